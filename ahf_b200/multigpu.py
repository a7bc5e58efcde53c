"""ONE box on several GPUs of one node: host-side driver of the slab decomposition in libahfgpu.so.

Everything on the data path is C++/CUDA behind the C-ABI (ahf_b200/csrc/slab.cu, comm.cu, mesh.cu):

    distribute      keys of what the rank read, histogram over the Hilbert cells of the decomposition level (all-reduce),
                    equal-particle key ranges (the reference's MPI load balancer, src/libutility/loadbalance.c:383), ONE
                    personalised exchange that sends every particle to its owner AND to the ranks whose range lies within the
                    ghost width of it (comm_dist_part + boundary duplication, src/comm.c:104-316, :324ff), ONE sort
    build_amr       every rank builds all levels over its own cells + ghost shell; per level one small all-gather (marked cells,
                    particles on the next level, row counts) and one all-gather of row keys
    construct_halos every rank serves the haloes whose centre lies in its key range: their gathering spheres are inside the
                    ghost shell, so the halo pass needs no communication; member lists come back as global input indices

Two transports: NCCL (one process per GPU; this module only hands the 128-byte NCCL id from rank 0 to the others through
torch.distributed) and an in-process group of host threads (`run_local`: 2, 4 or 8 "ranks" on ONE GPU, used by the tests).
There is no CPU path: every array lives on a GPU.
"""
from __future__ import annotations

import threading

import numpy as np

from . import ahf


class SlabRank:
    """One rank's share of ONE box."""

    def __init__(self, params: ahf.Params, rank: int, world: int, device: int, *, nccl_id: bytes | None = None, local_group: int | None = None):
        self.rank, self.world = rank, world
        p = ahf.Params.from_buffer_copy(bytes(params))
        p.device = device
        self.params = p
        self.g = ahf.AhfGpu(p)
        if local_group is not None:
            self.g.comm_init_local(rank, local_group)
        elif nccl_id is not None:
            self.g.comm_init_nccl(rank, world, nccl_id)
        else:
            raise ValueError("need nccl_id or local_group")
        self.mine = None

    def close(self):
        self.g.close()

    # ---- exchange ------------------------------------------------------------------------------------------
    def distribute(self, pos, mom, weight=None, u=None, *, id_base: int, ghost_width: float = 0.0, decomp_bits: int = 0):
        """pos / mom: what this rank read (any subset of the box, any order); id_base: global index of its first particle"""
        self.g.upload(pos, mom, weight, u)
        self.g.slab_distribute(id_base, ghost_width, decomp_bits)

    def distribute_ptr(self, pos_ptr: int, mom_ptr: int, n: int, *, id_base: int, ghost_width: float = 0.0, decomp_bits: int = 0):
        """same from raw (pinned) host pointers: the upload is part of the call"""
        import ctypes as C
        g = self.g
        g._chk(g._L.ahfgpu_upload_soa(g._h, C.c_void_p(pos_ptr), C.c_void_p(mom_ptr), None, None, n))
        g.slab_distribute(id_base, ghost_width, decomp_bits)

    def redistribute(self, *, id_base: int, ghost_width: float = 0.0, decomp_bits: int = 0):
        """again from the copy already uploaded (timing with the input resident in HBM)"""
        self.g.slab_distribute(id_base, ghost_width, decomp_bits)

    # ---- mesh ----------------------------------------------------------------------------------------------
    def build_amr(self) -> int:
        self.g.build_amr()
        return self.g.slab_info()["levels"]

    # ---- halo pass -----------------------------------------------------------------------------------------
    def construct_halos(self, centres: np.ndarray, gather_rad: np.ndarray, seed_npart: np.ndarray, fetch: bool = True, scal_only: bool = False):
        """the haloes whose centre this rank owns; returns (indices into the caller's list, result dict or None)"""
        info = self.g.slab_info()
        if len(gather_rad) and float(np.max(gather_rad)) > info["ghost_width"] * (1 + 1e-12):
            raise ahf.AhfGpuError("a gathering radius exceeds the ghost width the particles were distributed with")
        owner = self.g.slab_owner_of(centres)
        mine = np.nonzero(owner == self.rank)[0]
        self.mine = mine
        res = self.g.construct_halos(centres[mine], gather_rad[mine], seed_npart[mine], fetch=False)
        if fetch:
            res = self.g.fetch_halos(len(mine), scal_only=scal_only)
        return mine, res


def catalogue_from_ranks(fprefix, params: ahf.Params, seeds: dict, parts, part_id) -> dict:
    """The four catalogue files of ONE box that several ranks analysed (BASELINE.json configs[3]): `seeds` is what every rank derived
    (AhfGpu.halo_seeds: identical on all ranks, with the halo_sub lists), `parts` the ranks' (mine, res) pairs of
    SlabRank.construct_halos (fetched results; member lists are global input indices), `part_id` the ID of every particle of the box
    by input index.  Sub-halo re-hash, ordering and writers are the library's host code (ahfgpu_catalogue_write): the files equal the
    single-GPU ones byte for byte.  fprefix None: only re-hash and ordering."""
    nh = len(seeds["npart"])
    scal = np.zeros((nh, ahf.NSCAL)); members = [np.zeros(0, np.int64)] * nh; profs = [None] * nh
    seen = np.zeros(nh, np.int32)
    for mine, res in parts:
        for k, h in enumerate(mine):
            scal[h] = res["scal"][k]; members[h] = ahf.AhfGpu.halo_members(res, k).astype(np.int64); profs[h] = ahf.AhfGpu.halo_profile(res, k)
            seen[h] += 1
    if not np.all(seen == 1):
        raise ahf.AhfGpuError("every halo must be served by exactly one rank")
    a = params.r_fac / params.x_fac
    fac = dict(x_fac=params.x_fac, r_fac=params.r_fac, v_fac=params.v_fac, m_fac=params.m_fac, rho_fac=params.rho_fac, phi_fac=params.phi_fac,
               u_fac=(params.v_fac * a) ** 2, rho_vir=params.rho_vir, pmass=params.m_fac)
    profs = [p if scal[i, 9] >= params.min_part else None for i, p in enumerate(profs)]
    return ahf.catalogue_write(fprefix, scal, seeds["pos"], members, profs, seeds["host"], seeds["host_level"], seeds["halo_sub"], part_id, fac, params.min_part)


def gather_parts_torch(mine, res, dst: int = 0):
    """the ranks' halo results on rank `dst` (list of (mine, res) in rank order; None elsewhere) -- one process per GPU"""
    import torch.distributed as dist
    keep = {k: res[k] for k in ("scal", "members", "member_offset", "prof", "prof_offset") if k in res}
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object((np.asarray(mine), keep), out, dst=dst)
    return out


def nccl_id_via_torch(rank: int, device) -> bytes:
    """rank 0 makes the NCCL id, torch.distributed (already initialised by the caller) hands it to everybody"""
    import torch
    import torch.distributed as dist
    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(ahf.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())


def run_local(world: int, params: ahf.Params, fn, device: int = 0):
    """`world` ranks as host threads of this process on ONE device: fn(rank, SlabRank) -> result; returns the list of results.
    An exception in one rank aborts the group (the others leave their collectives with an error) and is re-raised."""
    group = ahf.local_group_create(world)
    results, errors = [None] * world, [None] * world

    def work(r):
        sb = None
        try:
            sb = SlabRank(params, r, world, device, local_group=group)
            results[r] = fn(r, sb)
        except BaseException as e:  # noqa: BLE001
            errors[r] = e
            ahf.local_group_abort(group)
        finally:
            if sb is not None:
                try:
                    sb.close()
                except Exception:  # noqa: BLE001
                    pass

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    ahf.local_group_destroy(group)
    first = [e for e in errors if e is not None and "aborted" not in str(e)] or [e for e in errors if e is not None]
    if first:
        raise first[0]
    return results
