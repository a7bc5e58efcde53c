"""One box on several GPUs of one node: one process per GPU, torch.distributed (NCCL over
NVLink / NVSwitch) for the plumbing, libahfgpu.so for all compute.

    slab exchange   every rank sorts the particles it read, equal-particle Hilbert key ranges are chosen from regular
                    samples (the reference's MPI load balancer uses a per-cell histogram, src/libutility/loadbalance.c:383)
                    and the particles travel to the owner of their key range with ONE all-to-all (comm.c:104-316)
    mesh            every rank holds the (small) cell structure of every level, deposits only its slab; the u64
                    fixed-point accumulators of each level are summed with an NCCL all-reduce -- the ghost-cell exchange,
                    exact and order independent, so every rank takes bit-identical refinement decisions
    halo pass       the sorted slabs are all-gathered once (48 B/particle fits every GPU), haloes are assigned
                    to ranks by greedy LPT on their seed particle count and each rank constructs its share

There is no CPU path here either: every tensor lives on the rank's GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import ahf, parallel


class _DevArray:
    """zero-copy torch view of device memory owned by libahfgpu (CUDA array interface)"""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def dev_tensor(ptr: int, shape, typestr: str, device) -> torch.Tensor:
    if int(np.prod(shape)) == 0:
        return torch.empty(tuple(shape), dtype={"<f4": torch.float32, "<i8": torch.int64}[typestr], device=device)
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device=device)


_ALLREDUCE_T = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64)


class SlabBox:
    """The path for ONE box split over `world` ranks."""

    def __init__(self, params: ahf.Params, rank: int, world: int, device: int):
        self.rank, self.world = rank, world
        self.device = torch.device("cuda", device)
        params.device = device
        self.g = ahf.AhfGpu(params)                 # mesh context: the rank's slab
        self.gh = ahf.AhfGpu(params)                # halo context: adopts the all-gathered box
        self.n_total = 0
        self._cb = _ALLREDUCE_T(self._allreduce)    # keep the ctypes thunk alive
        self._keep = []

    def close(self):
        self.gh.close(); self.g.close()

    # ------------------------------------------------------------------------------------------------ exchange
    def distribute(self, pos_local: np.ndarray, mom_local: np.ndarray):
        """pos_local / mom_local: the particles this rank read (any order).  Afterwards the rank holds its SFC slab,
        key sorted, resident in the mesh context."""
        g, L, dev = self.g, ahf.lib(), self.device
        n_local = int(pos_local.shape[0])
        if isinstance(pos_local, torch.Tensor):     # (pinned) host tensors: no staging copy inside the driver
            g.sfc_sort_ptr(pos_local.data_ptr(), mom_local.data_ptr(), n_local)
        else:
            g.sfc_sort(pos_local, mom_local, want_keys=False, want_order=False)
        pos4 = dev_tensor(L.ahfgpu_device_ptr(g._h, b"pos4"), (n_local, 4), "<f4", dev)
        mom4 = dev_tensor(L.ahfgpu_device_ptr(g._h, b"mom4"), (n_local, 4), "<f4", dev)
        keys = dev_tensor(L.ahfgpu_device_ptr(g._h, b"keys"), (n_local,), "<i8", dev)       # 63-bit keys: non-negative as int64
        if self.world == 1:
            self.n_total = n_local
            return
        tot = torch.tensor([n_local], device=dev, dtype=torch.int64)
        dist.all_reduce(tot)
        self.n_total = int(tot.item())
        # splitters from regular samples of every rank's sorted keys
        ns = 4096
        idx = (torch.arange(ns, device=dev, dtype=torch.int64) * max(n_local - 1, 0)) // (ns - 1)     # integer arithmetic: float32 linspace overflows 2^24
        samp = keys[idx] if n_local else torch.zeros(ns, dtype=torch.int64, device=dev)
        allsamp = [torch.empty_like(samp) for _ in range(self.world)]
        dist.all_gather(allsamp, samp)
        allsamp = torch.sort(torch.cat(allsamp)).values
        split = allsamp[(torch.arange(1, self.world, device=dev) * allsamp.numel()) // self.world]
        bounds = torch.searchsorted(keys, split).tolist() if n_local else [0] * (self.world - 1)
        bounds = [0] + bounds + [n_local]
        send_counts = [bounds[r + 1] - bounds[r] for r in range(self.world)]
        sc = torch.tensor(send_counts, device=dev, dtype=torch.int64)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc)
        recv_counts = rc.tolist()
        n_new = int(sum(recv_counts))
        rpos = torch.empty((n_new, 4), dtype=torch.float32, device=dev)
        rmom = torch.empty((n_new, 4), dtype=torch.float32, device=dev)
        dist.all_to_all_single(rpos, pos4.contiguous(), output_split_sizes=recv_counts, input_split_sizes=send_counts)
        dist.all_to_all_single(rmom, mom4.contiguous(), output_split_sizes=recv_counts, input_split_sizes=send_counts)
        torch.cuda.synchronize(dev)
        g._chk(L.ahfgpu_sfc_sort_device4(g._h, C.c_void_p(rpos.data_ptr()), C.c_void_p(rmom.data_ptr()), n_new, 0, 0))
        g.n = n_new
        del rpos, rmom

    # ------------------------------------------------------------------------------------------------ mesh
    def _allreduce(self, user, ptr, count):
        try:
            t = dev_tensor(ptr, (int(count),), "<i8", self.device)
            dist.all_reduce(t)                      # SUM of two's-complement words == sum of the u64 fixed-point accumulators
            torch.cuda.synchronize(self.device)
            return 0
        except Exception as e:                      # noqa: BLE001 -- reported through the C status
            print("all-reduce callback failed:", e)
            return 1

    def build_amr(self) -> int:
        g, L = self.g, ahf.lib()
        g._chk(L.ahfgpu_set_global_count(g._h, self.n_total))
        if self.world > 1:
            g._chk(L.ahfgpu_set_allreduce(g._h, C.cast(self._cb, C.c_void_p), None))
        return g.build_amr()

    # ------------------------------------------------------------------------------------------------ halo pass
    def gather_box(self):
        """all-gather the sorted slabs (rank order == key order) and adopt them in the halo context"""
        g, L, dev = self.g, ahf.lib(), self.device
        n_local = g.n
        pos4 = dev_tensor(L.ahfgpu_device_ptr(g._h, b"pos4"), (n_local, 4), "<f4", dev)
        mom4 = dev_tensor(L.ahfgpu_device_ptr(g._h, b"mom4"), (n_local, 4), "<f4", dev)
        keys = dev_tensor(L.ahfgpu_device_ptr(g._h, b"keys"), (n_local,), "<i8", dev)
        if self.world == 1:
            full = (pos4, mom4, keys)
        else:
            cnt = torch.tensor([n_local], device=dev, dtype=torch.int64)
            cnts = [torch.empty_like(cnt) for _ in range(self.world)]
            dist.all_gather(cnts, cnt)
            cnts = [int(c.item()) for c in cnts]
            nmax = max(cnts)

            def gather(t, width):
                pad = torch.zeros((nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
                pad[:n_local] = t
                out = torch.empty((self.world * nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
                dist.all_gather_into_tensor(out, pad)
                return torch.cat([out[r * nmax:r * nmax + cnts[r]] for r in range(self.world)]).contiguous()
            full = (gather(pos4, 4), gather(mom4, 4), gather(keys, 1))
        torch.cuda.synchronize(dev)
        self._keep = list(full)                     # the library only borrows these arrays
        n = int(full[2].shape[0])
        self.gh._chk(L.ahfgpu_adopt_sorted(self.gh._h, C.c_void_p(full[0].data_ptr()), C.c_void_p(full[1].data_ptr()),
                                           C.c_void_p(full[2].data_ptr()), n, 0, 0))
        self.gh.n = n
        return n

    def construct_halos(self, centres: np.ndarray, gather_rad: np.ndarray, seed_npart: np.ndarray) -> np.ndarray:
        """every rank gets the full (nhalo, 64) scalar table; member lists / profiles stay with the constructing rank"""
        owner = parallel.assign_halos_lpt(seed_npart, self.world)
        mine = np.nonzero(owner == self.rank)[0]
        scal = np.zeros((len(gather_rad), ahf.NSCAL))
        self.local_halos = mine
        self.local_result = None
        if len(mine):
            res = self.gh.construct_halos(centres[mine], gather_rad[mine], seed_npart[mine])
            scal[mine] = res["scal"]
            self.local_result = res
        if self.world > 1:
            t = torch.from_numpy(scal).to(self.device)
            dist.all_reduce(t)                      # disjoint rows: the sum is the union
            scal = t.cpu().numpy()
        return scal
