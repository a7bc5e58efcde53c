"""Host-side view of libahfgpu.so (include/ahfgpu.h) for Python callers: tests, bench.py, smoke().

The reference (NegriAndrea/AHF) is a C program; its own host code binds the same C-ABI directly
(INTEGRATION.md).  This module mirrors the reference's call order for the path --

    startrun -> [keys + sort, main.c:343-356] -> [gen_domgrids..gen_AMRhierarchy, main.c:616-648]
             -> (ahf_gridinfo / tree, host C) -> [halo loop, ahf_halos.c:504-510]

-- and keeps its names for parameters (AHF.input keys) and results (HALO / HALOPROFILE fields).
There is no CPU fallback: every method raises if the CUDA library is missing or reports an error.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libahfgpu.so")
NSCAL = 64
NPROFCOL = 25

H0 = 100.0
RHOC0 = 2.7755397e11
GRAV = 4.3006485e-9


class AhfGpuError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("device", C.c_int32), ("lgrid_dom", C.c_int32), ("lgrid_max", C.c_int32), ("min_part", C.c_int32),
                ("nth_dom", C.c_double), ("nth_ref", C.c_double), ("vesc_tune", C.c_double),
                ("r_fac", C.c_double), ("x_fac", C.c_double), ("v_fac", C.c_double), ("m_fac", C.c_double),
                ("rho_fac", C.c_double), ("phi_fac", C.c_double), ("hubble", C.c_double), ("ovlim", C.c_double),
                ("rho_vir", C.c_double)]


def make_params(*, boxsize: float, pmass: float, lgrid_dom: int, device: int = 0, lgrid_max: int = 1 << 21,
                nper_dom: float = 2.0, nper_ref: float = 2.5, vesc_tune: float = 1.5, nmin: int = 20,
                a: float = 1.0, omega0: float = 0.3, lambda0: float = 0.7, dvir: float = 200.0) -> Params:
    """AHF.input-example settings (RhoVir = 0: densities normalised to rho_crit; Dvir > 0 fixes ovlim) and the
    unit factors of reference src/libahf/ahf_halos.c:199-221 for a snapshot at expansion factor `a`."""
    t_unit = 1.0 / H0
    ez2 = omega0 / a ** 3 + (1.0 - omega0 - lambda0) / a ** 2 + lambda0        # (H/H0)^2
    hubble = H0 * np.sqrt(ez2)
    rho_crit_a = RHOC0 * ez2                                                     # physical rho_crit(a)
    return Params(device=device, lgrid_dom=lgrid_dom, lgrid_max=min(lgrid_max, 1 << 21), min_part=nmin,
                  nth_dom=nper_dom, nth_ref=nper_ref, vesc_tune=vesc_tune,
                  r_fac=boxsize * a, x_fac=boxsize, v_fac=boxsize / t_unit / a, m_fac=pmass,
                  rho_fac=pmass / boxsize ** 3, phi_fac=GRAV * pmass / (boxsize * a), hubble=hubble, ovlim=dvir,
                  rho_vir=a ** 3 * rho_crit_a)


def params_from_reference(glob: np.ndarray, *, lgrid_dom: int, device: int = 0, nper_dom: float = 2.0,
                          nper_ref: float = 2.5, lgrid_max: int = 1 << 21) -> Params:
    """Params from the 16 doubles the hooked reference dumps (oracle/ref_hooks.c dump_halos)."""
    return Params(device=device, lgrid_dom=lgrid_dom, lgrid_max=lgrid_max, min_part=int(glob[9]), nth_dom=nper_dom,
                  nth_ref=nper_ref, vesc_tune=glob[10], r_fac=glob[0], x_fac=glob[1], v_fac=glob[2], m_fac=glob[3],
                  rho_fac=glob[4], phi_fac=glob[5], hubble=glob[6], ovlim=glob[7], rho_vir=glob[8])


def min_ref(par: Params, l1dims, med_weight: float = 1.0) -> int:
    """ahf.min_ref: the first refinement level the patch colouring considers (reference src/libahf/ahf_gridinfo.c:147-175): the last
    level whose refinement threshold, expressed as an overdensity Nth_ref * pmass * med_weight / cell volume / rho_vir, is still below
    the virial overdensity (levels counted from the domain grid, start value 1; AHF_MIN_REF_OFFSET = 0, src/param.h:21).
    `med_weight` is simu.med_weight (1 for equal-mass runs; the heaviest species weight in a MULTIMASS run, src/startrun.c:663)."""
    start = 1
    for i, l1dim in enumerate(l1dims):
        refine_len = par.x_fac / float(l1dim)
        refine_ovdens = (par.nth_ref * (par.m_fac * med_weight) / (refine_len * refine_len * refine_len)) / par.rho_vir
        if refine_ovdens < par.ovlim:
            start = i
    return start


def tree_halos(stats_per_level, max_gather_rad: float, lists: bool = True) -> dict:
    """Host side of NEXT-2 (ahfgpu_tree_halos): refinement tree and halo seeds from the per-refinement tables ([niso, 18] per coloured
    level, first level = ahf.min_ref).  Returns daughter / close / sub (lists per level) and pos, gather_rad, npart, host per halo.
    lists=False: the per-refinement / per-halo Python lists (sub, halo_sub: what the tests compare) are left out -- at 2e4 haloes building
    them costs more than the tree itself."""
    L = lib()
    niso = np.array([len(s) for s in stats_per_level], np.int64)
    rows = int(niso.sum())
    st = np.ascontiguousarray(np.concatenate([np.asarray(s, np.float64).reshape(-1, 18) for s in stats_per_level]) if rows else np.zeros((0, 18)))
    dau = np.empty(max(rows, 1), np.int32); close = np.empty(max(rows, 1), np.float64); off = np.zeros(rows + 1, np.int64)
    sub = np.empty(max(rows, 1) * 2 + 8, np.int32); cap = rows + 1
    nh = C.c_int64(0)
    pos = np.zeros((cap, 3)); g = np.zeros(cap); npart = np.zeros(cap, np.int64); host = np.zeros(cap, np.int32)
    hlev = np.zeros(cap, np.int32); hsoff = np.zeros(cap + 1, np.int64); hsub = np.zeros(cap, np.int32)
    L.ahfgpu_tree_halos_ex.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    rc = L.ahfgpu_tree_halos_ex(len(niso), _p(niso), _p(st), max_gather_rad, _p(dau), _p(close), _p(off), _p(sub), len(sub), C.byref(nh),
                                _p(pos), _p(g), _p(npart), _p(host), cap, _p(hlev), _p(hsoff), _p(hsub), cap)
    if rc != 0:
        raise AhfGpuError(L.ahfgpu_last_error().decode())
    out = dict(daughter=[], close=[], sub=[], pos=pos[:nh.value].copy(), gather_rad=g[:nh.value].copy(), npart=npart[:nh.value].copy(),
               host=host[:nh.value].copy(), host_level=hlev[:nh.value].copy())
    if not lists:
        out["halo_sub_offset"] = hsoff[:nh.value + 1].copy(); out["halo_sub_flat"] = hsub[:int(hsoff[nh.value])].copy()
        return out
    out["halo_sub"] = [hsub[hsoff[i]:hsoff[i + 1]].copy() for i in range(nh.value)]
    r = 0
    for n in niso:
        out["daughter"].append(dau[r:r + n].astype(np.int64)); out["close"].append(close[r:r + n].copy())
        out["sub"].append([[int(v) for v in sub[off[q]:off[q + 1]]] for q in range(r, r + n)])
        r += int(n)
    return out


class _CatalogueIn(C.Structure):
    _fields_ = ([("nhalo", C.c_int64)] + [(k, C.c_void_p) for k in ("scal", "pos3", "member_off", "members", "prof_off", "prof", "species", "prof_species", "host",
                                                                     "host_level", "sub_off", "sub", "part_id", "part_weight", "part_u")]
                + [(k, C.c_double) for k in ("x_fac", "r_fac", "v_fac", "m_fac", "rho_fac", "phi_fac", "u_fac", "rho_vir", "pmass")]
                + [("min_part", C.c_int32), ("flags", C.c_int32)])


def catalogue_write(fprefix, scal, pos3, members, profiles, host, host_level, sub, part_id, fac: dict, min_part: int,
                    part_u=None, part_weight=None, species=None, prof_species=None) -> dict:
    """NEXT-3 (ahfgpu_catalogue_write): sub-halo re-hash, ordering and the four catalogue files.  members / profiles / sub: one array
    per halo (profiles [25, nbins] or None, prof_species [3, nbins] or None); fac: x_fac r_fac v_fac m_fac rho_fac phi_fac u_fac rho_vir
    pmass.  fprefix None: only the re-hash and the ordering.  Returns host / nsub after the re-hash and the rank (= halo ID) of every halo."""
    L = lib()
    nh = len(scal)
    keep = []

    def arr(a, dt):
        a = np.ascontiguousarray(a, dt); keep.append(a); return a.ctypes.data
    moff = np.zeros(nh + 1, np.int64); poff = np.zeros(nh + 1, np.int64); soff = np.zeros(nh + 1, np.int64)
    for i in range(nh):
        moff[i + 1] = moff[i] + len(members[i])
        poff[i + 1] = poff[i] + (0 if profiles[i] is None else profiles[i].shape[1])
        soff[i + 1] = soff[i] + len(sub[i])
    mem = np.concatenate([np.asarray(m, np.int64) for m in members]) if nh and moff[-1] else np.zeros(1, np.int64)
    prof = np.concatenate([np.asarray(p, np.float64).reshape(-1) for p in profiles if p is not None]) if poff[-1] else np.zeros(1)
    subs = np.concatenate([np.asarray(q, np.int32) for q in sub]) if soff[-1] else np.zeros(1, np.int32)
    cin = _CatalogueIn()
    cin.nhalo = nh
    cin.scal = arr(scal, np.float64); cin.pos3 = arr(pos3, np.float64); cin.member_off = arr(moff, np.int64); cin.members = arr(mem, np.int64)
    cin.prof_off = arr(poff, np.int64); cin.prof = arr(prof, np.float64)
    cin.host = arr(host, np.int32); cin.host_level = arr(host_level, np.int32); cin.sub_off = arr(soff, np.int64); cin.sub = arr(subs, np.int32)
    cin.part_id = arr(part_id, np.uint64)
    cin.part_u = arr(part_u, np.float32) if part_u is not None else None
    cin.part_weight = arr(part_weight, np.float32) if part_weight is not None else None
    flags = 0
    if species is not None:
        psp = np.concatenate([np.asarray(p, np.float64).reshape(-1) for p in prof_species if p is not None]) if poff[-1] else np.zeros(1)
        cin.species = arr(species, np.float64); cin.prof_species = arr(psp, np.float64)
        flags |= 1
    for k in ("x_fac", "r_fac", "v_fac", "m_fac", "rho_fac", "phi_fac", "u_fac", "rho_vir", "pmass"):
        setattr(cin, k, float(fac[k]))
    cin.min_part = int(min_part); cin.flags = flags
    ho = np.empty(max(nh, 1), np.int32); ns = np.empty(max(nh, 1), np.int32); rk = np.empty(max(nh, 1), np.int64)
    L.ahfgpu_catalogue_write.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = L.ahfgpu_catalogue_write(None if fprefix is None else fprefix.encode(), C.byref(cin), _p(ho), _p(ns), _p(rk))
    if rc != 0:
        raise AhfGpuError(L.ahfgpu_last_error().decode())
    return dict(host=ho[:nh].copy(), nsub=ns[:nh].copy(), rank=rk[:nh].copy())


_lib = None


def build(force: bool = False) -> None:
    """nvcc -gencode arch=compute_100a,code=sm_100a (cross-compiles without a GPU)."""
    src = os.path.join(HERE, "csrc")
    files = [os.path.join(src, f) for f in os.listdir(src) if f.endswith((".cu", ".cuh"))]
    files.append(os.path.join(HERE, "..", "include", "ahfgpu.h"))
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(f) for f in files)):
        return
    subprocess.check_call(["make", "-s", "-C", src, "-j4"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AhfGpuError(f"{LIB_PATH} not built (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.ahfgpu_last_error.restype = C.c_char_p
        L.ahfgpu_init.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Params)]
        L.ahfgpu_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
        L.ahfgpu_finalize.argtypes = [C.c_void_p]
        L.ahfgpu_sfc_sort_particles.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32] + [C.c_int32] * 6
        L.ahfgpu_sfc_sort_soa.argtypes = [C.c_void_p] * 5 + [C.c_uint64, C.c_void_p, C.c_void_p]
        L.ahfgpu_upload_soa.argtypes = [C.c_void_p] * 5 + [C.c_uint64]
        L.ahfgpu_sfc_sort_soa_async.argtypes = [C.c_void_p] * 5 + [C.c_uint64]
        L.ahfgpu_sfc_sort_resident.argtypes = [C.c_void_p]
        L.ahfgpu_event_record.argtypes = [C.c_void_p, C.c_int32]
        L.ahfgpu_event_elapsed_ms.restype = C.c_double
        L.ahfgpu_event_elapsed_ms.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.ahfgpu_synchronize.argtypes = [C.c_void_p]
        L.ahfgpu_device_ptr.restype = C.c_void_p
        L.ahfgpu_device_ptr.argtypes = [C.c_void_p, C.c_char_p]
        L.ahfgpu_set_global_count.argtypes = [C.c_void_p, C.c_uint64]
        L.ahfgpu_comm_nccl_unique_id.argtypes = [C.c_void_p]
        L.ahfgpu_comm_init_nccl.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.ahfgpu_comm_local_group_create.restype = C.c_void_p
        L.ahfgpu_comm_local_group_create.argtypes = [C.c_int32]
        L.ahfgpu_comm_local_group_destroy.argtypes = [C.c_void_p]
        L.ahfgpu_comm_local_group_abort.argtypes = [C.c_void_p]
        L.ahfgpu_comm_init_local.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.ahfgpu_slab_distribute.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_int32]
        L.ahfgpu_slab_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ahfgpu_slab_owner_of.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.ahfgpu_amr_level_owned.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.ahfgpu_particle_ids.argtypes = [C.c_void_p, C.c_void_p]
        L.ahfgpu_particle_ids_async.argtypes = [C.c_void_p, C.c_void_p]
        L.ahfgpu_halo_members_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.ahfgpu_ingest_gadget.argtypes = [C.c_void_p, C.c_char_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.ahfgpu_particles_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ahfgpu_input_peek.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.ahfgpu_ingest_prefetch.argtypes = [C.c_char_p]
        L.ahfgpu_adopt_sorted.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32, C.c_int32]
        L.ahfgpu_sfc_sort_device4.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32, C.c_int32]
        L.ahfgpu_hilbert_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
        L.ahfgpu_build_amr.argtypes = [C.c_void_p]
        L.ahfgpu_amr_nlevels.argtypes = [C.c_void_p]
        L.ahfgpu_amr_level_header.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.ahfgpu_amr_level_get.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 8
        L.ahfgpu_amr_patches.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ahfgpu_amr_patch_stats.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]
        L.ahfgpu_tree_halos.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.ahfgpu_amr_particle_levels.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.ahfgpu_construct_halos.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ahfgpu_halo_sizes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ahfgpu_halo_fetch.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.ahfgpu_stage_ms.restype = C.c_double
        L.ahfgpu_stage_ms.argtypes = [C.c_void_p, C.c_char_p]
        L.ahfgpu_stage_count.restype = C.c_int64
        L.ahfgpu_stage_count.argtypes = [C.c_void_p, C.c_char_p]
        _lib = L
    return _lib


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    if lib().ahfgpu_comm_nccl_unique_id(buf) != 0:
        raise AhfGpuError(lib().ahfgpu_last_error().decode())
    return buf.raw


def local_group_create(nranks: int) -> int:
    g = lib().ahfgpu_comm_local_group_create(nranks)
    if not g:
        raise AhfGpuError("could not create a local group")
    return int(g)


def local_group_abort(group: int) -> None:
    lib().ahfgpu_comm_local_group_abort(C.c_void_p(group))


def local_group_destroy(group: int) -> None:
    lib().ahfgpu_comm_local_group_destroy(C.c_void_p(group))


def exported_symbols() -> list[str]:
    """Every entry point include/ahfgpu.h declares."""
    hdr = open(os.path.join(HERE, "..", "include", "ahfgpu.h")).read()
    import re
    return sorted(set(re.findall(r"\b(ahfgpu_[a-z_0-9]+)\s*\(", hdr)))


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data if isinstance(a, np.ndarray) else int(a))


@dataclass
class GpuLevel:
    l1dim: int
    ncell: int
    npart_dep: int
    npart_final: int
    critdens: float
    masstopartdens: float
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray
    dens: np.ndarray
    runflags: np.ndarray
    interior: np.ndarray
    mark: np.ndarray
    count: np.ndarray

    def lin(self) -> np.ndarray:
        L = np.int64(self.l1dim)
        return (self.z.astype(np.int64) * L + self.y.astype(np.int64)) * L + self.x.astype(np.int64)


class AhfGpu:
    """One context = one device = one resident particle set."""

    def __init__(self, params: Params):
        self._L = lib()
        self._h = C.c_void_p()
        self.params = params
        self._chk(self._L.ahfgpu_init(C.byref(self._h), C.byref(params)))
        self.n = 0

    def _chk(self, rc: int):
        if rc != 0:
            raise AhfGpuError(self._L.ahfgpu_last_error().decode())

    def close(self):
        if self._h:
            self._L.ahfgpu_finalize(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_params(self, params: Params):
        self.params = params
        self._chk(self._L.ahfgpu_set_params(self._h, C.byref(params)))

    # ---- K1 / K2 -------------------------------------------------------------------------------
    def hilbert_keys(self, pos: np.ndarray, bits: int = 21) -> np.ndarray:
        pos = np.ascontiguousarray(pos, np.float32)
        keys = np.empty(pos.shape[0], np.uint64)
        self._chk(self._L.ahfgpu_hilbert_keys(self._h, _p(pos), pos.shape[0], bits, _p(keys)))
        return keys

    def sfc_sort(self, pos, mom, weight=None, u=None, want_keys=True, want_order=True, keys_out=None, order_out=None):
        """pos/mom: (N,3) float32 host arrays or raw host pointers (ints, e.g. pinned torch tensors' data_ptr)."""
        if isinstance(pos, np.ndarray):
            pos = np.ascontiguousarray(pos, np.float32); mom = np.ascontiguousarray(mom, np.float32)
            n = pos.shape[0]
        else:
            raise TypeError("use sfc_sort_ptr for raw pointers")
        weight = None if weight is None else np.ascontiguousarray(weight, np.float32)
        u = None if u is None else np.ascontiguousarray(u, np.float32)
        keys = keys_out if keys_out is not None else (np.empty(n, np.uint64) if want_keys else None)
        order = order_out if order_out is not None else (np.empty(n, np.uint32) if want_order else None)
        self._chk(self._L.ahfgpu_sfc_sort_soa(self._h, _p(pos), _p(mom), _p(weight), _p(u), n, _p(keys), _p(order)))
        self.n = n
        return keys, order

    def sfc_sort_ptr(self, pos_ptr: int, mom_ptr: int, n: int, keys_ptr: int = 0, order_ptr: int = 0):
        self._chk(self._L.ahfgpu_sfc_sort_soa(self._h, C.c_void_p(pos_ptr), C.c_void_p(mom_ptr), None, None, n,
                                              C.c_void_p(keys_ptr) if keys_ptr else None,
                                              C.c_void_p(order_ptr) if order_ptr else None))
        self.n = n

    def sfc_sort_async_ptr(self, pos_ptr: int, mom_ptr: int, n: int, weight_ptr: int = 0, u_ptr: int = 0):
        """Overlapped variant (ahfgpu_sfc_sort_soa_async): the momenta are still in flight on return; the host buffers behind
        mom_ptr / u_ptr must stay untouched until construct_halos() or synchronize() has returned."""
        self._chk(self._L.ahfgpu_sfc_sort_soa_async(self._h, C.c_void_p(pos_ptr), C.c_void_p(mom_ptr),
                                                    C.c_void_p(weight_ptr) if weight_ptr else None,
                                                    C.c_void_p(u_ptr) if u_ptr else None, n))
        self.n = n

    def ingest_gadget(self, path: str, posscale: float = 1.0, weightscale: float = 1.0, want_ids: bool = True, prefetch: bool = False):
        """bulk GADGET ingest with on-device unit scaling (NEXT-4): returns (info dict, ids or None); continue with sfc_sort_resident().
        prefetch: read the blocks through ahfgpu_ingest_prefetch first (the host-only half the drop-in program runs ahead of the context)"""
        if prefetch:
            self._chk(self._L.ahfgpu_ingest_prefetch(path.encode()))
        info = np.zeros(24, np.float64)
        self._chk(self._L.ahfgpu_ingest_gadget(self._h, path.encode(), posscale, weightscale, None, _p(info)))
        n = int(info[0])
        ids = None
        if want_ids:
            ids = np.empty(n, np.uint64)
            self._chk(self._L.ahfgpu_ingest_gadget(self._h, path.encode(), posscale, weightscale, _p(ids), _p(info)))
        self.n = n
        keys = ("n", "boxsize", "expansion", "omega0", "lambda0", "pmass", "shift_x", "shift_y", "shift_z", "scale_pos", "scale_mom", "version", "swapped",
                "hubble", "read_ms", "device_ms", "min_x", "min_y", "min_z", "max_x", "max_y", "max_z", "boxsize_file")
        return dict(zip(keys, info.tolist())), ids

    def particles(self):
        """the resident sorted particles: (pos4 [n,4] x y z weight, mom4 [n,4] px py pz u)"""
        pos4 = np.empty((self.n, 4), np.float32); mom4 = np.empty((self.n, 4), np.float32)
        self._chk(self._L.ahfgpu_particles_get(self._h, _p(pos4), _p(mom4)))
        return pos4, mom4

    def upload(self, pos, mom, weight=None, u=None):
        pos = np.ascontiguousarray(pos, np.float32); mom = np.ascontiguousarray(mom, np.float32)
        weight = None if weight is None else np.ascontiguousarray(weight, np.float32)
        u = None if u is None else np.ascontiguousarray(u, np.float32)
        self._chk(self._L.ahfgpu_upload_soa(self._h, _p(pos), _p(mom), _p(weight), _p(u), pos.shape[0]))
        self.n = pos.shape[0]

    def sfc_sort_resident(self):
        self._chk(self._L.ahfgpu_sfc_sort_resident(self._h))

    def event_record(self, slot: int):
        self._chk(self._L.ahfgpu_event_record(self._h, slot))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        return float(self._L.ahfgpu_event_elapsed_ms(self._h, a, b))

    def synchronize(self):
        self._chk(self._L.ahfgpu_synchronize(self._h))

    def sfc_sort_particles(self, part: np.ndarray, off_pos: int, off_mom: int, off_key: int, off_id: int,
                           off_weight: int = -1, off_u: int = -1):
        """`part`: structured/byte array laid out like the reference's struct particle[] (sorted in place)."""
        n, stride = part.shape[0], part.strides[0]
        self._chk(self._L.ahfgpu_sfc_sort_particles(self._h, _p(part), n, stride, off_pos, off_mom, off_key, off_id,
                                                    off_weight, off_u))
        self.n = n

    # ---- ONE box on several GPUs (include/ahfgpu.h "several GPUs working on ONE box") -----------
    def comm_init_nccl(self, rank: int, nranks: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._chk(self._L.ahfgpu_comm_init_nccl(self._h, rank, nranks, buf))

    def comm_init_local(self, rank: int, group: int):
        self._chk(self._L.ahfgpu_comm_init_local(self._h, rank, C.c_void_p(group)))

    def slab_distribute(self, id_base: int, ghost_width: float = 0.0, decomp_bits: int = 0):
        """keys + block histogram + equal-particle Hilbert ranges + ONE exchange (owner and ghost holders) + ONE sort of the particles
        uploaded with upload(); afterwards the context holds its key range plus the ghost shell"""
        self._chk(self._L.ahfgpu_slab_distribute(self._h, id_base, ghost_width, decomp_bits))
        self._split = True
        self.n = int(self.slab_info()["resident"])

    def slab_info(self) -> dict:
        io = np.zeros(12, np.int64); do = np.zeros(4, np.float64)
        self._chk(self._L.ahfgpu_slab_info(self._h, _p(io), _p(do)))
        keys = ("rank", "nranks", "resident", "own_lo", "own_hi", "n_total", "decomp_bits", "shell_blocks", "block_lo", "block_hi", "levels", "coll_calls")
        out = {k: int(v) for k, v in zip(keys, io)}
        out.update(ghost_width=float(do[0]), coll_ms=float(do[1]), coll_bytes=float(do[2]))
        return out

    def slab_owner_of(self, pos: np.ndarray) -> np.ndarray:
        pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
        owner = np.empty(pos.shape[0], np.int32)
        self._chk(self._L.ahfgpu_slab_owner_of(self._h, pos.shape[0], _p(pos), _p(owner)))
        return owner

    def level_owned(self, lev: int) -> np.ndarray:
        nc = int(self.level_header(lev)[0][1])
        owned = np.empty(nc, np.uint8)
        self._chk(self._L.ahfgpu_amr_level_owned(self._h, lev, _p(owned)))
        return owned.astype(bool)

    def particle_ids(self) -> np.ndarray:
        ids = np.empty(self.n, np.uint32)
        self._chk(self._L.ahfgpu_particle_ids(self._h, _p(ids)))
        return ids

    # ---- D / F / R / L -------------------------------------------------------------------------
    def build_amr(self) -> int:
        self._chk(self._L.ahfgpu_build_amr(self._h))
        return self._L.ahfgpu_amr_nlevels(self._h)

    def nlevels(self) -> int:
        return self._L.ahfgpu_amr_nlevels(self._h)

    def level_header(self, lev: int):
        io = np.zeros(4, np.int64); do = np.zeros(2, np.float64)
        self._chk(self._L.ahfgpu_amr_level_header(self._h, lev, _p(io), _p(do)))
        return io, do

    def level(self, lev: int, cells: bool = True) -> GpuLevel:
        io, do = self.level_header(lev)
        nc = int(io[1])
        x = np.empty(nc, np.int32); y = np.empty(nc, np.int32); z = np.empty(nc, np.int32)
        dens = np.empty(nc, np.float32); rf = np.empty(nc, np.uint8); it = np.empty(nc, np.uint8)
        mk = np.empty(nc, np.uint8); cnt = np.empty(nc, np.int32)
        if cells:
            self._chk(self._L.ahfgpu_amr_level_get(self._h, lev, _p(x), _p(y), _p(z), _p(dens), _p(rf), _p(it), _p(mk), _p(cnt)))
        else:
            self._chk(self._L.ahfgpu_amr_level_get(self._h, lev, None, None, None, _p(dens), None, None, _p(mk), _p(cnt)))
        return GpuLevel(int(io[0]), nc, int(io[2]), int(io[3]), float(do[0]), float(do[1]), x, y, z, dens, rf, it, mk, cnt)

    def patches(self, lev: int):
        """Patch colouring of ahf_gridinfo (src/libahf/ahf_gridinfo.c:236-577) on the device: (iso[ncell], periodic[niso, 3])."""
        nc = int(self.level_header(lev)[0][1])
        iso = np.empty(nc, np.int32); per = np.zeros((max(nc, 1), 3), np.uint8); niso = C.c_int64(0)
        self._chk(self._L.ahfgpu_amr_patches(self._h, lev, _p(iso), C.byref(niso), _p(per)))
        return iso, per[:niso.value].copy()

    def box_levels(self) -> int:
        """levels of the WHOLE box: a rank of a split box may hold fewer (its cells end on a coarser level)"""
        if getattr(self, "_split", False):
            return self.slab_info()["levels"]
        return self.nlevels()

    def min_ref(self, med_weight: float = 1.0) -> int:
        return min_ref(self.params, [int(self.params.lgrid_dom) << l for l in range(self.box_levels())], med_weight)

    def halo_seeds(self, max_gather_rad: float, med_weight: float = 1.0, lists: bool = True) -> dict:
        """From the resident hierarchy to the inputs of construct_halos without the reference's host-side mesh walk: first coloured
        level (ahf_gridinfo.c:147-175), per level patch labels + RefCentre tables on the device (ahfgpu_amr_patch_stats), then the tree
        and the seeds on the host (ahfgpu_tree_halos).  max_gather_rad = MaxGatherRad / boxsize."""
        m = self.min_ref(med_weight)
        stats = []
        for lev in range(m, self.box_levels()):             # split box: a collective call per level, every rank gets the same tables
            n = C.c_int64(0)
            self._chk(self._L.ahfgpu_amr_patch_stats(self._h, lev, C.byref(n), None, 0))
            stats.append(self.patch_stats(lev, n.value))
        out = tree_halos(stats, max_gather_rad, lists=lists)
        out["min_ref"] = m; out["stats"] = stats
        return out

    def patch_stats(self, lev: int, niso: int) -> np.ndarray:
        """RefCentre on the device (src/libahf/ahf_halos.c:935-1620): [niso, 18] per isolated refinement, columns as in include/ahfgpu.h."""
        st = np.zeros((max(niso, 1), 18), np.float64); n = C.c_int64(0)
        self._chk(self._L.ahfgpu_amr_patch_stats(self._h, lev, C.byref(n), _p(st), max(niso, 1)))
        assert n.value == niso, (n.value, niso)
        return st[:niso]

    def particle_levels(self, with_cells: bool = True):
        nl = self.nlevels()
        owner = np.empty(self.n, np.int8)
        cells = np.empty((nl, self.n), np.int32) if with_cells else None
        self._chk(self._L.ahfgpu_amr_particle_levels(self._h, _p(owner), _p(cells), nl))
        return owner, cells

    # ---- G / U / P -----------------------------------------------------------------------------
    def construct_halos(self, centres: np.ndarray, gather_rad: np.ndarray, seed_npart: np.ndarray | None = None,
                        fetch: bool = True):
        centres = np.ascontiguousarray(centres, np.float64); gather_rad = np.ascontiguousarray(gather_rad, np.float64)
        nh = gather_rad.shape[0]
        seed = None if seed_npart is None else np.ascontiguousarray(seed_npart, np.int64)
        self._chk(self._L.ahfgpu_construct_halos(self._h, nh, _p(centres), _p(gather_rad), _p(seed)))
        if not fetch:
            return None
        return self.fetch_halos(nh)

    def fetch_halos(self, nh: int, scal_only: bool = False, bufs: dict | None = None):
        """bufs: caller-owned host arrays to fetch into (e.g. numpy views of PINNED memory, so that the copies run at bus speed):
        'scal' (nh*64 f64), 'moff' / 'poff' (nh+1 i64), 'members' (i64), 'prof' (f64) -- each at least as large as needed"""
        tm = C.c_int64(); tb = C.c_int64()
        self._chk(self._L.ahfgpu_halo_sizes(self._h, C.byref(tm), C.byref(tb)))

        def take(name, count, dtype):
            if bufs is not None and name in bufs and bufs[name].size >= count:
                return bufs[name].reshape(-1)[:count]
            return np.empty(count, dtype)
        scal = take("scal", nh * NSCAL, np.float64).reshape(nh, NSCAL)
        if scal_only:
            self._chk(self._L.ahfgpu_halo_fetch(self._h, _p(scal), None, None, None, None))
            return dict(scal=scal)
        moff = take("moff", nh + 1, np.int64); poff = take("poff", nh + 1, np.int64)
        members = take("members", max(tm.value, 1), np.int64); prof = take("prof", max(tb.value, 1) * NPROFCOL, np.float64)
        self._chk(self._L.ahfgpu_halo_fetch(self._h, _p(scal), _p(moff), _p(members), _p(poff), _p(prof)))
        res = dict(scal=scal, member_offset=moff, members=members[:tm.value], prof_offset=poff, prof=prof[:tb.value * NPROFCOL])
        # GAS_PARTICLES build (particles carry u): HALO.gas_only / HALO.stars_only and the M_gas, M_star, u_gas profile columns
        species = np.empty((nh, 64), np.float64); psp = np.empty(max(tb.value, 1) * 3, np.float64)
        if self._L.ahfgpu_halo_fetch_species(self._h, _p(species), _p(psp)) == 0:
            res["species"] = species; res["prof_species"] = psp[:tb.value * 3]
        return res

    @staticmethod
    def halo_profile_species(res: dict, i: int):
        a, b = res["prof_offset"][i], res["prof_offset"][i + 1]
        nb = int(b - a)
        if nb == 0 or "prof_species" not in res:
            return None
        return res["prof_species"][a * 3:b * 3].reshape(3, nb)

    @staticmethod
    def halo_members(res: dict, i: int) -> np.ndarray:
        return res["members"][res["member_offset"][i]:res["member_offset"][i + 1]]

    @staticmethod
    def halo_profile(res: dict, i: int):
        a, b = res["prof_offset"][i], res["prof_offset"][i + 1]
        nb = int(b - a)
        if nb == 0:
            return None
        return res["prof"][a * NPROFCOL:b * NPROFCOL].reshape(NPROFCOL, nb)

    # ---- measurement ---------------------------------------------------------------------------
    def stage_ms(self, name: str) -> float:
        return float(self._L.ahfgpu_stage_ms(self._h, name.encode()))

    def stage_count(self, name: str) -> int:
        return int(self._L.ahfgpu_stage_count(self._h, name.encode()))

    def launches(self) -> int:
        return self.stage_count("launches")
