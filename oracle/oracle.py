"""TEST INFRASTRUCTURE: ctypes view of oracle/libahf_oracle.so (the CPU checker) and readers for the
dumps written by the hooked reference binary (oracle/_ref/ahf_ref, oracle/ref_hooks.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libahf_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "ahf_ref")
REF_BIN_MM = os.path.join(HERE, "_ref", "ahf_ref_mm")
REF_BIN_GT = os.path.join(os.path.dirname(REF_BIN), "ahf_ref_gt")      # -DAHFgridtreefile variant (writes .AHF_gridtree)

_lib = None


def build(force: bool = False) -> None:
    srcs = [os.path.join(HERE, f) for f in ("ahf_oracle_mesh.c", "ahf_oracle_halo.c", "ahf_oracle.h")]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return
    subprocess.check_call(["make", "-s", "-C", HERE, "libahf_oracle.so"])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_hilbert_key.restype = C.c_uint64
        L.orc_hilbert_key.argtypes = [C.c_double, C.c_double, C.c_double, C.c_uint]
        L.orc_hilbert_key_grid.restype = C.c_uint64
        L.orc_hilbert_key_grid.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint]
        L.orc_hilbert_keys.argtypes = [C.c_void_p, C.c_int64, C.c_uint, C.c_void_p]
        L.orc_hilbert_coords.argtypes = [C.c_uint64, C.c_uint, C.c_void_p]
        L.orc_argsort_keys.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_hier_build.restype = C.c_void_p
        L.orc_hier_build.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_double]
        L.orc_hier_nlevels.argtypes = [C.c_void_p]
        L.orc_hier_level_header.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_hier_level_get.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 11
        L.orc_hier_free.argtypes = [C.c_void_p]
        L.orc_hier_face_neighbours.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_hier_patch_centres.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_hier_patches.restype = C.c_int64
        L.orc_hier_patches.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        if hasattr(L, "orc_halo_construct"):
            L.orc_halo_construct.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
            L.orc_halo_result_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------------------------------------
def hilbert_keys(pos: np.ndarray, bits: int = 21) -> np.ndarray:
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    keys = np.empty(pos.shape[0], dtype=np.uint64)
    lib().orc_hilbert_keys(_p(pos), pos.shape[0], bits, _p(keys))
    return keys


def argsort_keys(keys: np.ndarray) -> np.ndarray:
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    order = np.empty(keys.shape[0], dtype=np.int64)
    lib().orc_argsort_keys(_p(keys), keys.shape[0], _p(order))
    return order


@dataclass
class Level:
    l1dim: int
    ncell: int
    critdens: float
    masstopartdens: float
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray
    dens: np.ndarray
    runflags: np.ndarray
    cnt_flag: np.ndarray
    plist_flag: np.ndarray
    interior: np.ndarray | None = None
    mark: np.ndarray | None = None
    cnt_final: np.ndarray | None = None
    plist_final: np.ndarray | None = None
    iso: np.ndarray | None = None               # isolated-refinement index per cell (ahf_gridinfo colouring), build_hierarchy(patches=True)
    iso_periodic: np.ndarray | None = None      # [niso, 3] periodic flags of the isolated refinements
    nb6: np.ndarray | None = None               # [ncell, 6] visible face neighbours (x-1, x+1, y-1, y+1, z-1, z+1), -1 = none
    patch: np.ndarray | None = None             # [niso, 12] RefCentre: numNodes, numParts, centre(3) (= particle centre of mass, AHFcomcentre), maxDens, centreGEOM(3), centreDens(3)

    def lin(self) -> np.ndarray:
        L = np.int64(self.l1dim)
        return (self.z.astype(np.int64) * L + self.y.astype(np.int64)) * L + self.x.astype(np.int64)


def build_hierarchy(pos_sorted: np.ndarray, lgrid_dom: int, lgrid_max: int = 1 << 21, nth_dom: float = 2.0,
                    nth_ref: float = 2.5, patches: bool = False) -> list[Level]:
    pos_sorted = np.ascontiguousarray(pos_sorted, dtype=np.float32)
    L = lib()
    h = L.orc_hier_build(_p(pos_sorted), pos_sorted.shape[0], lgrid_dom, lgrid_max, nth_dom, nth_ref)
    out = []
    try:
        for lev in range(L.orc_hier_nlevels(h)):
            io = np.zeros(4, dtype=np.int64)
            do = np.zeros(2, dtype=np.float64)
            L.orc_hier_level_header(h, lev, _p(io), _p(do))
            nc, nf, nfin = int(io[1]), int(io[2]), int(io[3])
            x = np.empty(nc, np.int32); y = np.empty(nc, np.int32); z = np.empty(nc, np.int32)
            dens = np.empty(nc, np.float32); rf = np.empty(nc, np.uint8); it = np.empty(nc, np.uint8)
            mk = np.empty(nc, np.uint8); cf = np.empty(nc, np.int32); pf = np.empty(max(nf, 1), np.int64)
            cfin = np.empty(nc, np.int32); pfin = np.empty(max(nfin, 1), np.int64)
            L.orc_hier_level_get(h, lev, _p(x), _p(y), _p(z), _p(dens), _p(rf), _p(it), _p(mk), _p(cf), _p(pf),
                                 _p(cfin), _p(pfin))
            out.append(Level(int(io[0]), nc, float(do[0]), float(do[1]), x, y, z, dens, rf, cf, pf[:nf], it, mk,
                             cfin, pfin[:nfin]))
            if patches and lev > 0:             # the reference colours refinement levels only (from ahf.min_ref >= 1)
                iso = np.empty(nc, np.int32); per = np.zeros((nc, 3), np.uint8)
                niso = int(L.orc_hier_patches(h, lev, _p(iso), _p(per)))
                if niso < 0:
                    raise RuntimeError(f"level {lev}: a cell joins more than three colours (undefined in the reference, ahf_gridinfo.c:1246)")
                out[-1].iso = iso; out[-1].iso_periodic = per[:niso].copy()
                nb6 = np.empty((nc, 6), np.int64)
                L.orc_hier_face_neighbours(h, lev, _p(nb6))
                out[-1].nb6 = nb6
                pc = np.zeros((max(niso, 1), 12), np.float64)
                per_c = np.ascontiguousarray(per[:max(niso, 1)])
                L.orc_hier_patch_centres(h, lev, _p(iso), _p(per_c), niso, _p(pc))
                out[-1].patch = pc[:niso]
    finally:
        L.orc_hier_free(h)
    return out


# ------------------------------------------------------------------------------------------------
# readers for the reference dumps
@dataclass
class RefParticles:
    n: int
    boxsize: float
    pmass: float
    t_unit: float
    a: float
    omega0: float
    lambda0: float
    no_vpart: float
    ids: np.ndarray
    keys: np.ndarray
    pos: np.ndarray
    mom: np.ndarray
    weight: np.ndarray | None = None
    u: np.ndarray | None = None


def read_particles(path: str, multimass: bool = False) -> RefParticles:
    with open(path, "rb") as f:
        n = int(np.fromfile(f, np.uint64, 1)[0])
        sc = np.fromfile(f, np.float64, 8)
        ids = np.fromfile(f, np.uint64, n)
        keys = np.fromfile(f, np.uint64, n)
        pos = np.fromfile(f, np.float32, 3 * n).reshape(n, 3)
        mom = np.fromfile(f, np.float32, 3 * n).reshape(n, 3)
        w = u = None
        if multimass:
            w = np.fromfile(f, np.float32, n)
            u = np.fromfile(f, np.float32, n)
    return RefParticles(n, sc[0], sc[1], sc[2], sc[3], sc[4], sc[5], sc[6], ids, keys, pos, mom, w, u)


def read_patches(path: str):
    """patches.bin of oracle/ref_hooks.c dump_patches(): (min_ref, [(iso[ncell], periodic[niso,3]) per coloured level])"""
    with open(path, "rb") as f:
        min_ref, ngrids = (int(v) for v in np.fromfile(f, np.int32, 2))
        out = []
        for _ in range(max(ngrids, 0)):
            nc, niso = (int(v) for v in np.fromfile(f, np.int64, 2))
            iso = np.fromfile(f, np.int32, nc)
            per = np.fromfile(f, np.int8, 3 * niso).reshape(niso, 3)
            out.append((iso, per))
    return min_ref, out


def read_gridtree(path: str):
    """<prefix>.AHF_gridtree of the reference built with -DAHFgridtreefile (ahf_halos.c:3066-3120):
    (min_ref, {level: dict(centre[n,3], close[n], nodes[n], parts[n], daughter[n,2], sub=[list of (level, index)])})"""
    with open(path) as f:
        tok = f.read().split()
    it = iter(tok)
    min_ref, ngrids = int(next(it)), int(next(it))
    out = {}
    for _ in range(ngrids):
        lev, n = int(next(it)), int(next(it))
        cen = np.zeros((n, 3)); close = np.zeros(n); nodes = np.zeros(n, np.int64); parts = np.zeros(n, np.int64)
        dau = np.zeros((n, 2), np.int64); sub = []
        for j in range(n):
            cen[j] = [float(next(it)) for _ in range(3)]; close[j] = float(next(it))
            nodes[j] = int(next(it)); parts[j] = int(next(it)); dau[j] = [int(next(it)), int(next(it))]
            ns = int(next(it))
            sub.append([(int(next(it)), int(next(it))) for _ in range(ns)])
        out[lev] = dict(centre=cen, close=close, nodes=nodes, parts=parts, daughter=dau, sub=sub)
    return min_ref, out


# ------------------------------------------------------------------------------------------------
# NEXT-2, second half: extents of the isolated refinements (end of RefCentre) and the refinement tree (analyseRef).
# Patch-level work (tens to thousands of patches): numpy for the per-cell extents, plain loops over patches for the tree.
# ------------------------------------------------------------------------------------------------
def patch_extents(lv: Level) -> np.ndarray:
    """[niso, 3, 2] (min, max) per dimension as RefCentre leaves them (src/libahf/ahf_halos.c:1400-1620): a periodic refinement is
    cut at boundRefDiv = fmod(centreDens + 1/2, 1) (:1413-1470; the radius term cancels in (a+b)/2) and keeps max over the nodes
    below the cut, min over those above (MinMaxBound, src/libutility/specific.c:227-254), so max < min there; untouched sentinels
    become 0 / 1 (:1597-1612)."""
    n = len(lv.patch)
    L = float(lv.l1dim)
    cd = lv.patch[:, 9:12]
    per = lv.iso_periodic.astype(bool)
    vol = lv.patch[:, 0] * ((1.0 / L) * (1.0 / L) * (1.0 / L))
    rad = np.array([((3.0 * v) / (4 * 3.14159265358979323846)) ** 0.333333333 * 1.1 for v in vol])
    div = np.full((n, 3), -1.0)
    for j in range(n):
        if per[j].any():
            for d in range(3):
                if per[j, d]:
                    a = rad[j] + cd[j, d]; b = 1.0 - rad[j] + cd[j, d]
                    div[j, d] = np.fmod((a + b) / 2.0, 1.0)
    ext = np.zeros((n, 3, 2))
    shift = 0.5 / L
    for d, c in enumerate((lv.x, lv.y, lv.z)):
        xx = np.fmod(c.astype(np.float64) / L + shift + 1.0, 1.0)
        mn = np.full(n, 100000.0); mx = np.full(n, -100000.0)
        dv = div[lv.iso, d]
        plain = dv < 0.0
        np.minimum.at(mn, lv.iso[plain], xx[plain]); np.maximum.at(mx, lv.iso[plain], xx[plain])
        low = ~plain & (xx < dv); up = ~plain & ~(xx < dv)
        np.maximum.at(mx, lv.iso[low], xx[low]); np.minimum.at(mn, lv.iso[up], xx[up])
        mn[mn == 100000.0] = 0.0; mx[mx == -100000.0] = 1.0
        ext[:, d, 0] = mn; ext[:, d, 1] = mx
    return ext


def _pdist2(a, b):
    d = np.abs(a - b)
    d = np.where(d > 0.5, 1.0 - d, d)
    return d[0] * d[0] + d[1] * d[1] + d[2] * d[2]


def patch_tree(levels: list[Level]):
    """analyseRef (src/libahf/ahf_halos.c:1652-2300) over the coloured levels (levels[0] = ahf.min_ref), default switches of the shipped
    define.h (PARDAU_PARTS).  Returns per level dict(sub=[lists of indices on the next level], daughter=[index or -1], close=[closeRefDist]).
    Steps as the reference takes them: (1) a finer refinement is listed under every coarser one whose [min,max] box holds its density
    centre (:1693-1800; periodic boxes have max < min); (2) one with several parents keeps the closest (first minimum) and is struck
    from the others -- only for levels 1 .. n-2, the loop bound of :1817 leaves the finest level alone; (3) one without a parent is
    given to the closest refinement of the level above and inherits its centres (:2030-2165); (4) the main branch is the listed
    refinement with most particles (first maximum, :2190-2235), the others get closeRefDist = half the distance to their nearest
    sibling (:2245-2285)."""
    n = len(levels)
    cd = [lv.patch[:, 9:12].copy() for lv in levels]
    ext = [patch_extents(lv) for lv in levels]
    sub = [[[] for _ in range(len(lv.patch))] for lv in levels]
    par = [[[] for _ in range(len(lv.patch))] for lv in levels]
    detail = False
    for i in range(n - 1):
        for j in range(len(sub[i])):
            for k in range(len(sub[i + 1])):
                ok = True
                for d in range(3):
                    lo, hi = ext[i][j, d]; v = cd[i + 1][k, d]
                    if lo < hi:
                        inside = (v > lo) and (v < hi)
                    else:
                        inside = ((v >= 0) and (v < hi)) or ((v > lo) and (v <= 1.0))
                    if not inside:
                        ok = False
                        break
                if ok:
                    sub[i][j].append(k); par[i + 1][k].append(j)
                    if len(par[i + 1][k]) > 1:
                        detail = True
    if detail:
        for i in range(1, n - 1):
            for j in range(len(par[i])):
                if len(par[i][j]) > 1:
                    best, tmin = -1, 10000000000000.0
                    for q in par[i][j]:
                        dist = _pdist2(cd[i][j], cd[i - 1][q])
                        if dist < tmin:
                            best, tmin = q, dist
                    for q in par[i][j]:
                        if q != best:
                            sub[i - 1][q] = [t for t in sub[i - 1][q] if t != j]
                    par[i][j] = [best]
    for i in range(1, n):
        for j in range(len(par[i])):
            if len(par[i][j]) == 0:
                best, tmin = -1, 10000000000000.0
                for q in range(len(cd[i - 1])):
                    dist = _pdist2(cd[i][j], cd[i - 1][q])
                    if dist < tmin:
                        best, tmin = q, dist
                par[i][j] = [best]; sub[i - 1][best].append(j)
                cd[i][j] = cd[i - 1][best]
    out = []
    close = [np.full(len(lv.patch), -1.0) for lv in levels]
    for i in range(n):
        dau = np.full(len(sub[i]), -1, np.int64)
        if i < n - 1:
            for j, sl in enumerate(sub[i]):
                if len(sl) > 1:
                    best, mp = -1, -1
                    for k in sl:
                        if levels[i + 1].patch[k, 1] > mp:
                            best, mp = k, levels[i + 1].patch[k, 1]
                    dau[j] = best
                    for a, k in enumerate(sl):
                        if k != best:
                            tmin = 10000000000000.0
                            for b, l in enumerate(sl):
                                if a != b:
                                    tmin = min(tmin, _pdist2(cd[i + 1][k], cd[i + 1][l]))
                            close[i + 1][k] = 0.5 * np.sqrt(tmin)
                elif len(sl) == 1:
                    dau[j] = sl[0]
        out.append(dict(sub=sub[i], daughter=dau, npar=[len(p) for p in par[i]]))
    for i in range(n):
        out[i]["close"] = close[i]
    return out


def tree_to_halos(levels: list[Level], tree, max_gather_rad: float):
    """spatialRef2halos (src/libahf/ahf_halos.c:2405-3058): the halo seeds the per-halo pass starts from.  Walks the refinement tree
    level by level as the reference does: a refinement without finer structure closes its halo (position = its centre, :2548-2570,
    :2680-2700), one with a single finer refinement hands its halo down the main branch, one with several opens a sub-halo for every
    listed refinement but the main-branch daughter (first guess R_vir = closeRefDist, :2620-2650, :2760-2790); particle counts add up
    along the main branch.  Gathering radius (:2985-3052): half the distance to the nearest halo with MORE particles (MaxGatherRad
    when there is none), at least R_vir, at most min(MaxGatherRad / boxsize, 1/4).  Returns (pos[nh,3], gatherRad[nh], npart[nh],
    hostHalo[nh]) in the order of the reference's halos[] array."""
    n = len(levels)
    niso = [len(lv.patch) for lv in levels]
    hidx = [np.full(k, -1, np.int64) for k in niso]
    pos, npart, rvir, host = [], [], [], []

    def new():
        pos.append(np.zeros(3)); npart.append(0); rvir.append(-1.0); host.append(-1)
        return len(pos) - 1
    expect = 0
    for i in range(n):
        for j in range(niso[i]):
            ns = len(tree[i]["sub"][j])
            expect += (1 if ns == 0 else ns) if i == 0 else (ns - 1 if ns > 1 else 0)
    for i in range(n):
        for j in range(niso[i]):
            sub = tree[i]["sub"][j]; dau = int(tree[i]["daughter"][j]); ns = len(sub)
            cen = levels[i].patch[j, 2:5]; np_ = int(levels[i].patch[j, 1])
            if i == 0:
                h = new()
                npart[h] = np_
                if ns == 0:
                    pos[h] = cen.copy()
            else:
                if tree[i]["npar"][j] == 0:
                    continue
                h = int(hidx[i][j])
                if h < 0:
                    raise RuntimeError("refinement without a halo (the reference exits here, ahf_halos.c:2676)")
                npart[h] += np_
                if ns != 1:
                    pos[h] = cen.copy()
            if ns >= 1 and dau != -1:
                hidx[i + 1][dau] = h
            if ns > 1:
                for k in sub:
                    if k != dau:
                        c = new()
                        host[c] = h; hidx[i + 1][k] = c; rvir[c] = float(tree[i + 1]["close"][k])
    while len(pos) < expect:
        new()
    nh = len(pos)
    P = np.array(pos).reshape(nh, 3); N = np.array(npart, np.int64); R = np.array(rvir)
    maxg = min(max_gather_rad, 0.25)
    G = np.empty(nh)
    for i in range(nh):
        more = N > N[i]
        if more.any():
            d = np.abs(P[more] - P[i]); d = np.where(d > 0.5, 1.0 - d, d)
            g = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]).min()) * 0.5
        else:
            g = maxg
        G[i] = min(max(g, R[i]), maxg)
    return P, G, N, np.array(host, np.int64)


def read_level(path: str) -> Level:
    with open(path, "rb") as f:
        hdr = np.fromfile(f, np.int64, 4)
        dh = np.fromfile(f, np.float64, 2)
        nc, npart = int(hdr[1]), int(hdr[2])
        x = np.fromfile(f, np.int32, nc); y = np.fromfile(f, np.int32, nc); z = np.fromfile(f, np.int32, nc)
        dens = np.fromfile(f, np.float32, nc)
        flg = np.fromfile(f, np.uint8, nc)
        cnt = np.fromfile(f, np.int32, nc)
        pl = np.fromfile(f, np.int64, npart)
    return Level(int(hdr[0]), nc, float(dh[0]), float(dh[1]), x, y, z, dens, flg, cnt, pl)


@dataclass
class RefHalos:
    n: int
    glob: np.ndarray            # 16 doubles: r_fac x_fac v_fac m_fac rho_fac phi_fac Hubble ovlim rho_vir minpart vtune maxgather a
    s: np.ndarray               # (n, 64) scalar slots, layout in oracle/ref_hooks.c dump_halos()
    members: list = field(default_factory=list)
    prof: list = field(default_factory=list)    # per halo (25, nbins) or None
    species: np.ndarray | None = None           # (n, 64): gas_only at 0, stars_only at 32 (GAS_PARTICLES build)
    prof_species: list | None = None            # per halo (3, nbins): M_gas, M_star, u_gas


def read_halos(dump_dir: str) -> RefHalos:
    with open(os.path.join(dump_dir, "halos.bin"), "rb") as f:
        hdr = np.fromfile(f, np.int64, 2)
        g = np.fromfile(f, np.float64, 16)
        n, ns = int(hdr[0]), int(hdr[1])
        s = np.fromfile(f, np.float64, n * ns).reshape(n, ns)
    members, prof = [], []
    with open(os.path.join(dump_dir, "halo_ipart.bin"), "rb") as f:
        for _ in range(n):
            m = int(np.fromfile(f, np.int64, 1)[0])
            members.append(np.fromfile(f, np.int64, m))
    with open(os.path.join(dump_dir, "halo_prof.bin"), "rb") as f:
        for _ in range(n):
            nb = int(np.fromfile(f, np.int64, 1)[0])
            prof.append(np.fromfile(f, np.float64, 25 * nb).reshape(25, nb) if nb > 0 else None)
    rh = RefHalos(n, g, s, members, prof)
    sp_path = os.path.join(dump_dir, "halo_species.bin")
    if os.path.exists(sp_path):                     # GAS_PARTICLES build: gas_only / stars_only blocks and the M_gas, M_star, u_gas profile columns
        rh.species = np.fromfile(sp_path, np.float64).reshape(n, 64)
        rh.prof_species = []
        with open(os.path.join(dump_dir, "halo_prof_species.bin"), "rb") as f:
            for _ in range(n):
                nb = int(np.fromfile(f, np.int64, 1)[0])
                rh.prof_species.append(np.fromfile(f, np.float64, 3 * nb).reshape(3, nb) if nb > 0 else None)
    return rh


def read_halo_tree(dump_dir: str) -> dict:
    """halo_tree.bin of oracle/ref_hooks.c dump_halo_tree(): hostHalo / hostHaloLevel / subStruct[] of every halo as spatialRef2halos
    left them (the inputs of the sub-halo re-hash, ahf_halos.c:550-640) and hostHalo / subStruct[] after it."""
    a = np.fromfile(os.path.join(dump_dir, "halo_tree.bin"), np.int32)
    n = int(a[0]); k = 1
    host_pre = np.empty(n, np.int32); level_pre = np.empty(n, np.int32); host_post = np.empty(n, np.int32)
    sub_pre, sub_post = [], []
    for i in range(n):
        host_pre[i], level_pre[i], ns = a[k], a[k + 1], int(a[k + 2]); k += 3
        sub_pre.append(a[k:k + ns].copy()); k += ns
        host_post[i], ns = a[k], int(a[k + 1]); k += 2
        sub_post.append(a[k:k + ns].copy()); k += ns
    assert k == len(a)
    return dict(host_pre=host_pre, level_pre=level_pre, sub_pre=sub_pre, host_post=host_post, sub_post=sub_post)


def run_reference(ahf_input: str, dump_dir: str | None = None, threads: int | None = None, multimass: bool = False,
                  cwd: str | None = None) -> dict:
    """Run the hooked, otherwise unmodified reference binary; returns the REFHOOK_TIMING fields."""
    env = dict(os.environ)
    if dump_dir is not None:
        os.makedirs(dump_dir, exist_ok=True)
        env["AHF_DUMP_DIR"] = dump_dir
    else:
        env.pop("AHF_DUMP_DIR", None)
    if threads is not None:
        env["OMP_NUM_THREADS"] = str(threads)
    exe = REF_BIN_MM if multimass else REF_BIN
    if not os.path.exists(exe):
        raise FileNotFoundError(f"{exe} missing: run oracle/build_ref.sh where /root/reference exists")
    pr = subprocess.run([exe, ahf_input], cwd=cwd or os.path.dirname(ahf_input), env=env, capture_output=True, text=True)
    if pr.returncode != 0:
        raise RuntimeError(f"reference failed ({pr.returncode}): {pr.stderr[-2000:]}")
    out = {}
    for line in pr.stderr.splitlines():
        if line.startswith("REFHOOK_TIMING"):
            for tok in line.split()[1:]:
                k, v = tok.split("=")
                out[k] = float(v)
    return out


# ------------------------------------------------------------------------------------------------
class _HaloParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("r_fac", "x_fac", "v_fac", "m_fac", "rho_fac", "phi_fac", "Hubble", "ovlim",
                                          "rho_vir", "vesc_tune")] + [("min_part", C.c_int)]


class _HaloResult(C.Structure):
    _fields_ = [("npart", C.c_int64), ("ipart", C.POINTER(C.c_int64)), ("n_gather", C.c_int64), ("n_rvir0", C.c_int64),
                ("n_unbound", C.c_int64), ("n_rvir1", C.c_int64), ("s", C.c_double * 64), ("nbins", C.c_int),
                ("prof", C.POINTER(C.c_double)), ("species", C.c_double * 64), ("prof_species", C.POINTER(C.c_double))]


def params_from_glob(g: np.ndarray) -> dict:
    """glob = the 16 doubles written by ref_hooks.c dump_halos()"""
    return dict(r_fac=g[0], x_fac=g[1], v_fac=g[2], m_fac=g[3], rho_fac=g[4], phi_fac=g[5], Hubble=g[6], ovlim=g[7],
                rho_vir=g[8], min_part=int(g[9]), vesc_tune=g[10])


def construct_halos(keys, pos, mom, weight, u, par: dict, centres, gather_rad, seed_npart=None) -> list[dict]:
    keys = np.ascontiguousarray(keys, np.uint64); pos = np.ascontiguousarray(pos, np.float32)
    mom = np.ascontiguousarray(mom, np.float32)
    weight = None if weight is None else np.ascontiguousarray(weight, np.float32)
    u = None if u is None else np.ascontiguousarray(u, np.float32)
    hp = _HaloParams(**{k: par[k] for k in ("r_fac", "x_fac", "v_fac", "m_fac", "rho_fac", "phi_fac", "Hubble", "ovlim",
                                            "rho_vir", "vesc_tune")}, min_part=int(par["min_part"]))
    L = lib()
    out = []
    for i in range(len(gather_rad)):
        if seed_npart is not None and seed_npart[i] == 0:      # ahf_halos_sfc.c:122 -- nothing to construct
            out.append(dict(npart=0, ipart=np.empty(0, np.int64), n_gather=0, n_rvir0=0, n_unbound=0, n_rvir1=0,
                            s=np.zeros(64), prof=None, nbins=0))
            continue
        r = _HaloResult()
        c = np.ascontiguousarray(centres[i], np.float64)
        L.orc_halo_construct(_p(keys), _p(pos), _p(mom), _p(weight), _p(u), keys.shape[0], C.byref(hp), _p(c),
                             float(gather_rad[i]), C.byref(r))
        ip = np.ctypeslib.as_array(r.ipart, shape=(max(int(r.npart), 1),))[:int(r.npart)].copy() if r.npart > 0 else np.empty(0, np.int64)
        pr = None
        if r.nbins > 0:
            pr = np.ctypeslib.as_array(r.prof, shape=(25 * r.nbins,)).copy().reshape(25, r.nbins)
        ps = None
        if r.nbins > 0 and u is not None and bool(r.prof_species):
            ps = np.ctypeslib.as_array(r.prof_species, shape=(3 * r.nbins,)).copy().reshape(3, r.nbins)
        out.append(dict(npart=int(r.npart), ipart=ip, n_gather=int(r.n_gather), n_rvir0=int(r.n_rvir0),
                        n_unbound=int(r.n_unbound), n_rvir1=int(r.n_rvir1), s=np.array(r.s[:]), prof=pr, nbins=int(r.nbins),
                        species=np.array(r.species[:]), prof_species=ps))
        L.orc_halo_result_free(C.byref(r))
    return out
