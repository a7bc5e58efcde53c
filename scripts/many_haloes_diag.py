"""Diagnostic: halo seeds + halo pass on the 256^3 box with 2e4 clumps; wall-clock split of the seed stage, CUDA-event stage times."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from ahf_b200 import ahf, synth
n1d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ncl = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
box = synth.make_box(n1d, seed=44, n_clumps=ncl)
par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d)
with ahf.AhfGpu(par) as g:
    g.upload(box.pos, box.mom); g.sfc_sort_resident(); g.build_amr()
    m = g.min_ref()
    import ctypes as C
    t0 = time.perf_counter(); stats = []
    for lev in range(m, g.nlevels()):
        t1 = time.perf_counter()
        n = C.c_int64(0)
        g._chk(g._L.ahfgpu_amr_patch_stats(g._h, lev, C.byref(n), None, 0))
        t2 = time.perf_counter()
        stats.append(g.patch_stats(lev, n.value))
        t3 = time.perf_counter()
        print("level", lev, "cells", int(g.level_header(lev)[0][1]), "patches", n.value, "compute ms %.2f fetch ms %.2f" % ((t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
    t4 = time.perf_counter()
    if os.environ.get("AHF_SAVE_STATS"):
        np.savez_compressed(os.environ["AHF_SAVE_STATS"], boxsize=box.boxsize, **{"lev%d" % i: s for i, s in enumerate(stats)})
    out = ahf.tree_halos(stats, 3.0 / box.boxsize)
    t5 = time.perf_counter()
    print("patch tables total ms %.1f, host tree ms %.1f, haloes %d" % ((t4 - t0) * 1e3, (t5 - t4) * 1e3, len(out["npart"])), flush=True)
    c, r, s = np.ascontiguousarray(out["pos"]), np.ascontiguousarray(out["gather_rad"]), np.ascontiguousarray(out["npart"], np.int64)
    for mode, env in (("default", {}), ("profiles one CTA per halo", {"AHFGPU_PROFILES_V1": "1"}), ("default + per-launch events", {"AHFGPU_KTIME": "1"})):
        os.environ.update(env)
        for it in range(3):
            g.synchronize(); t0 = time.perf_counter()
            g.construct_halos(c, r, s, fetch=False)
            g.synchronize(); t1 = time.perf_counter()
        for k in env:
            os.environ.pop(k)
        print(mode, "halo pass wall ms %.2f" % ((t1 - t0) * 1e3), {k: round(g.stage_ms(k), 3) for k in ("halo_gather", "halo_sort", "halo_localize", "halo_unbind", "halo_profiles")}, flush=True)
