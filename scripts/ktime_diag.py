"""Per-kernel times of ONE warm pass of the path on the bench workload: an event pair around every launch (AHFGPU_KTIME=1, the LAUNCH
macro of the library) -- kernel sums and the idle time between consecutive kernels, with warm caches (an ncu launch list is cold-cache
and serialised).  usage: python scripts/ktime_diag.py [n1d] [seeds: device|generator]   -> stderr of the library + a wall-clock line"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ahf_b200 import ahf, synth
n1d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
src = sys.argv[2] if len(sys.argv) > 2 else "device"
box = synth.make_box(n1d, seed=43)
par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d)
os.environ["AHFGPU_STAGES"] = "0"
with ahf.AhfGpu(par) as g:
    g.upload(box.pos, box.mom)
    g.sfc_sort_resident(); g.build_amr()
    if src == "device":
        hs = g.halo_seeds(3.0 / box.boxsize, lists=False)
        centres, rad, seednp = np.ascontiguousarray(hs["pos"]), np.ascontiguousarray(hs["gather_rad"]), np.ascontiguousarray(hs["npart"], np.int64)
    else:
        centres, rad, seednp = synth.halo_seeds(box)
    for it in range(4):
        if it == 3:
            os.environ["AHFGPU_KTIME"] = "1"
        w = []
        for f in (g.sfc_sort_resident, g.build_amr, lambda: g.construct_halos(centres, rad, seednp, fetch=False)):
            g.synchronize(); t0 = time.perf_counter(); f(); g.synchronize(); w.append((time.perf_counter() - t0) * 1e3)
        print("pass %d wall ms: sort %.3f mesh %.3f halo %.3f%s" % (it, w[0], w[1], w[2], "  (with per-launch events)" if it == 3 else ""), flush=True)
