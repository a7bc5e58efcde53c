"""Diagnostic: halo pass stage times on the bench box with the device-tree seeds, for several hybrid thresholds."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from ahf_b200 import ahf, synth
n1d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
box = synth.make_box(n1d, seed=43)
par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d)
with ahf.AhfGpu(par) as g:
    g.upload(box.pos, box.mom); g.sfc_sort_resident(); g.build_amr()
    hs = g.halo_seeds(3.0 / box.boxsize)
    c, r, s = np.ascontiguousarray(hs["pos"]), np.ascontiguousarray(hs["gather_rad"]), np.ascontiguousarray(hs["npart"], np.int64)
    for thr in ("0", "1024", "4096", "16384", "65536", "100000000"):
        os.environ["AHFGPU_UNBIND_SMALL"] = thr
        for it in range(3):
            g.construct_halos(c, r, s, fetch=False)
        st = {k: round(g.stage_ms(k), 3) for k in ("halo_gather", "halo_sort", "halo_localize", "halo_unbind", "halo_profiles")}
        print(thr, st, "iters", g.stage_count("halo_unbind_iterations"), "sweeps", g.stage_count("halo_unbind_mask_sweeps"), "small", g.stage_count("halo_unbind_small"), flush=True)
    S = g.fetch_halos(len(r), scal_only=True)["scal"]
    ng = S[:, 5]
    print("haloes", len(r), "gathered quantiles", np.quantile(ng, [0, .25, .5, .75, .9, .99, 1]).astype(int), "sum", int(ng.sum()))
    print("n(ng<=4096)", int((ng <= 4096).sum()), "n(ng>65536)", int((ng > 65536).sum()))
