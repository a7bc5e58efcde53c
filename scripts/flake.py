"""Flake hunt: repeated sort + build_amr (optionally + halo pass), signature = per-level (ncell, npart)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ahf_b200 import ahf, synth
n1d = int(sys.argv[1]); reps = int(sys.argv[2]); halos = int(sys.argv[3]); resort = int(sys.argv[4]) if len(sys.argv) > 4 else 1
box = synth.make_box(n1d, seed=43)
centres, rad, seednp = synth.halo_seeds(box)
par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d)
with ahf.AhfGpu(par) as g:
    g.upload(box.pos, box.mom)
    g.sfc_sort_resident()
    seen = {}
    for it in range(reps):
        if resort: g.sfc_sort_resident()
        g.build_amr()
        nl = g.nlevels()
        sig = tuple(tuple(int(x) for x in g.level_header(l)[0][:3]) for l in range(nl))
        if halos: g.construct_halos(centres, rad, seednp, fetch=False)
        seen.setdefault(sig, []).append(it)
    print("halos", halos, "resort", resort, "env", {k: v for k, v in os.environ.items() if k.startswith("AHFGPU")}, "distinct", len(seen), [(len(v), v[:6]) for v in seen.values()])
    if len(seen) > 1:
        sigs = list(seen.keys())
        for l in range(min(len(sigs[0]), len(sigs[1]))):
            if sigs[0][l] != sigs[1][l]: print("   first differing level", l, sigs[0][l], sigs[1][l]); break
