"""Experiment: domain deposit kernel time with the tile's particles in Hilbert order (product) vs ordered by cell (AHFGPU_DOM_CELLSORT=1)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from ahf_b200 import ahf, synth
n1d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
box = synth.make_box(n1d, seed=43)
par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d, lgrid_max=n1d)
out = {}
with ahf.AhfGpu(par) as g:
    g.upload(box.pos, box.mom); g.sfc_sort_resident()
    ref = None
    for mode in ("hilbert", "cellsort", "hilbert", "cellsort"):
        if mode == "cellsort": os.environ["AHFGPU_DOM_CELLSORT"] = "1"
        else: os.environ.pop("AHFGPU_DOM_CELLSORT", None)
        ts = []
        for _ in range(6):
            g.build_amr(); ts.append(g.stage_ms("deposit_dom_kernel"))
        d = g.level(0, cells=False).dens
        if ref is None: ref = d.copy()
        out.setdefault(mode, []).append(dict(ms=ts, same=bool(np.array_equal(d, ref))))
print(json.dumps(out))
