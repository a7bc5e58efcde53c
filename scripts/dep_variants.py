"""Timing-only decomposition of the domain deposit kernel (library built with -DAHFGPU_EXPERIMENTS as ahf_b200/libahfgpu_exp.so):
AHFGPU_DOM_VARIANT 0 product, 1 red without return / carry, 2 no atomics, 3 no flush, 4 one tile copy.  Densities of 1..4 are wrong on purpose."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from ahf_b200 import ahf
ahf.LIB_PATH = os.path.join(os.path.dirname(ahf.LIB_PATH), "libahfgpu_exp.so")
from ahf_b200 import synth
n1d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
out = {}
for name, kw in (("bench box", {}), ("lattice only", dict(clump_frac=0.0001, n_clumps=1))):
    box = synth.make_box(n1d, seed=43, **kw)
    par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d, lgrid_max=n1d)
    with ahf.AhfGpu(par) as g:
        g.upload(box.pos, box.mom); g.sfc_sort_resident()
        res = {}
        for var in (0, 1, 2, 3, 4, 0):
            os.environ["AHFGPU_DOM_VARIANT"] = str(var)
            ts = []
            for _ in range(5):
                g.build_amr(); ts.append(g.stage_ms("deposit_dom_kernel"))
            res.setdefault(str(var), []).append(round(min(ts), 4))
        out[name] = res
os.environ.pop("AHFGPU_DOM_VARIANT", None)
print(json.dumps(out))
