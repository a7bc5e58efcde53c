"""Timing experiments on the domain deposit kernel (variants produce wrong densities on purpose; timing only).
Needs a library built with NVFLAGS+=-DAHFGPU_EXPERIMENTS (ahf_b200/csrc/Makefile); the shipped library refuses them."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ahf_b200 import ahf, synth
n1d = 256
for frac in (0.0001, 0.3):
    box = synth.make_box(n1d, seed=43, clump_frac=frac)
    par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d, lgrid_max=n1d)
    for var in sys.argv[1:] or ["0", "1", "2", "3", "4"]:
        os.environ["AHFGPU_DOM_VARIANT"] = var
        with ahf.AhfGpu(par) as g:
            g.sfc_sort(box.pos, box.mom)
            ts = []
            for _ in range(5):
                g.build_amr(); ts.append(round(g.stage_ms('deposit_dom_kernel'), 4))
            print('variant', var, 'clump_frac', frac, 'ms', ts, flush=True)
