import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, zlib
from ahf_b200 import ahf, synth
n1d = 256
box = synth.make_box(n1d, seed=43, clump_frac=0.3)
par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d, lgrid_max=n1d)
for rmax in sys.argv[1:]:
    os.environ["AHFGPU_DOM_RMAX"] = rmax
    with ahf.AhfGpu(par) as g:
        g.sfc_sort(box.pos, box.mom)
        ts = []
        for _ in range(5):
            g.build_amr(); ts.append(round(g.stage_ms('deposit_dom_kernel'), 4))
        print('rmax', rmax, 'ms', ts, 'dens crc', zlib.crc32(g.level(0, cells=False).dens.tobytes()), flush=True)
