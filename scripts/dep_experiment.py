import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from ahf_b200 import ahf, synth
for frac in (0.0001, 0.3):
    box = synth.make_box(256, seed=43, clump_frac=frac)
    par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=256)
    with ahf.AhfGpu(par) as g:
        g.sfc_sort(box.pos, box.mom)
        ts=[]
        for _ in range(4):
            g.build_amr(); ts.append(g.stage_ms('deposit_dom_kernel'))
        print('clump_frac',frac,'deposit_dom_kernel ms',ts,'ctas',g.stage_count('deposit_dom_ctas'),'levels',g.nlevels())
