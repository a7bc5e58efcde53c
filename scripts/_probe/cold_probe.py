"""first-call vs second-call wall clock of the library's phases in a fresh process (cold costs of the drop-in program)"""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import numpy as np
from ahf_b200 import ahf, synth
n1d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
box = synth.make_box(n1d, seed=43)
par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d, device=0)
t = time.perf_counter()
g = ahf.AhfGpu(par)
print("init %.3f" % (time.perf_counter() - t))
for rep in range(3):
    t0 = time.perf_counter(); g.upload(box.pos, box.mom)
    t1 = time.perf_counter(); g.sfc_sort_resident()
    t2 = time.perf_counter(); g.build_amr()
    t3 = time.perf_counter()
    print("rep %d upload %.3f sort_resident %.3f build_amr %.3f" % (rep, t1 - t0, t2 - t1, t3 - t2), flush=True)
