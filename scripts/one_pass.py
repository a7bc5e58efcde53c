"""One warm pass + one measured pass of the whole path (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ahf_b200 import ahf, synth
n1d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
npass = int(sys.argv[2]) if len(sys.argv) > 2 else 2
box = synth.make_box(n1d, seed=43)
centres, rad, seednp = synth.halo_seeds(box)
par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d)
with ahf.AhfGpu(par) as g:
    g.upload(box.pos, box.mom)
    for it in range(npass):
        g.sfc_sort_resident(); g.build_amr(); g.construct_halos(centres, rad, seednp, fetch=False)
    print("launches", g.launches())
