"""A/B of the radix sort tile size (AHFGPU_RS_ITEMS=8|16, read once per process): keys/sort/gather and halo_sort stage times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ahf_b200 import ahf, synth
box = synth.make_box(256, seed=43)
c, r, npart = synth.halo_seeds(box)
par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=256)
with ahf.AhfGpu(par) as g:
    g.upload(box.pos, box.mom)
    for it in range(4):
        g.sfc_sort_resident()
        s = [g.stage_ms(k) for k in ("keys", "sort", "gather")]
        g.build_amr()
        g.construct_halos(c, r, npart, fetch=False)
        h = g.stage_ms("halo_sort")
    print("RS_ITEMS", os.environ.get("AHFGPU_RS_ITEMS", "default"), "keys %.3f sort %.3f gather %.3f halo_sort %.3f ms" % (*s, h))
