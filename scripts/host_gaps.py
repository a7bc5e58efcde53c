import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
from ahf_b200 import ahf, synth
box = synth.make_box(256, seed=43)
c, r, npart = synth.halo_seeds(box)
par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=256)
with ahf.AhfGpu(par) as g:
    g.upload(box.pos, box.mom)
    for it in range(4):
        t0=time.perf_counter(); g.sfc_sort_resident(); g.synchronize(); t1=time.perf_counter()
        s1=sum(g.stage_ms(k) for k in ('keys','sort','gather'))
        g.build_amr(); g.synchronize(); t2=time.perf_counter()
        s2=sum(max(g.stage_ms(k),0) for k in ('ll','deposit','flag','refine','relink'))
        g.construct_halos(c, r, npart, fetch=False); g.synchronize(); t3=time.perf_counter()
        s3=sum(g.stage_ms(k) for k in ('halo_gather','halo_sort','halo_unbind','halo_profiles'))
        print('iter %d  sort wall %.2f ms (stages %.2f) | amr wall %.2f (stages %.2f) | halos wall %.2f (stages %.2f)'%(it,(t1-t0)*1e3,s1,(t2-t1)*1e3,s2,(t3-t2)*1e3,s3))
