import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from conftest import Golden
from ahf_b200 import ahf as A
g0=Golden('edge32')
par=A.params_from_reference(g0.glob, lgrid_dom=g0.n1d, nper_dom=g0.nper_dom, nper_ref=g0.nper_ref)
out={}
for mode in ('tiles','generic'):
    if mode=='generic': os.environ['AHFGPU_GENERIC_DEPOSIT']='1'
    with A.AhfGpu(par) as g:
        g.sfc_sort(g0.pos,g0.mom); g.build_amr(); out[mode]=g.level(0)
R=g0.level(0)
a=out['tiles'].dens.astype(np.float64); b=out['generic'].dens.astype(np.float64); r=R['dens'].astype(np.float64)
d=np.abs(a-b); i=np.argsort(-d)[:10]
print('tiles vs generic worst abs', d[i]); print('values', b[i]); print('ref', r[i]); print('cnt', R['cnt'][i])
print('x,y,z', R['x'][i],R['y'][i],R['z'][i])
print('sum tiles',(a+1).sum(),'generic',(b+1).sum(),'ref',(r+1).sum())
print('---')
for k in i[:4]:
    print('cell',R['x'][k],R['y'][k],R['z'][k],'tiles %.6f generic %.6f ref %.6f cnt %d'%(a[k],b[k],r[k],R['cnt'][k]))
e=np.abs(a-r)/np.maximum(np.abs(r),1); j=np.argsort(-e)[:5]
for k in j: print('worst-vs-ref cell',R['x'][k],R['y'][k],R['z'][k],'tiles %.6f generic %.6f ref %.6f cnt %d err %.2e'%(a[k],b[k],r[k],R['cnt'][k],e[k]))
