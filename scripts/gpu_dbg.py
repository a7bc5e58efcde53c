import sys, traceback
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from conftest import Golden, lin
from ahf_b200 import ahf as A
for name in ['frag16','edge32','plain32']:
    g0=Golden(name)
    par=A.params_from_reference(g0.glob, lgrid_dom=g0.n1d, nper_dom=g0.nper_dom, nper_ref=g0.nper_ref)
    with A.AhfGpu(par) as g:
        k=g.hilbert_keys(g0.pos); print(name,'keys eq',np.array_equal(k,g0.keys))
        pin,min_=g0.input_order()
        keys,order=g.sfc_sort(pin,min_); print(' sort keys eq',np.array_equal(keys,g0.keys),'perm ok',np.array_equal(np.sort(order),np.arange(len(order))))
        g.sfc_sort(g0.pos,g0.mom)
        try:
            nl=g.build_amr(); print(' nlev',nl,'ref',g0.nlev)
            owner,cells=g.particle_levels()
            for l in range(min(nl,g0.nlev)):
                G=g.level(l); R=g0.level(l)
                same=G.ncell==len(R['x']) and np.array_equal(G.lin(),lin(R['x'],R['y'],R['z'],R['l1dim']))
                msg='  L%d ncell %d ref %d cells %s'%(l,G.ncell,len(R['x']),same)
                if same:
                    err=np.abs(G.dens.astype(np.float64)-R['dens'])/np.maximum(np.abs(R['dens']),1)
                    msg+=' rf %s cnt %s denserr %.2e'%(np.array_equal(G.runflags,R['runflags']),np.array_equal(G.count,R['cnt']),err.max())
                    fin=np.zeros(len(g0.keys),bool); fin[R['plist_final']]=True
                    msg+=' owner %s'%np.array_equal(owner==l,fin)
                print(msg)
        except Exception as e:
            traceback.print_exc()
        try:
            res=g.construct_halos(g0.hs[:,0:3].copy(), g0.hs[:,3].copy(), g0.hs[:,4].astype(np.int64))
            S=res['scal']
            for i in range(len(S)):
                ref=g0.hs[i]
                print('  halo',i,'stages gpu',S[i,5:10],'ref',ref[5:10],'members eq',np.array_equal(g.halo_members(res,i),g0.members(i)))
                if ref[9]>=20:
                    sl=list(range(10,58)); e=np.abs(S[i,sl]-ref[sl])/np.maximum(np.abs(ref[sl]),1e-300); e[S[i,sl]==ref[sl]]=0
                    bad=[(sl[k],ref[sl[k]],S[i,sl[k]]) for k in np.nonzero(e>1e-9)[0]]
                    print('     worst rel',e.max(),'bad',bad[:8])
                    pr=g0.prof(i); pg=g.halo_profile(res,i)
                    if pg is not None and pg.shape==pr.shape:
                        ep=np.abs(pr-pg)/np.maximum(np.abs(pr),1e-300); ep[pr==pg]=0
                        print('     prof worst',ep.max(), np.unravel_index(np.argmax(ep),ep.shape))
                    else: print('     prof shape mismatch', None if pg is None else pg.shape, pr.shape)
        except Exception as e:
            traceback.print_exc()
