"""Run-to-run determinism of sort + hierarchy (+ halo pass): counts and checksums over repeated passes."""
import os, sys, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ahf_b200 import ahf, synth
n1d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
box = synth.make_box(n1d, seed=43)
centres, rad, seednp = synth.halo_seeds(box)
par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d)
with ahf.AhfGpu(par) as g:
    g.upload(box.pos, box.mom)
    seen = {}
    for it in range(reps):
        g.sfc_sort_resident()
        g.build_amr()
        nl = g.nlevels()
        hdr = tuple(tuple(int(x) for x in g.level_header(l)[0]) for l in range(nl))
        d0 = g.level(0).dens
        lv = [g.level(l) for l in range(nl)]
        per = tuple((zlib.crc32(L.x.tobytes() + L.y.tobytes() + L.z.tobytes()), zlib.crc32(L.dens.tobytes()), zlib.crc32(L.mark.tobytes()), zlib.crc32(L.count.tobytes())) for L in lv)
        sig = (hdr, zlib.crc32(d0.tobytes()), g.stage_count("deposit"), per)
        g.construct_halos(centres, rad, seednp, fetch=False)
        sc = g.fetch_halos(len(rad), scal_only=True)["scal"]
        sig = sig + (zlib.crc32(np.ascontiguousarray(sc).tobytes()),)
        if it == 0: ref = sig
        elif sig != ref:
            for l in range(min(len(sig[3]), len(ref[3]))):
                names = ("coords", "dens", "mark", "count")
                bad = [names[q] for q in range(4) if sig[3][l][q] != ref[3][l][q]]
                if bad:
                    print("iteration", it, "level", l, "differs in", bad)
                    if "dens" in bad and sig[3][l][0] == ref[3][l][0]:
                        a = lv[l].dens; b = ref_lv[l].dens; w = np.nonzero(a != b)[0]
                        print("   cells differing", len(w), "first", w[:8], "vals", a[w[:8]], b[w[:8]], "x", lv[l].x[w[:8]], "y", lv[l].y[w[:8]], "z", lv[l].z[w[:8]], "cnt", lv[l].count[w[:8]])
        if it == 0: ref_lv = lv
        seen.setdefault(sig, []).append(it)
    print("distinct outcomes:", len(seen))
    for s, its in seen.items():
        print(" iterations", its, "deposit", s[2], "dens crc", s[1], "scal crc", s[4], "ncell", [h[1] for h in s[0]], "npart", [h[2] for h in s[0]])
