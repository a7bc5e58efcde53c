"""Per-level stage table of one pass of the path on the bench workload (CUDA-event stage timers of the library)."""
import os, sys, time
os.environ.setdefault("AHFGPU_LEVEL_STAGES", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ahf_b200 import ahf, synth
n1d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
box = synth.make_box(n1d, seed=43)
centres, rad, seednp = synth.halo_seeds(box)
par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d)
with ahf.AhfGpu(par) as g:
    g.upload(box.pos, box.mom)
    for it in range(3):
        g.sfc_sort_resident()
        g.synchronize(); t0 = time.perf_counter()
        g.build_amr()
        g.synchronize(); t1 = time.perf_counter()
        nl = g.nlevels()
        if it == 2:
            print("build_amr wall ms", (t1 - t0) * 1e3, "levels", nl)
            for k in ("ll", "deposit", "deposit_dom_kernel", "deposit_ref_kernel", "flag", "refine", "relink"):
                print("  %-22s %8.3f ms" % (k, g.stage_ms(k)))
            for l in range(nl):
                h, d = g.level_header(l)
                print("  L%d: %s" % (l, " ".join("%s %.3f" % (k, g.stage_ms("%s_L%d" % (k, l))) for k in ("deposit", "depk", "flag", "refine", "relink"))),
                      "ctas", g.stage_count("depk_L%d" % l), "hdr", list(h)[:6])
        g.synchronize(); t0 = time.perf_counter()
        g.construct_halos(centres, rad, seednp, fetch=False)
        g.synchronize(); t1 = time.perf_counter()
        if it == 2:
            print("construct_halos wall ms", (t1 - t0) * 1e3)
            for k in ("halo_gather", "halo_sort", "halo_unbind", "halo_profiles"):
                print("  %-22s %8.3f ms" % (k, g.stage_ms(k)))
    sc = g.fetch_halos(len(rad), scal_only=True)["scal"]
    npart = sc[:, 9]; ng = sc[:, 5]
    o = np.argsort(-ng)[:10]
    print("largest gathered:", ng[o].astype(int), "final npart:", npart[o].astype(int))
    print("sum gathered", int(ng.sum()), "halos", len(rad))
