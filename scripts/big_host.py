"""Cooperative multi-block halo pass on one very large host (BASELINE.json configs[4] stand-in): stage timings, sanity checks,
optional A/B against the one-CTA-per-halo kernels.  python scripts/big_host.py [n_host] [ab]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ahf_b200 import ahf, synth
n_host = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
ab = len(sys.argv) > 2
t0 = time.time(); box = synth.make_host_box(n_host); print("generated", box.npart, "particles in %.1f s" % (time.time() - t0), flush=True)
c, r, npart = synth.halo_seeds(box)
par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=128)
res = {}
for variant in (("coop", "v1") if ab else ("coop",)):
    for k in ("AHFGPU_UNBIND_V1", "AHFGPU_PROFILES_V1", "AHFGPU_GATHER_V1"):
        if variant == "v1": os.environ[k] = "1"
        else: os.environ.pop(k, None)
    with ahf.AhfGpu(par) as g:
        keys, order = g.sfc_sort(box.pos, box.mom)
        for it in range(3):
            g.synchronize(); t0 = time.perf_counter()
            out = g.construct_halos(c, r, npart, fetch=(it == 2))
            g.synchronize(); wall = (time.perf_counter() - t0) * 1e3
        st = {k: round(g.stage_ms(k), 3) for k in ("halo_gather", "halo_sort", "halo_unbind", "halo_profiles")}
        S = out["scal"]
        print(variant, "halo pass wall ms (incl. fetch)", round(wall, 2), st, "gathered", int(S[:, 5].sum()), "host: gathered", int(S[0, 5]), "final npart", int(S[0, 9]),
              "unbind iterations", g.stage_count("halo_unbind_iterations"), "mask sweeps", g.stage_count("halo_unbind_mask_sweeps"), "unbind members/s", round(g.stage_count("halo_unbind_iter_members") / (st["halo_unbind"] * 1e-3) / 1e9, 3), "G/s", flush=True)
        res[variant] = (S.copy(), out["members"].copy(), out["prof"].copy())
S = res["coop"][0]
pos = box.pos[order]
m = ahf.AhfGpu.halo_members(None, dict(members=res["coop"][1], member_offset=out["member_offset"]), 0) if False else None
if ab:
    a, b = res["coop"], res["v1"]
    print("members identical:", np.array_equal(a[1], b[1]), " scal max rel diff:", np.nanmax(np.abs(a[0] - b[0]) / np.maximum(np.abs(b[0]), 1e-300)),
          " prof max rel diff:", np.nanmax(np.abs(a[2] - b[2]) / np.maximum(np.abs(b[2]), 1e-300)))
