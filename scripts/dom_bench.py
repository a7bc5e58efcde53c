"""Domain-level TSC deposit alone (256^3 bench box and a lattice-only box): kernel time of k_deposit_dom2 / its heavy form / k_deposit_dom.
usage: python scripts/dom_bench.py [n1d] [reps]"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from ahf_b200 import ahf, synth
n1d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
only = sys.argv[3] if len(sys.argv) > 3 else None
out = {}
for boxname, kw in (("bench box", dict(seed=43)), ("lattice only", dict(seed=43, clump_frac=1e-9, n_clumps=1))):
    box = synth.make_box(n1d, **kw)
    par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d, lgrid_max=n1d)
    out[boxname] = {}
    for name, env in (("dom2", {"AHFGPU_DOM_V2": "1"}), ("dom2_heavy", {"AHFGPU_DOM_V2": "1", "AHFGPU_DOM2_HEAVY": "1"}), ("dom", {"AHFGPU_DOM_V2": "0"})):
        if only and name != only:
            continue
        for k in ("AHFGPU_DOM2_HEAVY", "AHFGPU_DOM_V2"):
            os.environ.pop(k, None)
        os.environ.update(env)
        os.environ["AHFGPU_DOM2_STATS"] = "1"
        with ahf.AhfGpu(par) as g:
            g.sfc_sort(box.pos, box.mom)
            g.build_amr()
            st = (g.stage_count("deposit_dom_ctas"), g.stage_count("dom2_heavy_ctas"), g.stage_count("dom2_failed_light"))
            os.environ.pop("AHFGPU_DOM2_STATS")
            ms = []
            for _ in range(reps):
                g.build_amr()
                ms.append(round(g.stage_ms("deposit_dom_kernel"), 4))
        out[boxname][name] = dict(ms=ms, ctas_upper=st[0], heavy=st[1], failed_light=st[2])
        print(boxname, name, out[boxname][name], flush=True)
print(json.dumps(out))
