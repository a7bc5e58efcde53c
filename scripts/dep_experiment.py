"""Domain TSC deposit kernel timing (CUDA events inside the library) on the bench workload; optional A/B against the
previous float-weight kernel (AHFGPU_DEPOSIT_V1=1).  Run on the GPU box: python scripts/dep_experiment.py [n1d]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ahf_b200 import ahf, synth
n1d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for frac in (0.0001, 0.3):
    box = synth.make_box(n1d, seed=43, clump_frac=frac)
    par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d)
    dens = {}
    for variant in ("v1", "v2"):
        if variant == "v1": os.environ["AHFGPU_DEPOSIT_V1"] = "1"
        else: os.environ.pop("AHFGPU_DEPOSIT_V1", None)
        with ahf.AhfGpu(par) as g:
            g.sfc_sort(box.pos, box.mom)
            ts = []
            for _ in range(5):
                g.build_amr(); ts.append(round(g.stage_ms('deposit_dom_kernel'), 4))
            dens[variant] = g.level(0).dens.astype(np.float64)
            alg = 16.0 * box.npart + 4.0 * n1d ** 3
            print(variant, 'clump_frac', frac, 'deposit_dom_kernel ms', ts, 'GB/s-alg', round(alg / min(ts[1:]) / 1e6, 1), 'ctas', g.stage_count('deposit_dom_ctas'), 'levels', g.nlevels(), flush=True)
    d = np.abs(dens["v1"] - dens["v2"]) / np.maximum(np.abs(dens["v1"]), 1.0)
    print('   v1 vs v2 dens: max rel', d.max(), 'sum v1', (dens["v1"] + 1).sum(), 'sum v2', (dens["v2"] + 1).sum(), 'N', box.npart)
