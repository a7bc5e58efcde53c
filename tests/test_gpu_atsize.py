"""GPU: parity AT THE SIZES THAT ARE BENCHMARKED.  The unmodified reference (oracle/_ref/ahf_ref, prebuilt, travels with the repo) runs
on the box itself on BASELINE.json's configs[0] (128^3, about 3 s) and configs[1] (256^3, the bench workload, about 25 s) and the CUDA
path is compared with its dumps stage by stage, with the north-star tolerances asserted explicitly (tests/parity_util.py): density
1e-5 per cell, halo count exact, M_vir / R_vir 1e-4, member overlap >= 99.9 % above 100 particles -- plus the exact cell-set /
run-structure / ownership checks and the profile columns (eigenvectors up to sign).  The measured deviations are written to
gpurun_out/parity_at_size_<n>.json (copied to profiles/ for the record)."""
import json
import os
import shutil
import tempfile

import pytest

import parity_util

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def A():
    from ahf_b200 import ahf
    return ahf


@pytest.mark.parametrize("n1d", [128, 256])
def test_benchmarked_size_against_reference_binary(A, n1d):
    from ahf_b200 import synth
    from oracle import oracle as O
    if not os.path.exists(O.REF_BIN):
        pytest.skip("oracle/_ref/ahf_ref not built")
    box = synth.make_box(n1d, seed=43)                      # bench.py's box (seed 43)
    work = tempfile.mkdtemp(prefix="ahf_atsize_%d_" % n1d)
    try:
        R = parity_util.run_reference_dump(box, work)
        out = parity_util.compare_with_reference(A, box, R, n1d)
        out.update(n1d=n1d, particles=box.npart, reference_timing=R["timing"])
        assert out["halos_ge_minpart"] >= 20 and out["halos_gt_100"] >= 10
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "parity_at_size_%d.json" % n1d), "w") as f:
                json.dump(out, f, indent=1)
        except OSError:
            pass
        print(out)
    finally:
        shutil.rmtree(work, ignore_errors=True)
