"""GPU (one device is enough): ONE box split over 2 / 4 / 8 ranks -- host threads of this process sharing the GPU, the in-process
transport of ahf_b200/csrc/comm.cu -- gives exactly the single-GPU result: the ranks' own cells partition every level with bit-identical
densities, marks, run structure and particle counts; the final level of every particle, every halo scalar, member list (by global particle
index) and profile are identical.  The same checker runs over NCCL in tests/test_gpu_multi.py when the box has several GPUs."""
import numpy as np
import pytest

import slab_util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    from ahf_b200 import ahf
    return ahf


@pytest.fixture(scope="module")
def case(A):
    from ahf_b200 import synth
    n1d = 64
    # clumps on a rank boundary (box centre planes), on the periodic faces and in a corner, plus random ones
    cb = np.array([[0.5, 0.5, 0.5], [0.5, 0.25, 0.75], [0.999, 0.5, 0.3], [0.001, 0.002, 0.998], [0.25, 0.5, 0.5]])
    box = synth.make_box(n1d, seed=17, n_clumps=14, centres_box=cb)
    c, r, seed = synth.halo_seeds(box)
    T = slab_util.single_gpu_truth(A, box, n1d, c, r, seed)
    return dict(box=box, n1d=n1d, c=c, r=r, seed=seed, T=T)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_split_box_equals_single_gpu(A, case, world):
    from ahf_b200 import multigpu
    box, n1d = case["box"], case["n1d"]
    n = box.npart
    b = (np.arange(world + 1) * n) // world            # every rank "reads" a file-order slice
    par = A.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d)

    def fn(rank, sb):
        sb.distribute(box.pos[b[rank]:b[rank + 1]], box.mom[b[rank]:b[rank + 1]], id_base=int(b[rank]))
        return slab_util.rank_report(sb, case["c"], case["r"], case["seed"])

    reports = multigpu.run_local(world, par, fn)
    out = slab_util.check_against_truth(reports, case["T"], n)
    print(world, out)


def test_split_box_with_scrambled_input_and_uneven_shares(A, case):
    """the ranks read very different amounts, in random order: same result (the exchange is stable, equal keys keep their global index order)"""
    from ahf_b200 import multigpu
    box, n1d = case["box"], case["n1d"]
    n = box.npart
    world = 3
    b = np.array([0, n // 10, n // 2, n])
    par = A.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d)

    def fn(rank, sb):
        sb.distribute(box.pos[b[rank]:b[rank + 1]], box.mom[b[rank]:b[rank + 1]], id_base=int(b[rank]))
        return slab_util.rank_report(sb, case["c"], case["r"], case["seed"])

    reports = multigpu.run_local(world, par, fn)
    slab_util.check_against_truth(reports, case["T"], n)
