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


@pytest.mark.parametrize("world", [2, 4, 8])
def test_split_box_derives_the_same_halo_seeds(A, case, world):
    """NEXT-1/2 over several ranks: patch labels as connected components across the rank boundaries (local union-find over own-sourced
    edges, names joined through the shared ghost cells, numbering by first cell), per-refinement tables combined over the ranks, the tree
    on every rank.  Every rank must end with the single-GPU tables -- counts, centres from the exact integer sums, maximum density and
    extents bit for bit, the density-weighted centre (double sums in a different order) to rounding -- and with the same halo seeds."""
    from ahf_b200 import multigpu
    box, n1d = case["box"], case["n1d"]
    n = box.npart
    par = A.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d)
    with A.AhfGpu(par) as g:
        g.sfc_sort(box.pos, box.mom); g.build_amr()
        ref = g.halo_seeds(3.0 / box.boxsize)
    assert len(ref["npart"]) >= 10 and sum(len(s) for s in ref["stats"]) >= 20
    b = (np.arange(world + 1) * n) // world

    def fn(rank, sb):
        sb.distribute(box.pos[b[rank]:b[rank + 1]], box.mom[b[rank]:b[rank + 1]], id_base=int(b[rank]))
        sb.build_amr()
        return sb.g.halo_seeds(3.0 / box.boxsize)

    outs = multigpu.run_local(world, par, fn)
    exact = [0, 1, 2, 3, 4, 5, 6, 7, 8, 12, 13, 14, 15, 16, 17]
    for o in outs:
        assert o["min_ref"] == ref["min_ref"] and len(o["stats"]) == len(ref["stats"])
        for a, r in zip(o["stats"], ref["stats"]):
            assert a.shape == r.shape
            assert np.array_equal(a[:, exact], r[:, exact]), np.nonzero(a[:, exact] != r[:, exact])
            assert np.allclose(a[:, 9:12], r[:, 9:12], rtol=0, atol=1e-12)
        for k in ("pos", "npart", "host", "host_level"):
            assert np.array_equal(o[k], ref[k]), k
        # sub-haloes gather at least out to closeRefDist, a distance between density-weighted centres: to rounding
        assert np.allclose(o["gather_rad"], ref["gather_rad"], rtol=1e-9, atol=0)
        plain = ref["host"] < 0
        assert np.array_equal(o["gather_rad"][plain], ref["gather_rad"][plain])


def test_split_box_writes_the_single_gpu_catalogue(A, case, tmp_path):
    """BASELINE.json configs[3] in small: particles -> decomposition -> mesh -> seeds (labels across rank boundaries) -> halo pass on the
    owning ranks -> re-hash, ordering and the four catalogue files; byte-identical to the files of the same box on one GPU."""
    from ahf_b200 import multigpu
    box, n1d = case["box"], case["n1d"]
    n = box.npart
    world = 4
    maxg = 3.0 / box.boxsize
    par = A.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d)
    with A.AhfGpu(par) as g:
        keys, order = g.sfc_sort(box.pos, box.mom); g.build_amr()
        seeds = g.halo_seeds(maxg)
        res = g.construct_halos(np.ascontiguousarray(seeds["pos"]), np.ascontiguousarray(seeds["gather_rad"]), np.ascontiguousarray(seeds["npart"], np.int64))
        res = dict(res); res["members"] = order.astype(np.int64)[res["members"]]          # sorted offsets -> input indices
        multigpu.catalogue_from_ranks(str(tmp_path / "one.z0.000"), par, seeds, [(np.arange(len(seeds["npart"])), res)], box.ids)
    b = (np.arange(world + 1) * n) // world

    def fn(rank, sb):
        sb.distribute(box.pos[b[rank]:b[rank + 1]], box.mom[b[rank]:b[rank + 1]], id_base=int(b[rank]), ghost_width=min(maxg, 0.25))
        sb.build_amr()
        sd = sb.g.halo_seeds(maxg)
        mine, r = sb.construct_halos(np.ascontiguousarray(sd["pos"]), np.ascontiguousarray(sd["gather_rad"]), np.ascontiguousarray(sd["npart"], np.int64))
        return sd, mine, r

    outs = multigpu.run_local(world, par, fn)
    multigpu.catalogue_from_ranks(str(tmp_path / "split.z0.000"), par, outs[0][0], [(m, r) for _, m, r in outs], box.ids)
    nh = 0
    for ext in ("AHF_halos", "AHF_profiles", "AHF_substructure", "AHF_particles"):
        a = open(tmp_path / ("one.z0.000." + ext), "rb").read(); bb = open(tmp_path / ("split.z0.000." + ext), "rb").read()
        assert a == bb, ext
        if ext == "AHF_halos":
            nh = a.count(b"\n") - 1
    assert nh >= 5


@pytest.mark.parametrize("world", [2, 4])
def test_split_multispecies_box_equals_single_gpu(A, world):
    """gas + dark matter + star particles (weights and thermal energies travel through the exchange in pos4.w / mom4.w): the mass-weighted
    deposit, the hierarchy and the halo pass of a split box equal the single-GPU run bit for bit"""
    from ahf_b200 import multigpu, synth
    n1d = 64
    cb = np.array([[0.5, 0.5, 0.5], [0.999, 0.5, 0.3], [0.25, 0.5, 0.5]])
    sbx = synth.make_species_box(n1d, seed=23, n_clumps=10, centres_box=cb)
    box = sbx.box
    n = box.npart
    c, r, seed = synth.halo_seeds(box)
    T = slab_util.single_gpu_truth(A, box, n1d, c, r, seed, weight=sbx.weight, u=sbx.u)
    b = (np.arange(world + 1) * n) // world
    par = A.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d)

    def fn(rank, sb):
        sl = slice(b[rank], b[rank + 1])
        sb.distribute(box.pos[sl], box.mom[sl], sbx.weight[sl], sbx.u[sl], id_base=int(b[rank]))
        return slab_util.rank_report(sb, c, r, seed)

    reports = multigpu.run_local(world, par, fn)
    slab_util.check_against_truth(reports, T, n)
