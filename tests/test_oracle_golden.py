"""CPU: the C restatement (oracle/) against the golden vectors produced by the unmodified reference.
Integer / index results must be identical; the oracle reproduces the reference's summation order, so
density and halo scalars are expected bit for bit (asserted to 1e-12 to stay robust to libm builds)."""
import numpy as np

from conftest import lin
from oracle import oracle as O


def test_keys_match_reference(golden):
    k = O.hilbert_keys(golden.pos)
    assert np.array_equal(k, golden.keys)
    assert np.all(np.diff(golden.keys.astype(np.int64)) >= 0)
    pos_in, _ = golden.input_order()
    order = O.argsort_keys(O.hilbert_keys(pos_in))
    assert np.array_equal(O.hilbert_keys(pos_in)[order], golden.keys)


def test_key_prefix_property(golden):
    # the halo gather relies on it (ahf_halos_sfc.c:370-412): the key at b bits is a prefix of the 21-bit key
    for bits in (1, 2, 5, 11, 20):
        kb = O.hilbert_keys(golden.pos[:2000], bits)
        assert np.array_equal(kb, golden.keys[:2000] >> np.uint64(3 * (21 - bits)))


def test_hierarchy_matches_reference(golden):
    H = O.build_hierarchy(golden.pos, golden.n1d, nth_dom=golden.nper_dom, nth_ref=golden.nper_ref)
    assert len(H) == golden.nlev
    for l, M in enumerate(H):
        R = golden.level(l)
        assert M.l1dim == int(R["l1dim"]) and M.ncell == len(R["x"])
        assert np.array_equal(M.lin(), lin(R["x"], R["y"], R["z"], R["l1dim"]))
        assert np.array_equal(M.runflags, R["runflags"])
        assert np.array_equal(M.cnt_flag, R["cnt"]) and np.array_equal(M.plist_flag, R["plist"])
        assert np.array_equal(M.dens, R["dens"]), "density differs from the reference bit pattern"
        assert abs(M.critdens - float(R["critdens"])) <= 1e-12 * abs(M.critdens)
        assert np.array_equal(M.cnt_final, R["cnt_final"]) and np.array_equal(M.plist_final, R["plist_final"])


def test_halo_pass_matches_reference(golden):
    par = O.params_from_glob(golden.glob)
    res = O.construct_halos(golden.keys, golden.pos, golden.mom, golden.weight, golden.u, par, golden.hs[:, 0:3].copy(),
                            golden.hs[:, 3].copy(), golden.hs[:, 4].copy())
    slots = list(range(10, 58))
    for i, r in enumerate(res):
        ref = golden.hs[i]
        if ref[4] == 0:
            continue
        assert [r["n_gather"], r["n_rvir0"], r["n_unbound"], r["n_rvir1"], r["npart"]] == [int(v) for v in ref[5:10]]
        assert np.array_equal(r["ipart"], golden.members(i))
        if r["npart"] < par["min_part"]:
            continue
        a, b = ref[slots], r["s"][slots]
        assert np.allclose(a, b, rtol=1e-12, atol=0), (i, np.nonzero(~np.isclose(a, b, rtol=1e-12, atol=0)))
        assert np.allclose(golden.prof(i), r["prof"], rtol=1e-12, atol=0)
        if golden.species(i) is not None:          # GAS_PARTICLES build: gas_only / stars_only and the species profile columns
            assert golden.species(i)[0] + golden.species(i)[32] > 0
            assert np.allclose(golden.species(i), r["species"], rtol=1e-12, atol=0), (i, np.nonzero(~np.isclose(golden.species(i), r["species"], rtol=1e-12, atol=0)))
            assert np.allclose(golden.prof_species(i), r["prof_species"], rtol=1e-12, atol=0)


def test_patch_colouring_matches_reference(golden):
    """NEXT-1: the literal restatement of the colouring sweep of ahf_gridinfo (ahf_gridinfo.c:236-577) numbers the isolated
    refinements of every coloured level exactly as the reference does, periodic flags included."""
    min_ref, ref = golden.patches()
    H = O.build_hierarchy(golden.pos, golden.n1d, nth_dom=golden.nper_dom, nth_ref=golden.nper_ref, patches=True)
    assert min_ref >= 1 and max(ref) == len(H) - 1
    for l, (iso, per) in ref.items():
        assert np.array_equal(H[l].iso, iso), l
        assert np.array_equal(H[l].iso_periodic, per), l
        assert iso.min() == 0 and iso.max() == per.shape[0] - 1


def test_patch_centres_match_reference_gridtree(golden):
    """NEXT-2, first half: RefCentre (ahf_halos.c:935-1390).  Node and particle counts of every isolated refinement are those of the
    reference's own .AHF_gridtree, its centres (centre of mass of the linked particles, AHFcomcentre in the shipped define.h) agree
    to the 14 decimals the file prints."""
    T = golden.gridtree()
    if T is None:
        import pytest
        pytest.skip("no -DAHFgridtreefile variant of the multi-species build")
    H = O.build_hierarchy(golden.pos, golden.n1d, nth_dom=golden.nper_dom, nth_ref=golden.nper_ref, patches=True)
    for l, t in T.items():
        pc = H[l].patch
        assert np.array_equal(pc[:, 0].astype(np.int64), t["nodes"]) and np.array_equal(pc[:, 1].astype(np.int64), t["parts"]), l
        d = np.abs(pc[:, 2:5] - t["centre"])
        assert np.minimum(d, 1.0 - d).max() <= 6e-15, (l, d.max())
        # a leaf refinement holds every particle that deposits on it, and TSC preserves the first moment: density-weighted == particle centre
        leaf = t["daughter"][:, 1] < 0
        if leaf.any():
            dd = np.abs(pc[leaf, 9:12] - pc[leaf, 2:5])
            assert np.minimum(dd, 1.0 - dd).max() < 1e-6


def test_refinement_tree_matches_reference_gridtree(golden):
    """NEXT-2, second half: extents of the isolated refinements and analyseRef (ahf_halos.c:1400-1620, :1652-2300).  Substructure lists
    (members and order), main-branch daughter and closeRefDist of every isolated refinement equal the reference's own .AHF_gridtree."""
    T = golden.gridtree()
    if T is None:
        import pytest
        pytest.skip("no -DAHFgridtreefile variant of the multi-species build")
    H = O.build_hierarchy(golden.pos, golden.n1d, nth_dom=golden.nper_dom, nth_ref=golden.nper_ref, patches=True)
    tree = O.patch_tree(H[min(T):])
    for i, l in enumerate(sorted(T)):
        t, o = T[l], tree[i]
        subs = [[int(v) for v in t["sub"][t["sub_off"][j]:t["sub_off"][j + 1], 1]] for j in range(len(t["nodes"]))]
        assert o["sub"] == subs, l
        assert np.array_equal(o["daughter"], t["daughter"][:, 1]), l
        assert np.abs(o["close"] - t["close"]).max() <= 1e-13, l
        if len(t["sub"]):
            assert np.all(t["sub"][:, 0] == l + 1)


def test_halo_seeds_match_reference(golden):
    """NEXT-2/3 glue: spatialRef2halos (ahf_halos.c:2405-3058) on top of the restated colouring, RefCentre and analyseRef gives the
    halo seeds -- centre, gathering radius, particle count, in the order of the reference's halos[] array -- that the reference hands
    to ahf_halos_sfc_constructHalo (golden `halo_s` columns 0-4, dumped at ahf_halos.c:508): bit for bit.  With this the oracle
    restates the whole chain particles -> keys -> hierarchy -> patches -> tree -> seeds -> halo pass."""
    min_ref, _ = golden.patches()
    H = O.build_hierarchy(golden.pos, golden.n1d, nth_dom=golden.nper_dom, nth_ref=golden.nper_ref, patches=True)
    tree = O.patch_tree(H[min_ref:])
    pos, gather, npart, host = O.tree_to_halos(H[min_ref:], tree, 3.0 / float(golden.d["boxsize"]))      # MaxGatherRad 3.0 (AHF.input-example)
    hs = golden.hs
    assert len(npart) == len(hs)
    assert np.array_equal(npart, hs[:, 4].astype(np.int64))
    assert np.array_equal(pos, hs[:, 0:3]) and np.array_equal(gather, hs[:, 3])
