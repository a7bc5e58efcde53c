"""GPU: libahfgpu.so (through the C-ABI, ahf_b200/ahf.py) against the golden vectors of the reference and
against the CPU oracle.  Tolerances (BASELINE.json north_star): keys / cells / flags / counts / halo
membership counts exact; density 1e-5 relative per cell (measured against max(|dens|,1): dens is a
contrast n/n_mean - 1 and crosses zero); halo masses and radii 1e-4 relative (we assert 1e-9)."""
import numpy as np
import pytest

from conftest import lin

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    from ahf_b200 import ahf
    return ahf


def _ctx(A, golden, **kw):
    par = A.params_from_reference(golden.glob, lgrid_dom=golden.n1d, nper_dom=golden.nper_dom, nper_ref=golden.nper_ref, **kw)
    return A.AhfGpu(par)


def test_hilbert_keys_bit_exact(A, golden):
    with _ctx(A, golden) as g:
        assert np.array_equal(g.hilbert_keys(golden.pos), golden.keys)
        for bits in (1, 2, 7, 13):
            assert np.array_equal(g.hilbert_keys(golden.pos[:5000], bits), golden.keys[:5000] >> np.uint64(3 * (21 - bits)))


def test_sort_matches_reference_order(A, golden):
    pos_in, mom_in = golden.input_order()
    with _ctx(A, golden) as g:
        keys, order = g.sfc_sort(pos_in, mom_in)
        assert np.array_equal(keys, golden.keys)
        # same multiset per key; where keys are unique the permutation is the reference's
        assert np.array_equal(np.sort(order), np.arange(len(order), dtype=np.uint32))
        uniq = np.concatenate([[True], keys[1:] != keys[:-1]]) & np.concatenate([keys[1:] != keys[:-1], [True]])
        assert np.array_equal(order[uniq], golden.ids[uniq])
        assert np.array_equal(pos_in[order][uniq], golden.pos[uniq])


def test_sort_aos_dropin(A, golden):
    """the reference's 48-byte struct particle (tdef.h:36-77): ll, pos[3], mom[3], pad, sfckey, id"""
    pos_in, mom_in = golden.input_order()
    dt = np.dtype([("ll", "<u8"), ("pos", "<f4", 3), ("mom", "<f4", 3), ("sfckey", "<u8"), ("id", "<u8")], align=True)
    assert dt.itemsize == 48
    part = np.zeros(len(pos_in), dt)
    part["pos"] = pos_in; part["mom"] = mom_in; part["id"] = np.arange(len(pos_in)); part["ll"] = 0xdeadbeef
    with _ctx(A, golden) as g:
        g.sfc_sort_particles(part, dt.fields["pos"][1], dt.fields["mom"][1], dt.fields["sfckey"][1], dt.fields["id"][1])
    assert np.array_equal(part["sfckey"], golden.keys)
    assert np.all(part["ll"] == 0)
    assert np.array_equal(part["pos"], pos_in[part["id"]]) and np.array_equal(part["mom"], mom_in[part["id"]])
    assert np.array_equal(np.sort(part["id"]), np.arange(len(pos_in), dtype=np.uint64))


def _check_levels(g, golden, order=None):
    """order: input id of our sorted particle i (None: our sorted order IS the golden one).  With it, everything that refers to a
    sorted offset is compared by particle id, so that equal Hilbert keys (free tie order) do not matter."""
    nl = g.build_amr()
    assert nl == golden.nlev, (nl, golden.nlev)
    owner, cells = g.particle_levels()
    n = len(golden.keys)
    mine = np.arange(n) if order is None else np.asarray(order, np.int64)
    theirs = np.arange(n) if order is None else golden.ids.astype(np.int64)
    worst = 0.0
    for l in range(nl):
        G = g.level(l); R = golden.level(l)
        assert G.l1dim == int(R["l1dim"]) and G.ncell == len(R["x"]), (l, G.ncell, len(R["x"]))
        assert np.array_equal(G.lin(), lin(R["x"], R["y"], R["z"], R["l1dim"])), "cell set differs on level %d" % l
        assert np.array_equal(G.runflags, R["runflags"]), "run structure differs on level %d" % l
        assert np.array_equal(G.count, R["cnt"]), "particles per node differ on level %d" % l
        err = np.abs(G.dens.astype(np.float64) - R["dens"]) / np.maximum(np.abs(R["dens"]), 1.0)
        worst = max(worst, err.max())
        assert err.max() <= 1e-5, (l, err.max())
        assert abs(G.critdens - float(R["critdens"])) <= 1e-12 * G.critdens
        # which particles sit on which node: as sets per node (list order is a linked-list artefact)
        cell_ref = np.full(n, -1, np.int64)
        cell_ref[theirs[R["plist"]]] = np.repeat(np.arange(G.ncell), R["cnt"])
        cell_gpu = np.empty(n, np.int64); cell_gpu[mine] = cells[l]
        assert np.array_equal(cell_gpu, cell_ref), "particle -> node map differs on level %d" % l
        fin_ref = np.zeros(n, bool); fin_ref[theirs[R["plist_final"]]] = True
        own_gpu = np.empty(n, np.int64); own_gpu[mine] = owner
        assert np.array_equal(own_gpu == l, fin_ref), "final ownership differs on level %d" % l
    return worst


def test_amr_matches_reference(A, golden):
    with _ctx(A, golden) as g:
        g.sfc_sort(golden.pos, golden.mom, golden.weight, golden.u)       # already key-sorted: the stable sort keeps the order
        worst = _check_levels(g, golden)
        print("max density error", worst)


def test_patch_labels_match_reference(A, golden):
    """NEXT-1: the patch colouring of ahf_gridinfo (src/libahf/ahf_gridinfo.c:236-577) on the device -- union-find over
    the neighbour table (ahfgpu_amr_patches) -- numbers the isolated refinements of every level as the unmodified reference does
    (tests/golden/patches.npz, its coloured levels) and as the CPU restatement of the sweep does (every refinement level)."""
    from oracle import oracle as O
    min_ref, ref = golden.patches()
    H = O.build_hierarchy(golden.pos, golden.n1d, nth_dom=golden.nper_dom, nth_ref=golden.nper_ref, patches=True)
    with _ctx(A, golden) as g:
        g.sfc_sort(golden.pos, golden.mom, golden.weight, golden.u)
        nl = g.build_amr()
        assert nl == len(H)
        for l in range(1, nl):
            iso, per = g.patches(l)
            assert np.array_equal(iso, H[l].iso), "patch labels differ from the restated sweep on level %d" % l
            assert np.array_equal(per, H[l].iso_periodic), l
            if l in ref:
                assert np.array_equal(iso, ref[l][0]), "patch labels differ from the reference on level %d" % l
                assert np.array_equal(per, ref[l][1]), l
        iso0, per0 = g.patches(0)                     # the periodic domain grid is one patch, periodic in all three directions
        assert iso0.min() == 0 and iso0.max() == 0 and per0.tolist() == [[1, 1, 1]]


def _check_halos(g, golden, rtol=1e-9, order=None):
    from parity_util import eigvec_cols_match
    n = len(golden.keys)
    mine = np.arange(n) if order is None else np.asarray(order, np.int64)
    theirs = np.arange(n) if order is None else golden.ids.astype(np.int64)
    res = g.construct_halos(golden.hs[:, 0:3].copy(), golden.hs[:, 3].copy(), golden.hs[:, 4].astype(np.int64))
    S = res["scal"]
    minpart = int(golden.glob[9])
    slots = [10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37,
             38, 39, 40, 41, 42, 43, 53, 54, 55, 56, 57]
    for i in range(len(S)):
        ref = golden.hs[i]
        if ref[4] == 0:
            assert S[i, 9] == 0
            continue
        assert np.array_equal(S[i, 5:10], ref[5:10]), (i, S[i, 5:10], ref[5:10])
        m = mine[g.halo_members(res, i)]
        mr = theirs[golden.members(i)]
        if order is None:
            assert np.array_equal(m, mr), "member list of halo %d" % i
        else:       # equal keys and equal radii go together (identical positions): same members, order free inside such ties
            assert np.array_equal(np.sort(m), np.sort(mr)), "member set of halo %d" % i
        if ref[9] < minpart:
            continue
        a, b = ref[slots], S[i, slots]
        ok = np.isclose(a, b, rtol=rtol, atol=1e-300)
        assert ok.all(), (i, [(slots[k], a[k], b[k]) for k in np.nonzero(~ok)[0]])
        # eigenvectors are defined up to sign
        for k0 in (44, 47, 50):
            va, vb = ref[k0:k0 + 3], S[i, k0:k0 + 3]
            assert np.allclose(va, vb, rtol=1e-6, atol=1e-9) or np.allclose(va, -vb, rtol=1e-6, atol=1e-9), (i, k0, va, vb)
        pr, pg = golden.prof(i), g.halo_profile(res, i)
        assert pg is not None and pg.shape == pr.shape
        cols = [c for c in range(25) if c not in (14, 15, 16, 18, 19, 20, 22, 23, 24)]
        okp = np.isclose(pr[cols], pg[cols], rtol=1e-8, atol=1e-300)
        assert okp.all(), (i, np.argwhere(~okp)[:5])
        bad = eigvec_cols_match(pr, pg)             # columns 14-16 / 18-20 / 22-24: eigenvectors per bin, up to sign
        assert not bad, (i, bad[:3])
        if golden.species(i) is not None:       # GAS_PARTICLES build: gas_only / stars_only blocks, M_gas / M_star / u_gas columns
            sr, sg = golden.species(i), res["species"][i]
            sl = [q + 32 * t for t in (0, 1) for q in list(range(0, 19)) + [28, 29]]
            oks = np.isclose(sr[sl], sg[sl], rtol=rtol, atol=1e-300)
            assert oks.all(), (i, [(sl[k], sr[sl[k]], sg[sl[k]]) for k in np.nonzero(~oks)[0]])
            for t in (0, 1):
                for k0 in (19, 22, 25):        # eigenvectors up to sign
                    va, vb = sr[32 * t + k0:32 * t + k0 + 3], sg[32 * t + k0:32 * t + k0 + 3]
                    assert np.allclose(va, vb, rtol=1e-6, atol=1e-9) or np.allclose(va, -vb, rtol=1e-6, atol=1e-9), (i, t, k0, va, vb)
            assert np.allclose(golden.prof_species(i), g.halo_profile_species(res, i), rtol=1e-8, atol=1e-300)


@pytest.mark.parametrize("unbind", ["hybrid", "cooperative", "one_cta", "sort_ties8", "sort_ties0"])
def test_halo_pass_matches_reference(A, golden, unbind):
    """species32: the -DMULTIMASS -DGAS_PARTICLES build (weights in M_vir / potential / profiles, thermal energy in the bound test
    and Ekin, R_max and r2 from the dark matter alone).  unbind: the default hybrid (haloes up to 16384 gathered members by one CTA each,
    larger ones by the cooperative multi-block pass), everything cooperative, everything by one CTA per halo -- all three must give the
    reference's member lists and scalars.  sort_ties*: the radial sort keeps only 8 / 0 mantissa bits of r^2 in its keys, so that
    k_fix_ties has to order long runs of equal keys (insertion and heap sort paths) -- the member ORDER must still be the reference's."""
    import os
    env = {"hybrid": {}, "cooperative": {"AHFGPU_UNBIND_SMALL": "0"}, "one_cta": {"AHFGPU_UNBIND_V1": "1"},
           "sort_ties8": {"AHFGPU_HALO_SORT_SKIP": "44"}, "sort_ties0": {"AHFGPU_HALO_SORT_SKIP": "52"}}[unbind]
    os.environ.update(env)
    try:
        with _ctx(A, golden) as g:
            g.sfc_sort(golden.pos, golden.mom, golden.weight, golden.u)
            _check_halos(g, golden)
    finally:
        for k in env:
            os.environ.pop(k, None)


def test_end_to_end_from_file_order(A, golden):
    """sort + mesh + halo pass from the snapshot's own particle order (what main.c hands over)"""
    pos_in, mom_in = golden.input_order()
    w_in = u_in = None
    if golden.weight is not None:
        w_in = np.empty_like(golden.weight); u_in = np.empty_like(golden.u)
        w_in[golden.ids] = golden.weight; u_in[golden.ids] = golden.u
    with _ctx(A, golden) as g:
        keys, order = g.sfc_sort(pos_in, mom_in, w_in, u_in)
        assert np.array_equal(keys, golden.keys)
        if np.all(keys[1:] != keys[:-1]):
            assert np.array_equal(order.astype(np.uint64), golden.ids)
            _check_levels(g, golden)
            _check_halos(g, golden)
        else:   # duplicate keys: tie order may differ from libc qsort -> compare by particle id
            _check_levels(g, golden, order=order)
            _check_halos(g, golden, order=order)


def test_overlapped_upload_gives_the_same_results(A, golden):
    """ahfgpu_sfc_sort_soa_async (momenta copied behind the sort and the hierarchy build, pinned host buffers) against the
    golden vectors, twice on one context so that the second call recycles buffers the copy stream used."""
    import torch
    hp = torch.from_numpy(np.ascontiguousarray(golden.pos, np.float32)).pin_memory()
    hm = torch.from_numpy(np.ascontiguousarray(golden.mom, np.float32)).pin_memory()
    hw = hu = None
    if golden.weight is not None:
        hw = torch.from_numpy(np.ascontiguousarray(golden.weight, np.float32)).pin_memory()
        hu = torch.from_numpy(np.ascontiguousarray(golden.u, np.float32)).pin_memory()
    ho = torch.full((hp.shape[0],), -1, dtype=torch.int32).pin_memory()
    with _ctx(A, golden) as g:
        for _ in range(2):
            g.sfc_sort_async_ptr(hp.data_ptr(), hm.data_ptr(), hp.shape[0], hw.data_ptr() if hw is not None else 0,
                                 hu.data_ptr() if hu is not None else 0)
            ho.fill_(-1)
            g._chk(g._L.ahfgpu_particle_ids_async(g._h, ho.data_ptr()))       # permutation: device -> host behind the momenta
            _check_levels(g, golden)
            _check_halos(g, golden)                                            # construct_halos has returned: the copy is complete
            assert np.array_equal(ho.numpy().view(np.uint32), g.particle_ids())


def test_member_lists_sent_early_equal_the_fetched_ones(A, golden):
    """ahfgpu_halo_members_buffer: with a registered pinned buffer the member lists leave for the host while the profiles are computed and
    ahfgpu_halo_fetch only waits for that copy; same lists as the plain fetch, also when the buffer is too small (fallback), and again
    after it is unregistered."""
    import torch
    hs = golden.hs
    with _ctx(A, golden) as g:
        g.sfc_sort(golden.pos, golden.mom, golden.weight, golden.u)
        plain = g.construct_halos(hs[:, 0:3].copy(), hs[:, 3].copy(), hs[:, 4].astype(np.int64))
        nm = len(plain["members"])
        assert nm > 0
        for cap in (nm + 100, max(nm // 2, 1)):
            buf = torch.full((nm + 100,), -7, dtype=torch.int64).pin_memory()
            g._chk(g._L.ahfgpu_halo_members_buffer(g._h, buf.data_ptr(), cap))
            g.construct_halos(hs[:, 0:3].copy(), hs[:, 3].copy(), hs[:, 4].astype(np.int64), fetch=False)
            res = g.fetch_halos(len(hs), bufs={"members": buf.numpy()})
            assert np.array_equal(res["members"], plain["members"]) and np.array_equal(res["member_offset"], plain["member_offset"])
            assert np.all(buf.numpy()[nm:] == -7)
            assert np.array_equal(res["scal"], plain["scal"]) and np.array_equal(res["prof"], plain["prof"])
        g._chk(g._L.ahfgpu_halo_members_buffer(g._h, None, 0))
        again = g.construct_halos(hs[:, 0:3].copy(), hs[:, 3].copy(), hs[:, 4].astype(np.int64))
        assert np.array_equal(again["members"], plain["members"])
