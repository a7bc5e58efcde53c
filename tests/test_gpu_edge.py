"""GPU: degenerate and adversarial inputs of the hot path, through the C-ABI, against the CPU oracle (which is bit-identical to
the reference on every golden vector, tests/test_oracle_golden.py): empty input, a handful of particles, every particle in one
cell (maximum collision: clump-core paths of all deposit kernels, deepest refinement), particles exactly on cell faces and on
the box faces (relink's inclusive faces, periodic wrap), haloes that gather nothing / wrap around the box / are skipped."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    from ahf_b200 import ahf
    return ahf


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def _par(A, L, n, boxsize=20.0, **kw):
    pmass = 0.3 * 2.7755397e11 * boxsize ** 3 / max(n, 1)
    return A.make_params(boxsize=boxsize, pmass=pmass, lgrid_dom=L, **kw)


def _exact_domain_dens(pos, L):
    """TSC on the domain grid in float64 (density.c:344-398), the yardstick when thousands of particles share a cell: there the
    reference's float32 accumulation (and the oracle's, which repeats its summation order) carries ~sqrt(n) 2^-24 of rounding
    noise, more than the 1e-5 the north star allows between the two implementations."""
    x = pos.astype(np.float64) * L
    i = np.minimum(np.floor(x).astype(np.int64), L - 1)
    s = x - (i + 0.5)
    w = np.stack([0.5 * (0.5 - s) ** 2, 0.75 - s * s, 0.5 * (0.5 + s) ** 2], axis=0)      # [3][n][dim]
    acc = np.zeros(L ** 3)
    for c in range(3):
        for b in range(3):
            for a in range(3):
                t = (((i[:, 2] + c - 1) % L) * L + (i[:, 1] + b - 1) % L) * L + (i[:, 0] + a - 1) % L
                np.add.at(acc, t, w[c][:, 2] * w[b][:, 1] * w[a][:, 0])
    return acc * (L ** 3 / len(pos)) - 1.0


def _compare_hierarchy(A, O, pos, L, tol=1e-5, exact_domain=False):
    """sort on the GPU, then every level against the oracle: cell sets, run flags, per-node counts, densities"""
    mom = np.zeros_like(pos)
    with A.AhfGpu(_par(A, L, len(pos))) as g:
        keys, order = g.sfc_sort(pos, mom)
        okeys = O.hilbert_keys(pos)
        oorder = O.argsort_keys(okeys)
        assert np.array_equal(keys, okeys[oorder])
        ps = pos[order]                                   # the GPU's own (stable) tie order: the oracle is order independent per cell
        nl = g.build_amr()
        H = O.build_hierarchy(ps, L)
        assert nl == len(H), (nl, len(H))
        worst = 0.0
        for l in range(nl):
            G = g.level(l)
            assert np.array_equal(G.lin(), H[l].lin()), "cell set differs on level %d" % l
            assert np.array_equal(G.runflags, H[l].runflags), "run structure differs on level %d" % l
            assert np.array_equal(G.count, H[l].cnt_flag), "particles per node differ on level %d" % l
            err = np.abs(G.dens.astype(np.float64) - H[l].dens) / np.maximum(np.abs(H[l].dens), 1.0)
            worst = max(worst, float(err.max()))
            if l == 0 and exact_domain:
                ex = _exact_domain_dens(ps, L)[G.lin()]
                eg = np.abs(G.dens.astype(np.float64) - ex) / np.maximum(np.abs(ex), 1.0)
                eo = np.abs(H[0].dens.astype(np.float64) - ex) / np.maximum(np.abs(ex), 1.0)
                assert eg.max() <= 2e-7, eg.max()          # fixed-point accumulation: float32 rounding of the result only
                print("domain level vs float64 TSC: GPU %.2e, oracle (float32 accumulation) %.2e" % (eg.max(), eo.max()))
        assert worst <= tol, worst
        return nl


def test_empty_input(A):
    pos = np.zeros((0, 3), np.float32)
    with A.AhfGpu(_par(A, 16, 0)) as g:
        keys, order = g.sfc_sort(pos, pos.copy())
        assert keys.shape == (0,) and order.shape == (0,)
        assert g.build_amr() == 1                         # the domain grid alone
        G = g.level(0)
        assert G.ncell == 16 ** 3 and np.all(G.dens == -1.0) and np.all(G.count == 0)
        res = g.construct_halos(np.zeros((0, 3)), np.zeros(0), np.zeros(0, np.int64))
        assert res["scal"].shape[0] == 0
        # haloes on an empty box gather nothing
        res = g.construct_halos(np.array([[0.5, 0.5, 0.5]]), np.array([0.1]), np.array([10], np.int64))
        assert res["scal"][0, 9] == 0 and len(g.halo_members(res, 0)) == 0


@pytest.mark.parametrize("n", [1, 2, 31, 33, 257])
def test_few_particles(A, O, n):
    rng = np.random.default_rng(100 + n)
    pos = rng.random((n, 3)).astype(np.float32)
    assert _compare_hierarchy(A, O, pos, 8) >= 1


@pytest.mark.parametrize("n,L", [(5000, 16), (40000, 32)])
def test_all_particles_in_one_cell(A, O, n, L):
    """maximum collision: one domain cell holds everything, refinement goes as deep as the particles allow"""
    rng = np.random.default_rng(7)
    c0 = (np.array([3, 5, 2]) + 0.5) / L
    pos = (c0 + (rng.random((n, 3)) - 0.5) * (0.98 / L)).astype(np.float32)
    # 1e-4 against the oracle: see _exact_domain_dens; the GPU's domain level is held to 2e-7 against float64
    assert _compare_hierarchy(A, O, pos, L, tol=1e-4, exact_domain=True) >= 3


def test_identical_positions(A, O):
    """many particles with the SAME coordinates (equal keys, equal weights): stable tie order, one cell on every level"""
    rng = np.random.default_rng(9)
    base = rng.random((40, 3)).astype(np.float32)
    pos = np.repeat(base, 200, axis=0)
    rng.shuffle(pos, axis=0)
    with A.AhfGpu(_par(A, 16, len(pos))) as g:
        keys, order = g.sfc_sort(pos, np.zeros_like(pos))
        assert np.all(keys[1:] >= keys[:-1])
        same = keys[1:] == keys[:-1]
        assert np.all(order[1:][same] > order[:-1][same])        # ties keep the input order
    _compare_hierarchy(A, O, pos, 16, tol=1e-4, exact_domain=True)


def test_particles_on_cell_and_box_faces(A, O):
    """coordinates that are exact multiples of the cell size on several levels (relink.c:153 inclusive faces), x = 0 and the
    largest float below 1 (lltools.c:61-64 clamp, periodic wrap of the TSC stencil)"""
    rng = np.random.default_rng(21)
    L = 16
    n = 6000
    clump = (np.array([0.5, 0.5, 0.5]) + rng.normal(0, 0.01, (n, 3)))
    # snap a third of the coordinates onto faces of the level-0..5 grids
    snap = rng.random((n, 3)) < 0.33
    lev = rng.integers(0, 6, (n, 3))
    q = np.round(clump * (L * 2.0 ** lev)) / (L * 2.0 ** lev)
    clump = np.where(snap, q, clump)
    edge = rng.random((3000, 3))
    edge[:1000, 0] = 0.0
    edge[1000:2000, 1] = np.nextafter(np.float32(1.0), np.float32(0.0))
    edge[2000:, 2] = rng.choice([0.0, float(np.nextafter(np.float32(1.0), np.float32(0.0)))], 1000)
    corner = rng.normal(0, 0.004, (3000, 3))                                     # a clump on the box corner, wrapped on all faces
    pos = np.mod(np.concatenate([clump, edge, corner]), 1.0).astype(np.float32)
    pos = np.minimum(pos, np.nextafter(np.float32(1.0), np.float32(0.0)))
    _compare_hierarchy(A, O, pos, L)


def test_halo_edge_cases(A, O):
    """haloes with seed count 0 (skipped, ahf_halos_sfc.c:122), a gathering sphere that holds nothing, one below NminPerHalo,
    one wrapped around the box corner -- against the oracle's halo pass"""
    from ahf_b200 import synth
    box = synth.make_box(32, seed=3, n_clumps=4, centres_box=np.array([[0.001, 0.999, 0.002], [0.5, 0.5, 0.5]]))
    par = A.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=32)
    c, r, npart = synth.halo_seeds(box)
    c = np.concatenate([c, [[0.25, 0.75, 0.25], [0.1, 0.1, 0.9], [0.7, 0.2, 0.3]]])
    r = np.concatenate([r, [1e-7, 0.004, r[0]]])             # nothing inside / a handful of lattice particles / a normal radius
    npart = np.concatenate([npart, [50, 50, 0]])             # the last one is skipped by its seed count
    with A.AhfGpu(par) as g:
        keys, order = g.sfc_sort(box.pos, box.mom)
        pos, mom = box.pos[order], box.mom[order]
        res = g.construct_halos(c, r, npart)
        opar = dict(r_fac=par.r_fac, x_fac=par.x_fac, v_fac=par.v_fac, m_fac=par.m_fac, rho_fac=par.rho_fac, phi_fac=par.phi_fac,
                    Hubble=par.hubble, ovlim=par.ovlim, rho_vir=par.rho_vir, vesc_tune=par.vesc_tune, min_part=par.min_part)
        ores = O.construct_halos(keys, pos, mom, None, None, opar, c, r, npart)
        S = res["scal"]
        assert S[-1, 9] == 0 and S[-3, 9] == 0
        for i, o in enumerate(ores):
            assert int(S[i, 9]) == o["npart"], (i, S[i, 5:10], o["npart"])
            assert np.array_equal(g.halo_members(res, i), o["ipart"]), i
            if o["npart"] >= par.min_part:
                assert np.allclose(S[i, 10:14], o["s"][10:14], rtol=1e-9), i
        assert int(S[0, 9]) >= par.min_part                  # the wrapped clump is a real halo


@pytest.mark.parametrize("case", ["pairs", "runs32", "long", "equal", "sorted_input"])
def test_key_sort_tie_fix_equals_the_full_stable_sort(A, O, case, monkeypatch):
    """K2 sorts bits 15..62 of the Hilbert keys in six radix passes and orders the particles that share all of them (one cell of
    2^-16 of the box per dimension) by the whole key afterwards (sfc.cu: k_fix_key_ties); runs longer than 32 fall back to the
    eight-pass sort.  Keys and permutation must equal the stable sort of the oracle's keys in every case."""
    rng = np.random.default_rng(11)
    n = 60000
    pos = rng.random((n, 3), dtype=np.float32)
    cell = np.float32(2.0 ** -16)
    if case == "pairs":            # 2..4 particles per 2^-16 cell, distinct low bits
        base = np.floor(rng.random((n // 3, 3)) * 65536).astype(np.float32) * cell
        pos = (np.repeat(base, 3, axis=0) + rng.random((n, 3), dtype=np.float32) * cell).astype(np.float32)
    elif case == "runs32":         # up to 32 per cell: the longest run the insertion sort takes
        base = np.floor(rng.random((n // 30, 3)) * 65536).astype(np.float32) * cell
        pos = (np.repeat(base, 30, axis=0) + rng.random((n, 3), dtype=np.float32) * cell).astype(np.float32)
    elif case == "long":           # thousands of distinct keys inside one cell: the fallback
        pos[:5000] = (np.float32(0.3) + rng.random((5000, 3), dtype=np.float32) * cell).astype(np.float32)
    elif case == "equal":          # thousands of identical particles: equal FULL keys keep their input order
        pos[1000:4000] = pos[7]
        pos[10000:10020] = pos[8]
    elif case == "sorted_input":
        pass
    pos = np.clip(pos, 0.0, np.float32(1.0) - np.float32(2.0 ** -24)).astype(np.float32)
    rng.shuffle(pos, axis=0)
    mom = rng.standard_normal((n, 3)).astype(np.float32)
    okeys = O.hilbert_keys(pos)
    if case == "sorted_input":
        o = np.argsort(okeys, kind="stable"); pos, mom, okeys = pos[o], mom[o], okeys[o]
    oorder = np.argsort(okeys, kind="stable")
    top = okeys[oorder] >> np.uint64(15)
    runlen = np.diff(np.flatnonzero(np.concatenate([[True], top[1:] != top[:-1], [True]])))
    with A.AhfGpu(_par(A, 32, n)) as g:
        keys, order = g.sfc_sort(pos, mom)
        fell_back = g.stage_count("sort_full_fallback")
    assert np.array_equal(keys, okeys[oorder])
    assert np.array_equal(order.astype(np.int64), oorder)
    assert fell_back == (1 if runlen.max() > 32 else 0), (case, runlen.max(), fell_back)
    if case in ("pairs", "runs32"):
        assert runlen.max() >= 2 and runlen.max() <= 32
    monkeypatch.setenv("AHFGPU_SORT_FULL", "1")
    with A.AhfGpu(_par(A, 32, n)) as g:
        keys8, order8 = g.sfc_sort(pos, mom)
    assert np.array_equal(keys8, keys) and np.array_equal(order8, order)
