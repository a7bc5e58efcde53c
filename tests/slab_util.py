"""Checker shared by the slab-decomposition tests (in-process ranks on one GPU: tests/test_gpu_slab.py; NCCL ranks under torchrun:
tests/mgpu_check.py): what a rank of a split box holds on ITS OWN cells / particles / haloes must equal the single-GPU run bit for bit."""
import numpy as np


def single_gpu_truth(A, box, n1d, centres, rad, seed, device=0, weight=None, u=None):
    """the whole box on one GPU: per level {cell key -> (dens bits, mark, runflags, count)}, final level of every particle, halo table"""
    par = A.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d, device=device)
    T = {}
    with A.AhfGpu(par) as g:
        keys, order = g.sfc_sort(box.pos, box.mom, weight, u)
        nl = g.build_amr()
        T["nl"] = nl
        T["levels"] = []
        for l in range(nl):
            G = g.level(l)
            T["levels"].append(dict(lin=G.lin(), dens=G.dens.view(np.uint32).copy(), mark=G.mark.copy(), runflags=G.runflags.copy(), count=G.count.copy(),
                                    interior=G.interior.copy(), header=g.level_header(l)))
        owner, _ = g.particle_levels(with_cells=False)
        by_id = np.empty(len(order), np.int8); by_id[order] = owner
        T["owner_by_id"] = by_id
        res = g.construct_halos(centres, rad, seed)
        T["scal"] = res["scal"]
        T["members"] = [order[g.halo_members(res, i)].astype(np.int64) for i in range(len(rad))]
        T["prof"] = [g.halo_profile(res, i) for i in range(len(rad))]
    return T


def rank_report(sb, centres, rad, seed):
    """everything a rank contributes, restricted to what it owns"""
    g = sb.g
    nl_box = sb.build_amr()
    info = g.slab_info()
    out = dict(rank=sb.rank, nl_box=nl_box, info=info, levels=[])
    for l in range(g.nlevels()):
        G = g.level(l)
        own = g.level_owned(l)
        out["levels"].append(dict(lin=G.lin()[own], dens=G.dens.view(np.uint32)[own].copy(), mark=G.mark[own].copy(), runflags=G.runflags[own].copy(),
                                  count=G.count[own].copy(), interior=G.interior[own].copy(), header=g.level_header(l)))
    owner, _ = g.particle_levels(with_cells=False)
    ids = g.particle_ids().astype(np.int64)
    lo, hi = info["own_lo"], info["own_hi"]
    out["own_ids"] = ids[lo:hi]
    out["own_level"] = owner[lo:hi].copy()
    mine, res = sb.construct_halos(centres, rad, seed)
    out["mine"] = mine
    out["scal"] = res["scal"]
    out["members"] = [g.halo_members(res, k).astype(np.int64) for k in range(len(mine))]
    out["prof"] = [g.halo_profile(res, k) for k in range(len(mine))]
    return out


def check_against_truth(reports, T, npart):
    """union of the ranks' own parts == the single-GPU result"""
    nl = T["nl"]
    assert all(r["nl_box"] == nl for r in reports), ([r["nl_box"] for r in reports], nl)
    for l in range(nl):
        parts = [r["levels"][l] for r in reports if l < len(r["levels"])]
        lin = np.concatenate([p["lin"] for p in parts])
        o = np.argsort(lin, kind="stable")
        ref = T["levels"][l]
        assert len(lin) == len(ref["lin"]) and np.array_equal(lin[o], ref["lin"]), "level %d: the ranks' own cells are not a partition of the level (%d vs %d)" % (l, len(lin), len(ref["lin"]))
        for k in ("dens", "mark", "runflags", "count", "interior"):
            v = np.concatenate([p[k] for p in parts])[o]
            bad = np.nonzero(v != ref[k])[0]
            assert bad.size == 0, "level %d: %s differs on %d own cells (first %s)" % (l, k, bad.size, bad[:5])
        # box-wide counts: particles deposited on the level = sum over ranks of their own share; finally owned likewise
        assert sum(int(p["header"][0][3]) for p in parts) == int(ref["header"][0][3]), "level %d: particles finally owned" % l
    ids = np.concatenate([r["own_ids"] for r in reports])
    assert len(ids) == npart and np.array_equal(np.sort(ids), np.arange(npart)), "own particles are not a partition of the box"
    lev = np.concatenate([r["own_level"] for r in reports])
    got = np.empty(npart, np.int8); got[ids] = lev
    assert np.array_equal(got, T["owner_by_id"]), "final level of the particles"
    nh = len(T["scal"])
    seen = np.zeros(nh, int)
    for r in reports:
        for k, h in enumerate(r["mine"]):
            seen[h] += 1
            assert np.array_equal(r["scal"][k], T["scal"][h], equal_nan=True), ("halo scalars", h, r["rank"])
            assert np.array_equal(r["members"][k], T["members"][h]), ("halo members (global ids)", h, r["rank"])
            a, b = r["prof"][k], T["prof"][h]
            assert (a is None) == (b is None) and (a is None or np.array_equal(a, b, equal_nan=True)), ("halo profile", h)
    assert np.all(seen == 1), "every halo must be served by exactly one rank"
    return dict(levels=nl, halos=nh, resident=[r["info"]["resident"] for r in reports], own=[r["info"]["own_hi"] - r["info"]["own_lo"] for r in reports])
