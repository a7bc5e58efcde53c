"""BASELINE.json configs[4] at size: a multi-species (dark matter + gas + stars, -DMULTIMASS -DGAS_PARTICLES semantics) host halo of
`n_host` particles (default 10^7) with subclumps -- the cooperative multi-block halo pass against the CPU oracle (oracle/, the C
restatement of ahf_halos_sfc_constructHalo pinned bit for bit on the reference's own dumps).  Prints one JSON object.
  python scripts/big_host_check.py [n_host] [n_sub_checked]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                    # noqa: E402
from ahf_b200 import ahf, synth       # noqa: E402
from oracle import oracle as O        # noqa: E402


def species_of(box, seed=77, gas_frac=0.15, star_frac=0.05):
    rng = np.random.default_rng([seed, 5])
    t = rng.random(box.npart)
    typ = np.where(t < gas_frac, 0, np.where(t < gas_frac + star_frac, 4, 1))
    w = np.where(typ == 0, 0.2, np.where(typ == 4, 0.1, 1.0)).astype(np.float32)
    u = np.where(typ == 0, rng.uniform(1e3, 2e5, size=box.npart), np.where(typ == 4, synth.PSTAR, synth.PDM)).astype(np.float32)
    return w, u


def run(n_host, n_check, species=True):
    t0 = time.time()
    box = synth.make_host_box(n_host)
    w, u = species_of(box) if species else (None, None)
    c, r, npart = synth.halo_seeds(box)
    par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=128)
    out = dict(n_host=n_host, particles=box.npart, species=species, gen_s=time.time() - t0)
    with ahf.AhfGpu(par) as g:
        keys, order = g.sfc_sort(box.pos, box.mom, w, u)
        for it in range(2):
            g.synchronize(); t0 = time.perf_counter()
            res = g.construct_halos(c, r, npart, fetch=(it == 1))
            g.synchronize(); wall = time.perf_counter() - t0
        out["gpu_halo_pass_ms"] = {k: g.stage_ms(k) for k in ("halo_gather", "halo_sort", "halo_unbind", "halo_profiles")}
        out["gpu_wall_ms_incl_fetch"] = wall * 1e3
        out["unbind_iterations"] = g.stage_count("halo_unbind_iterations")
        out["unbind_iter_members"] = g.stage_count("halo_unbind_iter_members")
        S = res["scal"]
        out["host_gathered"] = int(S[0, 5]); out["host_npart"] = int(S[0, 9])
        pos, mom = box.pos[order], box.mom[order]
        ws, us = (w[order], u[order]) if species else (None, None)
        opar = dict(r_fac=par.r_fac, x_fac=par.x_fac, v_fac=par.v_fac, m_fac=par.m_fac, rho_fac=par.rho_fac, phi_fac=par.phi_fac,
                    Hubble=par.hubble, ovlim=par.ovlim, rho_vir=par.rho_vir, vesc_tune=par.vesc_tune, min_part=par.min_part)
        sel = [0] + list(np.argsort(-npart[1:])[:n_check] + 1)         # the host and the richest subclumps
        t0 = time.time()
        ores = O.construct_halos(keys, pos, mom, ws, us, opar, c[sel], r[sel], npart[sel])
        out["oracle_s"] = time.time() - t0
        worst_s = worst_p = 0.0
        slots = [10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37,
                 38, 39, 40, 41, 42, 43, 53, 54, 55, 56, 57]
        for k, h in enumerate(sel):
            o = ores[k]
            assert [int(S[h, 5]), int(S[h, 6]), int(S[h, 7]), int(S[h, 8]), int(S[h, 9])] == [o["n_gather"], o["n_rvir0"], o["n_unbound"], o["n_rvir1"], o["npart"]], (h, S[h, 5:10])
            assert np.array_equal(g.halo_members(res, h), o["ipart"]), ("members", h)
            if o["npart"] < par.min_part:
                continue
            a, b = o["s"][slots], S[h, slots]
            ok = np.isclose(a, b, rtol=1e-8, atol=1e-300)
            assert ok.all(), (h, [(slots[q], a[q], b[q]) for q in np.nonzero(~ok)[0]])
            nz = a != 0
            worst_s = max(worst_s, float(np.max(np.abs(a[nz] - b[nz]) / np.abs(a[nz]))))
            pr, pg = o["prof"], g.halo_profile(res, h)
            cols = [q for q in range(25) if q not in (14, 15, 16, 18, 19, 20, 22, 23, 24)]
            okp = np.isclose(pr[cols], pg[cols], rtol=1e-7, atol=1e-300)
            assert okp.all(), (h, np.argwhere(~okp)[:5])
            nzp = pr[cols] != 0
            worst_p = max(worst_p, float(np.max(np.abs(pr[cols][nzp] - pg[cols][nzp]) / np.abs(pr[cols][nzp]))))
            if species:
                sl = [q + 32 * t for t in (0, 1) for q in list(range(0, 19)) + [28, 29]]
                assert np.allclose(o["species"][sl], res["species"][h][sl], rtol=1e-8, atol=1e-300), ("species block", h)
        out.update(checked_halos=len(sel), members_identical=True, scalar_max_rel_err=worst_s, profile_max_rel_err=worst_p)
    return out


if __name__ == "__main__":
    n_host = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    n_check = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    print(json.dumps(run(n_host, n_check)))
