"""Regenerates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/ahf_ref, built from
/root/reference by oracle/build_ref.sh) on small synthetic boxes and collecting the hook dumps
(oracle/ref_hooks.c).  Only runnable where the reference binary exists; the .npz files are committed.

    python tests/golden/make_golden.py              # the per-case fixtures
    python tests/golden/make_golden.py patches      # tests/golden/patches.npz: the patch colouring of ahf_gridinfo for the same cases
    python tests/golden/make_golden.py gridtree     # tests/golden/gridtree.npz: the reference's own .AHF_gridtree (-DAHFgridtreefile build)
"""
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from ahf_b200 import synth          # noqa: E402
from oracle import oracle as O      # noqa: E402

CASES = {
    # name: (n1d, seed, n_clumps, nper_dom, nper_ref, explicit clump centres (box units) or None)
    "edge32": (32, 7, 8, 2.0, 2.5, [[0.0, 0.0, 0.0], [0.999, 0.5, 0.5], [0.5, 0.001, 0.999], [0.25, 0.999, 0.0005], [0.5, 0.5, 0.0]]),
    "frag16": (16, 3, 4, 1.1, 1.3, [[0.02, 0.5, 0.97]]),
    "plain32": (32, 42, 6, 2.0, 2.5, None),
}


# multi-species cases run the -DMULTIMASS -DGAS_PARTICLES build (oracle/_ref/ahf_ref_mm): gas + dark matter + stars
CASES_MM = {
    "species32": (32, 11, 6, 2.0, 2.5, None),
}


def collect(name, n1d, seed, ncl, nper_dom, nper_ref, centres, species=False):
    work = tempfile.mkdtemp(prefix="ahf_golden_")
    try:
        if species:
            sb = synth.make_species_box(n1d, seed=seed, n_clumps=ncl, centres_box=None if centres is None else np.array(centres))
            inp = synth.write_reference_case_species(sb, work, nper_dom=nper_dom, nper_ref=nper_ref)
        else:
            box = synth.make_box(n1d, seed=seed, n_clumps=ncl, centres_box=None if centres is None else np.array(centres))
            inp = synth.write_reference_case(box, work, nper_dom=nper_dom, nper_ref=nper_ref)
        O.run_reference(inp, dump_dir=os.path.join(work, "dump"), threads=1, multimass=species)
        d = os.path.join(work, "dump")
        P = O.read_particles(os.path.join(d, "particles.bin"), multimass=species)
        out = dict(n1d=n1d, seed=seed, nper_dom=nper_dom, nper_ref=nper_ref, boxsize=P.boxsize, pmass=P.pmass,
                   ids=P.ids.astype(np.uint32), keys=P.keys, pos=P.pos, mom=P.mom)   # key-sorted; input (file) order: x_in[ids] = x
        if species:
            out["weight"] = P.weight; out["u"] = P.u
        nlev = 0
        while os.path.exists(os.path.join(d, "flag_level_%02d.bin" % nlev)):
            R = O.read_level(os.path.join(d, "flag_level_%02d.bin" % nlev))
            F = O.read_level(os.path.join(d, "final_level_%02d.bin" % nlev))
            p = "L%d_" % nlev
            out[p + "l1dim"] = R.l1dim; out[p + "critdens"] = R.critdens; out[p + "masstopartdens"] = R.masstopartdens
            out[p + "x"] = R.x; out[p + "y"] = R.y; out[p + "z"] = R.z; out[p + "dens"] = R.dens
            out[p + "runflags"] = R.runflags; out[p + "cnt"] = R.cnt_flag; out[p + "plist"] = R.plist_flag.astype(np.int32)
            out[p + "cnt_final"] = F.cnt_flag; out[p + "plist_final"] = F.plist_flag.astype(np.int32)
            nlev += 1
        out["nlev"] = nlev
        H = O.read_halos(d)
        out["halo_glob"] = H.glob; out["halo_s"] = H.s
        out["halo_moff"] = np.concatenate([[0], np.cumsum([len(m) for m in H.members])]).astype(np.int64)
        out["halo_members"] = np.concatenate(H.members).astype(np.int32) if H.n else np.zeros(0, np.int32)
        nb = [0 if p is None else p.shape[1] for p in H.prof]
        out["halo_poff"] = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
        out["halo_prof"] = np.concatenate([p.reshape(-1) for p in H.prof if p is not None]) if sum(nb) else np.zeros(0)
        if H.species is not None:
            out["halo_species"] = H.species
            out["halo_prof_species"] = np.concatenate([p.reshape(-1) for p in H.prof_species if p is not None]) if sum(nb) else np.zeros(0)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
        print(name, "levels", nlev, "halos", H.n, "bytes", os.path.getsize(os.path.join(ROOT, "tests", "golden", name + ".npz")))
    finally:
        shutil.rmtree(work, ignore_errors=True)


def collect_patches():
    """patches.npz: per case the first coloured level (ahf.min_ref) and, per coloured level, the isolated-refinement index of every
    node in traversal order plus the periodic flags of the isolated refinements (oracle/ref_hooks.c dump_patches)."""
    out = {}
    for table, species in ((CASES, False), (CASES_MM, True)):
        for name, (n1d, seed, ncl, nper_dom, nper_ref, centres) in table.items():
            work = tempfile.mkdtemp(prefix="ahf_golden_")
            try:
                cb = None if centres is None else np.array(centres)
                if species:
                    inp = synth.write_reference_case_species(synth.make_species_box(n1d, seed=seed, n_clumps=ncl, centres_box=cb), work,
                                                             nper_dom=nper_dom, nper_ref=nper_ref)
                else:
                    inp = synth.write_reference_case(synth.make_box(n1d, seed=seed, n_clumps=ncl, centres_box=cb), work,
                                                     nper_dom=nper_dom, nper_ref=nper_ref)
                O.run_reference(inp, dump_dir=os.path.join(work, "dump"), threads=1, multimass=species)
                min_ref, levels = O.read_patches(os.path.join(work, "dump", "patches.bin"))
                out[name + "_min_ref"] = min_ref; out[name + "_nlev"] = len(levels)
                for i, (iso, per) in enumerate(levels):
                    out["%s_L%d_iso" % (name, min_ref + i)] = iso.astype(np.int16 if iso.max() < 32000 else np.int32)
                    out["%s_L%d_per" % (name, min_ref + i)] = per.astype(np.uint8)
                print(name, "min_ref", min_ref, "coloured levels", len(levels), "patches", [p.shape[0] for _, p in levels])
            finally:
                shutil.rmtree(work, ignore_errors=True)
    path = os.path.join(ROOT, "tests", "golden", "patches.npz")
    np.savez_compressed(path, **out)
    print("patches.npz bytes", os.path.getsize(path))


def collect_gridtree():
    """gridtree.npz: what the reference built with its -DAHFgridtreefile option (oracle/_ref/ahf_ref_gt) writes after RefCentre and
    analyseRef (ahf_halos.c:3066-3120): per isolated refinement the centre, closeRefDist, node / particle counts, the daughter and the
    substructure links.  Default-build cases only."""
    import glob
    import subprocess
    out = {}
    for name, (n1d, seed, ncl, nper_dom, nper_ref, centres) in CASES.items():
        work = tempfile.mkdtemp(prefix="ahf_golden_")
        try:
            cb = None if centres is None else np.array(centres)
            inp = synth.write_reference_case(synth.make_box(n1d, seed=seed, n_clumps=ncl, centres_box=cb), work, nper_dom=nper_dom, nper_ref=nper_ref)
            env = dict(os.environ, OMP_NUM_THREADS="1")
            env.pop("AHF_DUMP_DIR", None)
            subprocess.run([O.REF_BIN_GT, inp], cwd=work, env=env, capture_output=True, text=True, check=True)
            min_ref, T = O.read_gridtree(glob.glob(os.path.join(work, "*.AHF_gridtree"))[0])
            out[name + "_min_ref"] = min_ref; out[name + "_nlev"] = len(T)
            for lev, t in T.items():
                p = "%s_L%d_" % (name, lev)
                out[p + "centre"] = t["centre"]; out[p + "close"] = t["close"]; out[p + "nodes"] = t["nodes"]; out[p + "parts"] = t["parts"]
                out[p + "daughter"] = t["daughter"]
                out[p + "sub_off"] = np.concatenate([[0], np.cumsum([len(x) for x in t["sub"]])]).astype(np.int64)
                out[p + "sub"] = np.array([y for x in t["sub"] for y in x], np.int64).reshape(-1, 2)
            print(name, "min_ref", min_ref, "patches per level", [len(t["nodes"]) for t in T.values()])
        finally:
            shutil.rmtree(work, ignore_errors=True)
    path = os.path.join(ROOT, "tests", "golden", "gridtree.npz")
    np.savez_compressed(path, **out)
    print("gridtree.npz bytes", os.path.getsize(path))


if __name__ == "__main__":
    only = sys.argv[1:]
    if only == ["gridtree"]:
        collect_gridtree()
        sys.exit(0)
    if only == ["patches"]:
        collect_patches()
        sys.exit(0)
    for k, v in CASES.items():
        if not only or k in only:
            collect(k, *v)
    for k, v in CASES_MM.items():
        if not only or k in only:
            collect(k, *v, species=True)
