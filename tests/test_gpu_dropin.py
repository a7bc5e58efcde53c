"""GPU: the drop-in itself.  ahf_b200/host/_build/AHF-b200 is the reference's own program (main.c, startrun, readers, grid
tree, subhalo re-hash, catalogue writers -- compiled unmodified) with the key/sort call site and the per-halo loop
redirected to libahfgpu.so (ahf_b200/host/ahf_glue.c).  Its catalogues must equal those of the pure-CPU reference binary
on the same snapshot: same haloes, same particle lists, numeric columns equal to print precision."""
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
DROPIN_DIR = os.path.join(ROOT, "ahf_b200", "host", "_build")


def _num_table(path):
    rows = []
    for line in open(path):
        if line.startswith("#") or not line.strip():
            continue
        rows.append([float(t) for t in line.split()])
    return rows


# full path on the GPU / CPU mesh + GPU sort & haloes / ahf_gridinfo + ahf_halos replaced as well (device patch tables, library tree, re-hash, writers)
# AHF-b200-full reads the snapshot through the bulk ingest (ahfgpu_ingest_gadget behind io_gadget_readpart); "...:reader" = the same
# binary with AHFB200_NO_INGEST=1, i.e. the reference's own reader and the host AoS path
@pytest.mark.parametrize("variant", ["AHF-b200", "AHF-b200-kh", "AHF-b200-full", "AHF-b200-full:reader"])
@pytest.mark.parametrize("n1d,seed,ncl", [(32, 21, 6), (64, 12, 14)])
def test_catalogues_equal_reference(n1d, seed, ncl, variant):
    variant, _, mode = variant.partition(":")
    DROPIN = os.path.join(DROPIN_DIR, variant)
    from ahf_b200 import synth
    from oracle import oracle as O
    if not (os.path.exists(DROPIN) and os.path.exists(O.REF_BIN)):
        pytest.skip("drop-in / reference binaries not built (need /root/reference at build time)")
    box = synth.make_box(n1d, seed=seed, n_clumps=ncl)
    work = tempfile.mkdtemp(prefix="ahf_dropin_")
    try:
        out = {}
        for tag, exe in (("ref", O.REF_BIN), ("gpu", DROPIN)):
            d = os.path.join(work, tag)
            inp = synth.write_reference_case(box, d)
            env = dict(os.environ); env.pop("AHF_DUMP_DIR", None); env["AHFB200_TIMING"] = "1"
            if mode == "reader":
                env["AHFB200_NO_INGEST"] = "1"
            pr = subprocess.run([exe, inp], cwd=d, env=env, capture_output=True, text=True)
            assert pr.returncode == 0, pr.stderr[-3000:]
            out[tag] = d
            if tag == "gpu" and variant == "AHF-b200-full":
                assert ("ingest_gadget=" in pr.stderr) == (mode != "reader"), pr.stderr[-600:]
        pre = "ref.z0.000.AHF_"
        if variant == "AHF-b200-full":
            # what startrun derives from the file object (boxsize, pmass, no_vpart, weights, species ...) must not depend on who read the file
            pa, pb = (open(os.path.join(out[t], "ref.parameter")).read() for t in ("ref", "gpu"))
            assert pa == pb
            ha, hb = (open(os.path.join(out[t], pre + "halos")).read() for t in ("ref", "gpu"))
            assert ha == hb
        # haloes: same count, integer columns identical, float columns to the precision the writer prints
        hr, hg = _num_table(os.path.join(out["ref"], pre + "halos")), _num_table(os.path.join(out["gpu"], pre + "halos"))
        assert len(hr) == len(hg) and len(hr) >= 3
        for a, b in zip(hr, hg):
            assert a[:3] == b[:3] and a[4] == b[4]                    # ID, hostHalo, numSubStruct, npart
            assert np.allclose(a, b, rtol=2e-5, atol=1e-4), [(i, x, y) for i, (x, y) in enumerate(zip(a, b)) if not np.isclose(x, y, rtol=2e-5, atol=1e-4)]
        # particle lists: byte identical
        assert open(os.path.join(out["ref"], pre + "particles")).read() == open(os.path.join(out["gpu"], pre + "particles")).read()
        assert open(os.path.join(out["ref"], pre + "substructure")).read() == open(os.path.join(out["gpu"], pre + "substructure")).read()
        pr_, pg_ = _num_table(os.path.join(out["ref"], pre + "profiles")), _num_table(os.path.join(out["gpu"], pre + "profiles"))
        assert len(pr_) == len(pg_)
        for a, b in zip(pr_, pg_):
            assert a[1] == b[1]                                        # npart per bin
            cols = [i for i in range(len(a)) if i not in range(16, 25)] # eigenvector columns: sign convention free
            assert np.allclose(np.array(a)[cols], np.array(b)[cols], rtol=2e-5, atol=1e-4)
    finally:
        shutil.rmtree(work, ignore_errors=True)


@pytest.mark.parametrize("variant", ["AHF-b200-mm", "AHF-b200-mm-full"])
@pytest.mark.parametrize("n1d,seed,ncl", [(32, 11, 6), (64, 5, 10)])
def test_species_catalogues_equal_reference(n1d, seed, ncl, variant):
    """the multi-species build (-DMULTIMASS -DGAS_PARTICLES: gas + dark matter + stars): AHF-b200-mm against the reference's own
    multi-species binary -- the _halos file then carries the gas_only / stars_only columns, _profiles the M_gas / M_star / u_gas ones"""
    DROPIN = os.path.join(DROPIN_DIR, variant)
    from ahf_b200 import synth
    from oracle import oracle as O
    if not (os.path.exists(DROPIN) and os.path.exists(O.REF_BIN_MM)):
        pytest.skip("drop-in / reference binaries not built (need /root/reference at build time)")
    sb = synth.make_species_box(n1d, seed=seed, n_clumps=ncl)
    work = tempfile.mkdtemp(prefix="ahf_dropin_mm_")
    try:
        out = {}
        for tag, exe in (("ref", O.REF_BIN_MM), ("gpu", DROPIN)):
            d = os.path.join(work, tag)
            inp = synth.write_reference_case_species(sb, d)
            env = dict(os.environ); env.pop("AHF_DUMP_DIR", None)
            pr = subprocess.run([exe, inp], cwd=d, env=env, capture_output=True, text=True)
            assert pr.returncode == 0, pr.stderr[-3000:]
            out[tag] = d
        pre = "ref.z0.000.AHF_"
        hr, hg = _num_table(os.path.join(out["ref"], pre + "halos")), _num_table(os.path.join(out["gpu"], pre + "halos"))
        assert len(hr) == len(hg) and len(hr) >= 3 and len(hr[0]) > 43          # species columns present
        assert any(a[43] > 0 for a in hr)                                        # n_gas of some halo
        eig = set(range(26, 35)) | set(range(52, 61)) | set(range(72, 81))      # eigenvector columns (total, gas, stars): sign convention free
        for a, b in zip(hr, hg):
            assert a[:3] == b[:3] and a[4] == b[4]
            cols = [i for i in range(len(a)) if i not in eig]
            assert np.allclose(np.array(a)[cols], np.array(b)[cols], rtol=2e-5, atol=1e-4), [(i, a[i], b[i]) for i in cols if not np.isclose(a[i], b[i], rtol=2e-5, atol=1e-4)]
        assert open(os.path.join(out["ref"], pre + "particles")).read() == open(os.path.join(out["gpu"], pre + "particles")).read()
        assert open(os.path.join(out["ref"], pre + "substructure")).read() == open(os.path.join(out["gpu"], pre + "substructure")).read()
        pr_, pg_ = _num_table(os.path.join(out["ref"], pre + "profiles")), _num_table(os.path.join(out["gpu"], pre + "profiles"))
        assert len(pr_) == len(pg_)
        for a, b in zip(pr_, pg_):
            assert a[1] == b[1]
            cols = [i for i in range(len(a)) if i not in range(16, 25)]
            assert np.allclose(np.array(a)[cols], np.array(b)[cols], rtol=2e-5, atol=1e-4), [(i, a[i], b[i]) for i in cols if not np.isclose(a[i], b[i], rtol=2e-5, atol=1e-4)]
    finally:
        shutil.rmtree(work, ignore_errors=True)


def test_full_dropin_with_surviving_subhaloes():
    """AHF-b200-full on a host with sub-clumps: the sub-halo re-hash keeps most of them (the boxes above have none that survive), so
    hostHalo / numSubStruct columns and the .AHF_substructure file are exercised with content"""
    DROPIN = os.path.join(DROPIN_DIR, "AHF-b200-full")
    from ahf_b200 import synth
    from oracle import oracle as O
    if not (os.path.exists(DROPIN) and os.path.exists(O.REF_BIN)):
        pytest.skip("drop-in / reference binaries not built (need /root/reference at build time)")
    box = synth.make_host_box(60000, n_sub=8, n1d_bg=32, seed=47)
    work = tempfile.mkdtemp(prefix="ahf_dropin_sub_")
    try:
        out = {}
        for tag, exe in (("ref", O.REF_BIN), ("gpu", DROPIN)):
            d = os.path.join(work, tag)
            inp = synth.write_reference_case(box, d, lgrid_domain=64)
            env = dict(os.environ); env.pop("AHF_DUMP_DIR", None)
            pr = subprocess.run([exe, inp], cwd=d, env=env, capture_output=True, text=True)
            assert pr.returncode == 0, pr.stderr[-3000:]
            out[tag] = d
        pre = "ref.z0.000.AHF_"
        sub = open(os.path.join(out["ref"], pre + "substructure")).read()
        assert len(sub.split()) >= 6
        assert sub == open(os.path.join(out["gpu"], pre + "substructure")).read()
        assert open(os.path.join(out["ref"], pre + "particles")).read() == open(os.path.join(out["gpu"], pre + "particles")).read()
        hr, hg = _num_table(os.path.join(out["ref"], pre + "halos")), _num_table(os.path.join(out["gpu"], pre + "halos"))
        assert len(hr) == len(hg)
        for a, b in zip(hr, hg):
            assert a[:3] == b[:3] and a[4] == b[4]
            assert np.allclose(a, b, rtol=2e-5, atol=1e-4)
    finally:
        shutil.rmtree(work, ignore_errors=True)
