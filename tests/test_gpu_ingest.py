"""GPU: NEXT-4 of SURVEY 8f -- ahfgpu_ingest_gadget (bulk GADGET read + unit scaling on the device) against the UNMODIFIED reference's
reader and scaling (io_gadget.c:427-568, :857-995; oracle/_ref/ahf_ref run on the box, particles dumped after its sort): positions,
momenta, the key-sorted sequence and the header quantities must be bit-identical, for GADGET-1 / GADGET-2 framing, either byte order
and a snapshot with negative coordinates (the reader's shift)."""
import os
import shutil
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    from ahf_b200 import ahf
    return ahf


@pytest.mark.parametrize("variant", ["le_v1", "be_v1", "le_v2", "be_v2", "shifted"])
def test_ingest_equals_reference_reader(A, variant):
    from ahf_b200 import synth
    from oracle import oracle as O
    if not os.path.exists(O.REF_BIN):
        pytest.skip("oracle/_ref/ahf_ref not built")
    box = synth.make_box(32, seed=11, n_clumps=6)
    work = tempfile.mkdtemp(prefix="ahf_ingest_")
    try:
        snap = os.path.join(work, "snap.gadget")
        synth.write_gadget1(box, snap, big_endian=variant.startswith("be"), version=2 if variant.endswith("v2") else 1,
                            pos_offset=-3.5 if variant == "shifted" else 0.0)
        inp = os.path.join(work, "AHF.input")
        synth.write_ahf_input(inp, snap, os.path.join(work, "ref"), 32)
        d = os.path.join(work, "dump")
        O.run_reference(inp, dump_dir=d)
        P = O.read_particles(os.path.join(d, "particles.bin"))
        par = A.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=32)
        with A.AhfGpu(par) as g:
            # the big-endian variants go through the host-only prefetch (blocks read before the upload, as the drop-in program does)
            info, ids = g.ingest_gadget(snap, prefetch=variant.startswith("be"))
            assert int(info["n"]) == P.n == box.npart
            assert info["version"] == (2 if variant.endswith("v2") else 1) and info["swapped"] == (1 if variant.startswith("be") else 0)
            assert np.array_equal(ids, box.ids.astype(np.uint64))
            assert abs(info["boxsize"] - P.boxsize) <= 1e-12 * P.boxsize, (info["boxsize"], P.boxsize)
            g.sfc_sort_resident()
            order = g.particle_ids().astype(np.int64)                      # sorted offset -> position in the file
            pos4, mom4 = g.particles()
            # the reference's ids are the file's ID block; the synthetic files number the particles in file order
            ref_of_id = np.empty(P.n, np.int64); ref_of_id[P.ids.astype(np.int64)] = np.arange(P.n)
            fid = ids.astype(np.int64)[order]
            r = ref_of_id[fid]
            assert np.array_equal(pos4[:, :3].view(np.uint32), P.pos[r].view(np.uint32)), "positions after scaling differ from the reference's"
            assert np.array_equal(mom4[:, :3].view(np.uint32), P.mom[r].view(np.uint32)), "momenta after scaling differ from the reference's"
            keys = g.hilbert_keys(pos4[:, :3].copy())
            assert np.array_equal(keys, P.keys), "key-sorted sequence differs"
            nl = g.build_amr()
            assert nl == len([f for f in os.listdir(d) if f.startswith("flag_level_")])
    finally:
        shutil.rmtree(work, ignore_errors=True)


def test_ingest_refuses_unsupported_files(A, tmp_path):
    from ahf_b200 import synth
    sb = synth.make_species_box(16, seed=3)
    snap = str(tmp_path / "species.gadget")
    synth.write_gadget1_species(sb, snap)
    par = A.make_params(boxsize=sb.box.boxsize, pmass=sb.box.pmass, lgrid_dom=16)
    with A.AhfGpu(par) as g:
        with pytest.raises(RuntimeError):
            g.ingest_gadget(snap)
        junk = str(tmp_path / "junk.bin")
        open(junk, "wb").write(b"\x01\x02\x03\x04" * 100)
        with pytest.raises(RuntimeError):
            g.ingest_gadget(junk)
