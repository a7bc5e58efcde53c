"""CPU-only: host-side logic, the C-ABI surface, the synthetic generator, multi-rank partitioning (gloo)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_exports_every_declared_symbol():
    from ahf_b200 import ahf
    ahf.build()
    L = C.CDLL(ahf.LIB_PATH)
    names = ahf.exported_symbols()
    assert len(names) >= 20
    for s in names:
        assert hasattr(L, s), s


def test_no_cpu_fallback_without_device():
    from ahf_b200 import ahf
    L = ahf.lib()
    if L.ahfgpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    par = ahf.make_params(boxsize=64.0, pmass=1e10, lgrid_dom=64)
    with pytest.raises(ahf.AhfGpuError):
        ahf.AhfGpu(par)


def test_params_follow_reference_unit_factors(golden):
    from ahf_b200 import ahf
    g = golden.glob      # r_fac x_fac v_fac m_fac rho_fac phi_fac Hubble ovlim rho_vir minpart vtune maxgather a
    p = ahf.make_params(boxsize=float(golden.d["boxsize"]), pmass=float(golden.d["pmass"]), lgrid_dom=golden.n1d)
    mine = [p.r_fac, p.x_fac, p.v_fac, p.m_fac, p.rho_fac, p.phi_fac, p.hubble, p.ovlim, p.rho_vir, p.min_part, p.vesc_tune]
    assert np.allclose(mine, g[:11], rtol=1e-6), (mine, g[:11])


def test_synth_box_and_gadget_roundtrip(tmp_path):
    from ahf_b200 import synth
    b = synth.make_box(16, seed=1, n_clumps=3)
    assert b.pos.shape == (4096, 3) and b.pos.dtype == np.float32
    assert b.pos.min() >= 0.0 and b.pos.max() < 1.0
    inp = synth.write_reference_case(b, str(tmp_path))
    raw = open(os.path.join(tmp_path, "snap.gadget"), "rb").read()
    assert int.from_bytes(raw[:4], "little") == 256
    npart = np.frombuffer(raw[4:4 + 24], "<i4")
    assert npart[1] == 4096 and npart.sum() == 4096
    x = np.frombuffer(raw[4 + 256 + 4 + 4:4 + 256 + 4 + 4 + 4096 * 12], "<f4").reshape(-1, 3)
    assert np.array_equal((x * np.float32(1.0 / b.boxsize)).astype(np.float32), b.pos)      # exact: box is 2^k
    assert "LgridDomain       = 16" in open(inp).read()
    c, r, n = synth.halo_seeds(b)
    assert len(r) == 3 and r.max() <= 0.25 and (n >= 30).all()


def test_lpt_partition_properties():
    from ahf_b200 import parallel as P
    rng = np.random.default_rng(0)
    w = rng.pareto(1.2, 500) * 100 + 20
    a = P.assign_halos_lpt(w, 8)
    loads = np.bincount(a, weights=w, minlength=8)
    assert set(a.tolist()) <= set(range(8))
    assert loads.max() <= loads.mean() + w.max()            # LPT bound
    assert np.array_equal(a, P.assign_halos_lpt(w, 8))      # deterministic
    b = P.slab_bounds(1000003, 8)
    assert b[0] == 0 and b[-1] == 1000003 and np.all(np.diff(b) >= 125000)


_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from ahf_b200 import parallel as P, synth
from oracle import oracle as O
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
box = synth.make_box(16, seed=9, n_clumps=4)
keys = O.hilbert_keys(box.pos); order = O.argsort_keys(keys); keys = keys[order]
b = P.slab_bounds(len(keys), 2)
mine = np.arange(b[rank], b[rank + 1])
# every rank owns a contiguous SFC slab; together they cover the box exactly once
cnt = torch.tensor([len(mine)]); dist.all_reduce(cnt); assert int(cnt) == len(keys)
lo = torch.tensor([int(keys[mine[0]] >> np.uint64(1))]); hi = torch.tensor([int(keys[mine[-1]] >> np.uint64(1))])
los = [torch.zeros(1, dtype=torch.int64) for _ in range(2)]; his = [torch.zeros(1, dtype=torch.int64) for _ in range(2)]
dist.all_gather(los, lo); dist.all_gather(his, hi)
assert int(his[0]) <= int(los[1])
# haloes: same LPT assignment on every rank, each halo constructed by exactly one rank, results identical to a serial run
c, r, n = synth.halo_seeds(box)
a = P.assign_halos_lpt(n, 2)
par = dict(r_fac=box.boxsize, x_fac=box.boxsize, v_fac=box.boxsize*100, m_fac=box.pmass, rho_fac=box.pmass/box.boxsize**3,
           phi_fac=4.3006485e-9*box.pmass/box.boxsize, Hubble=100.0, ovlim=200.0, rho_vir=2.7755397e11, vesc_tune=1.5, min_part=20)
pos = box.pos[order]; mom = box.mom[order]
sel = np.nonzero(a == rank)[0]
res = O.construct_halos(keys, pos, mom, None, None, par, c[sel], r[sel], n[sel])
np_local = torch.zeros(len(r), dtype=torch.int64)
for k, h in enumerate(sel): np_local[h] = res[k]["npart"]
dist.all_reduce(np_local)
serial = O.construct_halos(keys, pos, mom, None, None, par, c, r, n)
assert [int(v) for v in np_local] == [s["npart"] for s in serial]
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_partition_gloo(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % dict(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


_WORKER_CAT = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from ahf_b200 import ahf, multigpu, parallel as P, synth
from oracle import oracle as O
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
out_dir = sys.argv[2]
box = synth.make_host_box(30000, n_sub=12, n1d_bg=32, seed=48)            # a host halo with sub-haloes: the re-hash has work to do
n1d = 64
keys = O.hilbert_keys(box.pos); order = np.argsort(keys, kind="stable"); keys = keys[order]
pos = box.pos[order]; mom = box.mom[order]
par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d)
# every rank derives the same seeds (CPU oracle hierarchy -> the library's host tree), as every rank of a split box does
H = O.build_hierarchy(pos, n1d, patches=True)
m = ahf.min_ref(par, [lv.l1dim for lv in H])
stats = [np.hstack([lv.patch, O.patch_extents(lv).reshape(len(lv.patch), 6)]) for lv in H[m:]]
seeds = ahf.tree_halos(stats, 3.0 / box.boxsize)
c, r, n = np.ascontiguousarray(seeds["pos"]), np.ascontiguousarray(seeds["gather_rad"]), np.ascontiguousarray(seeds["npart"], np.int64)
assert len(n) >= 8
opar = dict(r_fac=par.r_fac, x_fac=par.x_fac, v_fac=par.v_fac, m_fac=par.m_fac, rho_fac=par.rho_fac, phi_fac=par.phi_fac,
            Hubble=par.hubble, ovlim=par.ovlim, rho_vir=par.rho_vir, vesc_tune=par.vesc_tune, min_part=par.min_part)

def serve(sel):
    # the halo pass of the haloes `sel` (CPU oracle) in the layout ahfgpu_halo_fetch delivers
    rs = O.construct_halos(keys, pos, mom, None, None, opar, c[sel], r[sel], n[sel])
    scal = np.array([q["s"] for q in rs]).reshape(len(sel), 64)
    moff = np.zeros(len(sel) + 1, np.int64); poff = np.zeros(len(sel) + 1, np.int64)
    for k, q in enumerate(rs):
        moff[k + 1] = moff[k] + q["npart"]; poff[k + 1] = poff[k] + q["nbins"]
    mem = np.concatenate([q["ipart"].astype(np.int64) for q in rs]) if moff[-1] else np.zeros(0, np.int64)
    prof = np.concatenate([q["prof"].reshape(-1) for q in rs if q["nbins"] > 0]) if poff[-1] else np.zeros(0)
    return dict(scal=scal, member_offset=moff, members=mem, prof_offset=poff, prof=prof)

mine = np.nonzero(P.assign_halos_lpt(n, 2) == rank)[0]                     # the same assignment on every rank
assert len(mine) > 0
parts = multigpu.gather_parts_torch(mine, serve(mine), dst=0)              # results travel to rank 0 (gloo here, NCCL's host side on a GPU box)
if rank == 0:
    assert sorted(np.concatenate([p[0] for p in parts]).tolist()) == list(range(len(n)))
    ids = order.astype(np.uint64)                                          # ID of the particle at sorted offset i = its input index
    two = multigpu.catalogue_from_ranks(os.path.join(out_dir, "two.z0.000"), par, seeds, parts, ids)
    allh = np.arange(len(n))
    one = multigpu.catalogue_from_ranks(os.path.join(out_dir, "one.z0.000"), par, seeds, [(allh, serve(allh))], ids)
    for k in ("host", "nsub", "rank"):
        assert np.array_equal(two[k], one[k]), k
    for ext in ("AHF_halos", "AHF_profiles", "AHF_substructure", "AHF_particles"):
        a = open(os.path.join(out_dir, "one.z0.000." + ext), "rb").read(); b = open(os.path.join(out_dir, "two.z0.000." + ext), "rb").read()
        assert len(a) > 0 and a == b, ext
    # a halo served twice, or not at all, is refused
    try:
        multigpu.catalogue_from_ranks(None, par, seeds, parts + [parts[0]], ids); raise SystemExit("duplicate halo accepted")
    except ahf.AhfGpuError:
        pass
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_catalogue_gloo(tmp_path):
    """N > 1 host path on the CPU (world_size 2, gloo): both ranks derive the same halo seeds, each serves the haloes assigned to it, the
    results are gathered on rank 0 (multigpu.gather_parts_torch) and written by multigpu.catalogue_from_ranks -- the four catalogue files
    and the re-hashed links must equal those of one process serving every halo."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker_cat.py"
    script.write_text(_WORKER_CAT % dict(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_table_driven_hilbert_index_equals_the_step_form(tmp_path, golden):
    """hilbert.cuh is host+device code: the three-levels-per-look-up form the key kernel uses (hilbert_index21_tab) against the
    21-step form, compiled for the host, on random cells, the box corners and the golden particle keys of the reference."""
    src = tmp_path / "hil.cpp"
    src.write_text('''
#include "ahf_b200/csrc/hilbert.cuh"
#include <cstdio>
#include <cstdlib>
extern "C" int hil_check(const float *pos, const unsigned long long *keys, long n) {
  static uint16_t tab[12 * 512];
  ahf::hilbert_build_tab3(tab);
  unsigned long long bad = 0;
  srand(1);
  for (int i = 0; i < 1000000; i++) {
    uint32_t x = rand() & 0x1fffff, y = rand() & 0x1fffff, z = rand() & 0x1fffff;
    if (i < 8) { x = (i & 1) ? 0x1fffff : 0; y = (i & 2) ? 0x1fffff : 0; z = (i & 4) ? 0x1fffff : 0; }
    if (ahf::hilbert_index(x, y, z, 21) != ahf::hilbert_index21_tab(x, y, z, tab)) bad++;
  }
  for (long i = 0; i < n; i++) if (ahf::hilbert_key_pos21_tab(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], tab) != keys[i]) bad++;
  return (int)(bad > 1000000 ? 1000000 : bad);
}
''')
    so = tmp_path / "hil.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-I", ROOT, "-o", str(so), str(src)])
    L = C.CDLL(str(so))
    L.hil_check.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
    pos = np.ascontiguousarray(golden.pos, np.float32); keys = np.ascontiguousarray(golden.keys, np.uint64)
    assert L.hil_check(pos.ctypes.data, keys.ctypes.data, len(keys)) == 0


def test_patch_labels_as_union_find_equal_the_sweep(tmp_path, golden):
    """patches.cuh is host+device code: the union-find the device kernels run (edges to the VISIBLE face neighbours EARLIER in
    traversal order, larger root hooked under the smaller, patches numbered by the rank of their first cell) gives the labels of
    the reference's sequential colouring sweep -- in traversal order and with the cells visited in a scrambled order."""
    from oracle import oracle as O
    src = tmp_path / "uf.cpp"
    src.write_text('''
#include "ahf_b200/csrc/patches.cuh"
#include <vector>
extern "C" long uf_labels(const long long *nb6, long ncell, int *iso, int seed) {
  std::vector<int32_t> parent(ncell);
  std::vector<long> ord(ncell);
  for (long c = 0; c < ncell; c++) { parent[c] = (int32_t)c; ord[c] = c; }
  unsigned long long s = (unsigned long long)seed;
  if (seed) for (long c = ncell - 1; c > 0; c--) { s = s * 6364136223846793005ull + 1442695040888963407ull; long j = (long)((s >> 33) % (unsigned long long)(c + 1)); long t = ord[c]; ord[c] = ord[j]; ord[j] = t; }
  for (long q = 0; q < ncell; q++) {
    const long c = ord[q];
    for (int d = 0; d < 6; d++) { const long long n = nb6[6 * c + d]; if (n >= 0 && n < c) ahf::uf_unite(parent.data(), (int32_t)c, (int32_t)n); }
  }
  std::vector<int> rank(ncell, -1);
  long niso = 0;
  for (long c = 0; c < ncell; c++) if (ahf::uf_find(parent.data(), (int32_t)c) == c) rank[c] = (int)niso++;
  for (long c = 0; c < ncell; c++) iso[c] = rank[ahf::uf_find(parent.data(), (int32_t)c)];
  return niso;
}
''')
    so = tmp_path / "uf.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-I", ROOT, "-o", str(so), str(src)])
    L = C.CDLL(str(so))
    L.uf_labels.restype = C.c_long
    L.uf_labels.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_int]
    H = O.build_hierarchy(golden.pos, golden.n1d, nth_dom=golden.nper_dom, nth_ref=golden.nper_ref, patches=True)
    for lv in H[1:]:
        for seed in (0, 7):
            iso = np.empty(lv.ncell, np.int32)
            n = L.uf_labels(lv.nb6.ctypes.data, lv.ncell, iso.ctypes.data, seed)
            assert n == lv.iso_periodic.shape[0] and np.array_equal(iso, lv.iso)


def test_union_find_primitives_on_random_graphs(tmp_path):
    """patches.cuh (host build) against scipy's connected components on random graphs: same partition, and with every union hooking
    the larger root under the smaller one the root of a component is its smallest member (what the patch numbering relies on)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    src = tmp_path / "ufg.cpp"
    src.write_text('''
#include "ahf_b200/csrc/patches.cuh"
extern "C" void uf_graph(const int *a, const int *b, long ne, int *parent, long n) {
  for (long i = 0; i < n; i++) parent[i] = (int)i;
  for (long e = 0; e < ne; e++) ahf::uf_unite(parent, a[e], b[e]);
  for (long i = 0; i < n; i++) parent[i] = ahf::uf_find(parent, (int32_t)i);
}
''')
    so = tmp_path / "ufg.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-I", ROOT, "-o", str(so), str(src)])
    L = C.CDLL(str(so))
    L.uf_graph.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long]
    rng = np.random.default_rng(3)
    for n, ne in ((1, 0), (2, 1), (50, 20), (1000, 700), (20000, 15000), (20000, 60000)):
        a = rng.integers(0, n, ne).astype(np.int32); b = rng.integers(0, n, ne).astype(np.int32)
        root = np.empty(n, np.int32)
        L.uf_graph(a.ctypes.data, b.ctypes.data, ne, root.ctypes.data, n)
        ncomp, lab = connected_components(coo_matrix((np.ones(ne), (a, b)), shape=(n, n)), directed=False)
        assert len(np.unique(root)) == ncomp
        first = np.full(ncomp, n, np.int64)
        np.minimum.at(first, lab, np.arange(n))                  # smallest member of every scipy component
        assert np.array_equal(root, first[lab])


def test_committed_bench_lines_follow_the_contract():
    """the bench lines kept under profiles/ carry every key the measurement contract names (a guard for edits of bench.py)"""
    import glob
    import json
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r1[f-z]_bench_256_n1.json")))
    assert files
    for f in files:
        d = json.loads([l for l in open(f) if l.startswith("{")][0])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                  "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
            assert k in d, (f, k)
        assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and "workload" in d["config"] and "l2" in d["config"]
        assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
        r = d["roofline"]
        assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
        assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    full = json.load(open(os.path.join(ROOT, "profiles", "r1f_bench_256_n1.json")))
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(full["cpu_baseline"]) and full["cpu_baseline"]["kind"] == "reference"


def test_final_bench_line_and_traffic_source():
    """the last bench line of round 2 follows the contract as well, and `roofline.traffic` is read from a `k_deposit_dom` launch of the
    newest ncu summary that holds one (an earlier version picked the first launch of the alphabetically last file: a k_deposit_runs launch)"""
    import json
    import bench
    d = json.loads([l for l in open(os.path.join(ROOT, "profiles", "r3q_bench_256_n1.json")) if l.startswith("{")][0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["cpu_baseline"]["kind"] == "reference" and "256^3" in d["cpu_baseline"]["sample"] and "256^3" in d["config"]["workload"]
    assert d["e2e"]["h2d_bytes_per_step"] > 4e8 and d["e2e"]["d2h_bytes_per_step"] > 1e8          # members, profiles and the permutation come back
    t = bench.ncu_traffic_bytes(256)
    alg = 16.0 * 256 ** 3 + 4.0 * 256 ** 3
    assert t is not None and 1.0 * alg < t < 2.0 * alg, t                    # u64 accumulators: ~1.45x the algorithmic bytes
    assert d["roofline"]["traffic"] == t
    assert bench.ncu_traffic_bytes(128) is None


def test_min_ref_follows_reference(golden):
    """host arithmetic of ahf_gridinfo.c:147-175 (ahf.min_ref) against the first coloured level of the reference (tests/golden/patches.npz);
    frag16 sits on the edge: its level 3 has refine_ovdens = 199.99999999999997 against ovlim = 200"""
    from ahf_b200 import ahf
    par = ahf.params_from_reference(golden.glob, lgrid_dom=golden.n1d, nper_dom=golden.nper_dom, nper_ref=golden.nper_ref)
    l1dims = [int(golden.level(l)["l1dim"]) for l in range(golden.nlev)]
    medw = float(golden.weight.max()) if golden.weight is not None else 1.0
    assert ahf.min_ref(par, l1dims, medw) == golden.patches()[0]


def test_host_tree_and_halo_seeds_follow_reference(golden):
    """ahfgpu_tree_halos (host code of the library: analyseRef + spatialRef2halos on per-refinement tables) fed with the tables the
    oracle's restated RefCentre produces -- the same 18 columns ahfgpu_amr_patch_stats delivers from the device: substructure lists,
    main-branch daughters and closeRefDist equal the restated analyseRef (pinned on the reference's .AHF_gridtree), and the halo seeds
    (centre, gathering radius, particle count, order) equal the ones the reference hands to ahf_halos_sfc_constructHalo, bit for bit."""
    from ahf_b200 import ahf
    from oracle import oracle as O
    min_ref = golden.patches()[0]
    H = O.build_hierarchy(golden.pos, golden.n1d, nth_dom=golden.nper_dom, nth_ref=golden.nper_ref, patches=True)
    stats = [np.hstack([lv.patch, O.patch_extents(lv).reshape(len(lv.patch), 6)]) for lv in H[min_ref:]]
    out = ahf.tree_halos(stats, 3.0 / float(golden.d["boxsize"]))
    tree = O.patch_tree(H[min_ref:])
    for i, t in enumerate(tree):
        assert out["sub"][i] == t["sub"], i
        assert np.array_equal(out["daughter"][i], t["daughter"]) and np.array_equal(out["close"][i], t["close"]), i
    hs = golden.hs
    assert len(out["npart"]) == len(hs) and np.array_equal(out["npart"], hs[:, 4].astype(np.int64))
    assert np.array_equal(out["pos"], hs[:, 0:3]) and np.array_equal(out["gather_rad"], hs[:, 3])
    _, _, _, host = O.tree_to_halos(H[min_ref:], tree, 3.0 / float(golden.d["boxsize"]))
    assert np.array_equal(out["host"], host)


# ---- NEXT-3: sub-halo re-hash, ordering, catalogue writers ------------------------------------------------------------------------------
def _catalogue_case(kind):
    from ahf_b200 import synth
    if kind == "bench64":
        return synth.make_box(64, seed=43), None, False
    if kind == "clumps64":
        return synth.make_box(64, seed=7, n_clumps=40, clump_frac=0.5), None, False
    if kind == "host_a":
        return synth.make_host_box(60000, n_sub=8, n1d_bg=32, seed=47), 64, False
    if kind == "host_b":
        return synth.make_host_box(30000, n_sub=12, n1d_bg=32, seed=48), 64, False
    if kind == "species":
        return synth.make_species_box(32, seed=42), None, True
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["bench64", "clumps64", "host_a", "host_b", "species"])
def test_catalogue_writer_equals_reference_files(kind, tmp_path):
    """ahfgpu_catalogue_write (host code of the library) fed with what the UNMODIFIED reference held when its halo loop had finished --
    halo scalars, member lists, profiles, and hostHalo / hostHaloLevel / subStruct[] as its tree left them (oracle/ref_hooks.c) -- must
    reproduce the re-hashed host links and substructure lists, the halo order, and the four catalogue files byte for byte."""
    from ahf_b200 import ahf, synth
    from oracle import oracle as O
    box, lgrid, mm = _catalogue_case(kind)
    exe = O.REF_BIN_MM if mm else O.REF_BIN
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    w = str(tmp_path)
    inp = synth.write_reference_case_species(box, w) if mm else synth.write_reference_case(box, w, lgrid_domain=lgrid)
    d = os.path.join(w, "dump")
    O.run_reference(inp, dump_dir=d, multimass=mm)
    P = O.read_particles(os.path.join(d, "particles.bin"), multimass=mm)
    H = O.read_halos(d)
    T = O.read_halo_tree(d)
    g = H.glob
    fac = dict(r_fac=g[0], x_fac=g[1], v_fac=g[2], m_fac=g[3], rho_fac=g[4], phi_fac=g[5], rho_vir=g[8], u_fac=g[13], pmass=g[14])
    assert abs(g[15]) < 1e-12                                  # z = 0: the reference's file prefix is <prefix>.z0.000
    profiles = [H.prof[i] if H.s[i, 9] >= g[9] else None for i in range(H.n)]
    psp = None if not mm else [H.prof_species[i] if H.s[i, 9] >= g[9] else None for i in range(H.n)]
    out = ahf.catalogue_write(os.path.join(w, "own.z0.000"), H.s, H.s[:, 0:3].copy(), H.members, profiles, T["host_pre"], T["level_pre"], T["sub_pre"],
                              P.ids, fac, int(g[9]), part_u=P.u if mm else None, species=H.species if mm else None, prof_species=psp)
    assert np.array_equal(out["host"], T["host_post"])
    assert np.array_equal(out["nsub"], [len(q) for q in T["sub_post"]])
    for ext in ("AHF_halos", "AHF_profiles", "AHF_substructure", "AHF_particles"):
        a = open(os.path.join(w, "ref.z0.000." + ext), "rb").read(); b = open(os.path.join(w, "own.z0.000." + ext), "rb").read()
        if a != b:
            la, lb = a.split(b"\n"), b.split(b"\n")
            k = next((q for q in range(min(len(la), len(lb))) if la[q] != lb[q]), min(len(la), len(lb)))
            raise AssertionError("%s differs at line %d:\n ref %r\n own %r" % (ext, k, la[k:k + 1], lb[k:k + 1]))
    if kind.startswith("host"):
        assert (T["host_post"] >= 0).sum() >= 5                # the case does exercise surviving sub-haloes
    if not mm:
        # the same links from the library's own tree (ahfgpu_tree_halos_ex on the per-refinement tables): hostHalo, hostHaloLevel and the
        # substructure lists of halos[] as spatialRef2halos leaves them
        n1d = lgrid or box.n1d
        par = ahf.params_from_reference(g, lgrid_dom=n1d)
        Hh = O.build_hierarchy(P.pos, n1d, patches=True)
        m = ahf.min_ref(par, [lv.l1dim for lv in Hh])
        stats = [np.hstack([lv.patch, O.patch_extents(lv).reshape(len(lv.patch), 6)]) for lv in Hh[m:]]
        tr = ahf.tree_halos(stats, g[11] / g[1])
        assert np.array_equal(tr["host"], T["host_pre"]) and np.array_equal(tr["host_level"], T["level_pre"])
        assert all(np.array_equal(a, b) for a, b in zip(tr["halo_sub"], T["sub_pre"]))
        assert np.array_equal(tr["pos"], H.s[:, 0:3]) and np.array_equal(tr["npart"], H.s[:, 4].astype(np.int64))


def test_tree_cell_lists_equal_the_all_pairs_loops():
    """The library's tree replaces the reference's all-pairs loops (children x parents of consecutive levels, ahf_halos.c:1693-1800; adoption
    of parent-less refinements, :2030-2165; every halo against every halo for the gathering radius, :2985-3052) by searches through periodic
    cell lists.  On random tables that are large enough to take those paths (>= 512 refinements above an orphan, thousands of haloes, periodic
    extents, equal particle numbers, coincident centres) the results must equal a direct numpy restatement of the loops."""
    from ahf_b200 import ahf
    rng = np.random.default_rng(7)

    def level(n, half, frac_periodic=0.05):
        c = rng.random((n, 3))
        st = np.zeros((n, 18))
        st[:, 0] = rng.integers(8, 500, n); st[:, 1] = rng.integers(1, 60, n)          # few distinct particle numbers: many ties
        st[:, 2:5] = c; st[:, 9:12] = c; st[:, 6:9] = c
        lo, hi = c - half * rng.random((n, 3)), c + half * rng.random((n, 3))
        per = rng.random(n) < frac_periodic                                                # extents across a periodic face: max < min
        lo[per] = np.mod(lo[per], 1.0); hi[per] = np.mod(hi[per], 1.0)
        lo[~per] = np.clip(lo[~per], 0.0, 1.0); hi[~per] = np.clip(hi[~per], 0.0, 1.0)
        st[:, 12:18:2] = lo; st[:, 13:18:2] = hi
        return st
    stats = [level(1500, 0.03), level(2500, 0.01), level(1200, 0.004)]
    stats[1][:40, 9:12] = stats[1][40:80, 9:12]                                          # coincident centres
    out = ahf.tree_halos(stats, 0.2)

    def inside(v, lo, hi):
        return np.where(lo < hi, (v > lo) & (v < hi), ((v >= 0) & (v < hi)) | ((v > lo) & (v <= 1.0)))

    def pd2(a, b):
        d = np.abs(a - b); d = np.where(d > 0.5, 1.0 - d, d)
        return d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]
    # (1) lists before the several-parents / adoption steps cannot be read back; the final lists can: rebuild all steps directly
    cd = [s[:, 9:12].copy() for s in stats]
    sub = []; par = []
    for i in range(len(stats)):
        par.append([[] for _ in range(len(stats[i]))])
    for i in range(len(stats) - 1):
        P, Cc = stats[i], cd[i + 1]
        lst = []
        for j in range(len(P)):
            m = inside(Cc[:, 0], P[j, 12], P[j, 13]) & inside(Cc[:, 1], P[j, 14], P[j, 15]) & inside(Cc[:, 2], P[j, 16], P[j, 17])
            ks = np.nonzero(m)[0].tolist(); lst.append(ks)
            for k in ks:
                par[i + 1][k].append(j)
        sub.append(lst)
    sub.append([[] for _ in range(len(stats[-1]))])
    for i in range(1, len(stats) - 1):                                                       # several parents: keep the closest (first minimum)
        for j in range(len(stats[i])):
            if len(par[i][j]) > 1:
                d = pd2(cd[i][j][None, :], cd[i - 1][par[i][j]]); best = par[i][j][int(np.argmin(d))]
                for q in par[i][j]:
                    if q != best:
                        sub[i - 1][q] = [t for t in sub[i - 1][q] if t != j]
                par[i][j] = [best]
    n_orphans = 0
    for i in range(1, len(stats)):                                                           # adoption by the closest refinement above
        up = cd[i - 1].copy()
        for j in range(len(stats[i])):
            if not par[i][j]:
                best = int(np.argmin(pd2(cd[i][j][None, :], up))); n_orphans += 1
                par[i][j] = [best]; sub[i - 1][best].append(j); cd[i][j] = cd[i - 1][best]
    assert n_orphans > 100 and len(stats[0]) >= 512
    for i in range(len(stats)):
        assert [list(map(int, q)) for q in out["sub"][i]] == sub[i], f"substructure lists of level {i}"
    # gathering radius: half the distance to the nearest halo with MORE particles, clipped
    pos, npart = out["pos"], out["npart"]
    nh = len(npart)
    assert nh > 2048
    g = np.empty(nh)
    for i in range(nh):
        m = npart > npart[i]
        g[i] = 0.5 * np.sqrt(pd2(pos[i][None, :], pos[m]).min()) if m.any() else 0.2
    # sub-haloes carry closeRefDist as lower bound: take it from the output's own lower clip by comparing only where the plain value rules
    lower = np.minimum(g, 0.2)
    assert np.all(out["gather_rad"] >= lower - 0.0) and np.all(out["gather_rad"] <= 0.2)
    plain = out["host"] < 0
    assert np.array_equal(out["gather_rad"][plain], lower[plain])
