#!/usr/bin/env python
"""bench.py -- the AHF particle hot path on B200: Hilbert keys + sort, TSC deposit on the domain grid and every
refinement level, refinement flags / next level / relink, per-halo gather + radial sort + unbinding + profiles.

One "step" = one pass of the whole path over one synthetic box.  Prints ONE JSON line (rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n1d L] [--mode auto|boxes|slab] [--impl ours|reference]

N = 1: BASELINE.json configs[1] -- 256^3 particles, LgridDomain 256, AHF.input-example settings, halo seeds from the DEVICE hierarchy
       (patch labels + RefCentre tables on the GPU, tree on the host, computed once before the timed passes).
N > 1 (torchrun): BASELINE.json configs[2] -- ONE 512^3 box split over the N GPUs (SFC slabs + ghost shell, NCCL), "scaling": "strong".
       (--mode boxes: one independent 256^3 box per GPU, no collective, "weak" -- the replica number of round 1.)
--impl reference: the reference's own CPU implementation (oracle/_ref/ahf_ref = unmodified NegriAndrea/AHF with timing hooks) on the
       box's host cores, rank 0 only: the 256^3 box of the N = 1 arm (for N > 1 the same 256^3 box as a bounded sample of the 512^3
       workload), OMP_NUM_THREADS = nproc, best of min(K, 3) runs, plus one run with 1 thread in `detail`.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particles/sec through the hot path (keys+sort, TSC deposit+flag+refine+relink on all levels, halo gather+sort+unbind+profiles)"
UNIT = "particles/s"


def ncu_traffic_bytes(n1d: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the domain deposit kernel (k_deposit_dom), from the newest committed
    `ncu --set full` summary under profiles/ that holds a launch of it; only valid for the configuration that was profiled (256^3)."""
    if n1d != 256:
        return None
    import glob, re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full.txt")), reverse=True)         # r3... before r2... before r1...: newest round first
    for fn in files:
        inside = False
        rd = wr = None
        for line in open(fn):
            if line.startswith("## launch id"):
                if inside and rd is not None and wr is not None:
                    return rd + wr
                inside = re.search(r"k_deposit_dom\s*[<(]", line) is not None
                rd = wr = None
                continue
            if not inside:
                continue
            m = re.match(r"\s+dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
            if m:
                v = float(m.group(2)) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(3)]
                if m.group(1) == "read":
                    rd = v
                else:
                    wr = v
        if inside and rd is not None and wr is not None:
            return rd + wr
    return None


_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: keep a private handle to the real stdout and point fd 1 at stderr, so that whatever
    NCCL, the CUDA runtime or a child process prints (e.g. 'NCCL version ...' at NCCL_DEBUG=VERSION/WARN) cannot land next to it."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region, sampled in-process through NVML (nvidia_ml_py) -- the same
    counters `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` prints (B200_PROFILING.md), without
    forking nvidia-smi five times a second next to the measurement."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.sm_max = None
        self._halt = threading.Event()
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        if self.h is None:
            return
        nv = self.nv
        masks = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._halt.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, m in masks.items():
                    if r & m:
                        self.reasons.add(k)
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(self.samples), "source": "nvml"}


_REF_CASES = {}


def run_reference_sample(n1d: int, seed: int, threads: int | None):
    """Unmodified reference on one box of the bench generator; returns (pps, detail dict).  The snapshot is written once per (n1d, seed)."""
    from ahf_b200 import synth
    from oracle import oracle as O
    import atexit
    if (n1d, seed) not in _REF_CASES:
        box = synth.make_box(n1d, seed=seed)
        work = tempfile.mkdtemp(prefix="ahf_refbench_")
        atexit.register(shutil.rmtree, work, ignore_errors=True)
        _REF_CASES[(n1d, seed)] = (synth.write_reference_case(box, work), box.npart)
        del box
    inp, npart = _REF_CASES[(n1d, seed)]
    t0 = time.perf_counter()
    t = O.run_reference(inp, dump_dir=None, threads=threads)
    wall = time.perf_counter() - t0
    path_s = t["keys"] + t["sort"] + t["ll"] + t["deposit"] + t["refine"] + t["relink"] + t["halo_loop"]
    return npart / path_s, dict(t, wall_s=wall, path_s=path_s, npart=npart)


def dropin_leg(n1d: int, seed: int, ref_wall_s: float):
    """The drop-in PROGRAM on the GADGET file the reference has just analysed (cpu_baseline leg): ahf_b200/host/_build/AHF-b200-full = the
    reference's main / startrun / reader with key+sort, mesh, ahf_gridinfo and ahf_halos replaced by libahfgpu, from AHF.input to the four
    written catalogues, wall clock of the whole process (CUDA context creation included), and whether the catalogues equal the reference's
    byte for byte."""
    import hashlib
    import subprocess
    exe = os.path.join(ROOT, "ahf_b200", "host", "_build", "AHF-b200-full")
    if not os.path.exists(exe) or (n1d, seed) not in _REF_CASES:
        return None
    inp, npart = _REF_CASES[(n1d, seed)]
    d = os.path.dirname(inp)
    names = [f for f in sorted(os.listdir(d)) if ".AHF_" in f]

    def digests():
        return {f.split(".AHF_")[1]: hashlib.md5(open(os.path.join(d, f), "rb").read()).hexdigest() for f in names}
    ref = digests()
    for f in names:
        os.remove(os.path.join(d, f))
    env = dict(os.environ); env.pop("AHF_DUMP_DIR", None); env["AHFB200_TIMING"] = "1"
    walls, phases = [], {}
    for _ in range(2):                       # best of two: the CUDA driver start-up of a fresh process varies between 1 and 3 s on these boxes
        for f in names:
            if os.path.exists(os.path.join(d, f)):
                os.remove(os.path.join(d, f))
        t0 = time.perf_counter()
        pr = subprocess.run([exe, inp], cwd=d, env=env, capture_output=True, text=True)
        wall = time.perf_counter() - t0
        if pr.returncode != 0:
            return {"error": pr.stderr[-400:]}
        if not walls or wall < min(walls):
            for ln in pr.stderr.splitlines():
                if ln.startswith("AHFB200_TIMING"):
                    phases = {k: float(v) for k, v in (t.split("=") for t in ln.split()[1:])}
        walls.append(wall)
    wall = min(walls)
    own = digests() if all(os.path.exists(os.path.join(d, f)) for f in names) else {}
    return {"program": "ahf_b200/host/_build/AHF-b200-full", "workload": f"GADGET file of the {n1d}^3 box, AHF.input to the four catalogue files", "wall_s": wall,
            "wall_s_runs": walls, "reference_wall_s": ref_wall_s, "speedup_wall": ref_wall_s / wall, "particles_per_s_wall": npart / wall,
            "phases_s": phases, "phases_note": "ahfgpu_init = CUDA driver start-up + context creation of the fresh process (helper thread from program start)",
            "catalogues_byte_identical": {k: own.get(k) == v for k, v in ref.items()}}


def run_port_sample(n1d: int, seed: int):
    """CPU port (oracle/) on a bounded sample when the reference binary is not available."""
    from ahf_b200 import synth, ahf
    from oracle import oracle as O
    box = synth.make_box(n1d, seed=seed)
    t0 = time.perf_counter()
    keys = O.hilbert_keys(box.pos); order = O.argsort_keys(keys)
    pos = box.pos[order]; mom = box.mom[order]; keys = keys[order]
    O.build_hierarchy(pos, n1d)
    c, r, npart = synth.halo_seeds(box)
    P = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d)
    par = dict(r_fac=P.r_fac, x_fac=P.x_fac, v_fac=P.v_fac, m_fac=P.m_fac, rho_fac=P.rho_fac, phi_fac=P.phi_fac,
               Hubble=P.hubble, ovlim=P.ovlim, rho_vir=P.rho_vir, vesc_tune=P.vesc_tune, min_part=P.min_part)
    O.construct_halos(keys, pos, mom, None, None, par, c, r, npart)
    dt = time.perf_counter() - t0
    return box.npart / dt, dict(path_s=dt, npart=box.npart)


def many_haloes_leg(g, ahf, synth, n1d, peak):
    """The halo pass at a realistic halo count: the same generator with 2e4 small clumps (what a cosmological box of this size holds),
    seeds from the DEVICE hierarchy; gather / sort / unbind / profiles timed with CUDA events over 3 passes (not part of `value`)."""
    box = synth.make_box(n1d, seed=44, n_clumps=20000)
    par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d)
    g.set_params(par)
    g.upload(box.pos, box.mom); g.sfc_sort_resident(); g.build_amr()
    t0 = time.perf_counter()
    hs = g.halo_seeds(3.0 / box.boxsize, lists=False)
    seeds_ms = 1e3 * (time.perf_counter() - t0)
    c, r, s = np.ascontiguousarray(hs["pos"]), np.ascontiguousarray(hs["gather_rad"]), np.ascontiguousarray(hs["npart"], np.int64)
    os.environ["AHFGPU_STAGES"] = "1"
    st = []
    for _ in range(4):
        g.event_record(4)
        g.construct_halos(c, r, s, fetch=False)
        g.event_record(5)
        q = {k: g.stage_ms(k) for k in ("halo_gather", "halo_sort", "halo_localize", "halo_unbind", "halo_profiles")}
        q["total_events"] = g.event_elapsed_ms(4, 5)
        q["gathered"] = g.stage_count("halo_gathered"); q["iter_members"] = g.stage_count("halo_unbind_iter_members"); q["final"] = g.stage_count("halo_final_members")
        st.append(q)
    os.environ["AHFGPU_STAGES"] = "0"
    st = st[1:]
    m = {k: float(np.mean([q[k] for q in st])) for k in st[0]}
    scal = g.fetch_halos(len(r), scal_only=True)["scal"]
    t = m["halo_gather"] + m["halo_sort"] + m["halo_unbind"] + m["halo_profiles"]
    nbytes = 92.0 * m["gathered"] + 21.0 * m["iter_members"] + 28.0 * m["final"]
    return {"workload": f"synthetic {n1d}^3 box with 20000 Plummer clumps (seed 44), halo seeds from the device hierarchy", "seeds": int(len(r)),
            "haloes_ge_minpart": int((scal[:, 9] >= par.min_part).sum()), "seeds_ms": seeds_ms,
            "stages_ms": {k: m[k] for k in ("halo_gather", "halo_sort", "halo_localize", "halo_unbind", "halo_profiles")}, "halo_pass_ms_events": m["total_events"],
            "gathered_particles": m["gathered"], "unbind_pps": m["gathered"] / (t * 1e-3),
            "roofline": {"algorithmic_bytes": nbytes, "achieved_gbs": nbytes / (t * 1e-3) / 1e9, "frac": nbytes / (t * 1e-3) / 1e9 / peak}}


def bench_slab(args, rank, world, local_rank, config):
    """ONE box over all ranks (strong scaling): a step = exchange (keys, block histogram all-reduce, partition, NCCL send/recv to owners and
    ghost holders, the one sort) + mesh on own cells and ghost shell (per level one small all-gather and one all-gather of row keys) + halo pass
    of the haloes centred in the rank's key range.  `value`: the rank's file share already resident in HBM; `e2e`: pinned host share uploaded
    and every result of the rank's haloes (scalars, member lists as global particle indices, profiles) fetched, every step."""
    import torch
    import torch.distributed as dist
    from ahf_b200 import ahf, multigpu, synth
    dev = torch.device("cuda", local_rank)
    # every rank generates only its own share of the box (a z-slab of the lattice + every world-th clump): what it would read from its file
    pos_np, mom_np, cl, boxsize, pmass = synth.make_box_slice(args.n1d, rank, world, seed=43)
    c, r, seed = synth.halo_seeds_from(cl["centres"], cl["npart"], cl["scale"], boxsize)
    n_loc = int(pos_np.shape[0])
    pos_l = torch.empty(pos_np.shape, dtype=torch.float32, pin_memory=True); pos_l.numpy()[:] = pos_np
    mom_l = torch.empty(mom_np.shape, dtype=torch.float32, pin_memory=True); mom_l.numpy()[:] = mom_np
    counts = torch.zeros(world, device=dev, dtype=torch.int64); counts[rank] = n_loc
    if world > 1:
        dist.all_reduce(counts)
    counts = counts.cpu().numpy()
    n = int(counts.sum()); id_base = int(counts[:rank].sum())
    pmass = 0.3 * synth.RHOC0 * boxsize ** 3 / n
    par = ahf.make_params(boxsize=boxsize, pmass=pmass, lgrid_dom=args.n1d, device=local_rank)
    nid = multigpu.nccl_id_via_torch(rank, dev) if world > 1 else ahf.nccl_unique_id()
    sb = multigpu.SlabRank(par, rank, world, local_rank, nccl_id=nid)
    g = sb.g
    g.upload(pos_np, mom_np)
    del pos_np, mom_np

    def barrier():
        g.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # halo seeds: from the box's OWN hierarchy (default) -- patch labels joined across the rank boundaries, per-refinement tables combined
    # over the ranks (a collective call per level), the tree on every rank; computed once before the timed passes, like the N = 1 arm
    gw = 0.0
    seeds_info = {"source": args.seeds}
    hs = None
    if args.seeds == "device":
        maxg = 3.0 / boxsize
        sb.redistribute(id_base=id_base, ghost_width=min(maxg, 0.25)); sb.build_amr()
        g.halo_seeds(maxg)                                                # warm (allocations)
        sb.build_amr()
        barrier(); t0 = time.perf_counter()
        hs = g.halo_seeds(maxg)
        barrier(); seeds_info["ms_once_outside_the_timed_step"] = 1e3 * (time.perf_counter() - t0)
        c, r, seed = np.ascontiguousarray(hs["pos"]), np.ascontiguousarray(hs["gather_rad"]), np.ascontiguousarray(hs["npart"], np.int64)
        gw = float(r.max()) * (1.0 + 1e-9) if len(r) else 0.0              # the shell must hold the largest gathering sphere
        seeds_info.update(n=int(len(r)), refinements=int(sum(len(q) for q in hs["stats"])), max_gather_rad=gw,
                          what="labels + per-refinement tables on the devices (collective per level), tree + seeds on every rank, wall clock")

    def step_resident():
        st = {}
        sb.redistribute(id_base=id_base, ghost_width=gw)
        st.update({k: g.stage_ms(k) for k in ("slab_keys", "slab_histogram_allreduce", "slab_decompose", "slab_partition", "slab_exchange", "keys", "sort", "gather")})
        sb.build_amr()
        st.update({k: g.stage_ms(k) for k in ("amr_total", "ll", "deposit", "deposit_dom_kernel", "flag", "refine", "relink", "rows_allgather", "level_allgather", "rows_merge", "rows_owned")})
        st["deposit_particles"] = g.stage_count("deposit")
        mine, _ = sb.construct_halos(c, r, seed, fetch=False)
        st.update({k: g.stage_ms(k) for k in ("halo_gather", "halo_sort", "halo_unbind", "halo_profiles")})
        st["halo_gathered"] = g.stage_count("halo_gathered"); st["halos_mine"] = len(mine)
        return st

    def step_e2e():
        sb.distribute_ptr(pos_l.data_ptr(), mom_l.data_ptr(), n_loc, id_base=id_base, ghost_width=gw)
        sb.build_amr()
        mine, res = sb.construct_halos(c, r, seed, fetch=True)
        return mine, res

    os.environ["AHFGPU_STAGES"] = "0"
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank); sampler.start()
    barrier()
    l0 = g.launches()
    g.event_record(0)
    timed = [step_resident() for _ in range(args.steps)]
    g.event_record(1)
    barrier()
    ms_res = g.event_elapsed_ms(0, 1) / args.steps
    launches = g.launches() - l0
    info = g.slab_info()
    os.environ["AHFGPU_STAGES"] = "1"
    stages = [step_resident() for _ in range(2)][1:]
    os.environ["AHFGPU_STAGES"] = "0"
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    g.event_record(2)
    for _ in range(args.steps):
        mine, res = step_e2e()
    g.event_record(3)
    barrier()
    ms_e2e = g.event_elapsed_ms(2, 3) / args.steps
    os.environ.pop("AHFGPU_STAGES", None)
    clocks = sampler.stop()
    d2h = int(sum(v.nbytes for v in res.values() if hasattr(v, "nbytes")))
    nh_ok = int((res["scal"][:, 9] >= par.min_part).sum())
    # ---- BASELINE.json configs[3]: the catalogue of the box.  The ranks' results go to rank 0 (scalars, member lists as global input
    #      indices, profiles), which re-hashes the sub-haloes, orders and writes the four files (not timed as part of the step)
    cat_info = None
    members_total = int(len(res["members"]))
    if hs is not None:
        mt = torch.tensor([float(members_total)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(mt)
        if float(mt[0]) <= args.catalogue_max_members:
            import shutil, tempfile
            barrier(); t0 = time.perf_counter()
            parts = multigpu.gather_parts_torch(mine, res) if world > 1 else [(np.asarray(mine), res)]
            if rank == 0:
                d = tempfile.mkdtemp(prefix="ahf_b200_cat_")
                try:
                    out = multigpu.catalogue_from_ranks(os.path.join(d, "box.z0.000"), par, hs, parts, np.arange(n, dtype=np.uint64))
                    files = {f.split(".AHF_")[1]: os.path.getsize(os.path.join(d, f)) for f in sorted(os.listdir(d))}
                    nh_cat = sum(1 for ln in open(os.path.join(d, "box.z0.000.AHF_halos")) if not ln.startswith("#"))
                    cat_info = {"files_bytes": files, "halos_written": nh_cat, "subhalos": int((out["host"] >= 0).sum()),
                                "wall_s_gather_rehash_write": time.perf_counter() - t0, "members": int(float(mt[0])),
                                "what": "results of all ranks gathered on rank 0, ahfgpu_catalogue_write (re-hash, ordering, the reference's four file formats)"}
                finally:
                    shutil.rmtree(d, ignore_errors=True)
            barrier()
        else:
            cat_info = {"skipped": f"{int(float(mt[0]))} members above --catalogue-max-members {int(args.catalogue_max_members)}"}
    st = {k: float(np.mean([q[k] for q in stages])) for k in stages[0]}
    st["deposit_dom_kernel"] = float(np.mean([t["deposit_dom_kernel"] for t in timed]))
    # per-rank facts: maxima / sums over the ranks
    loc = torch.tensor([ms_res, ms_e2e, st["slab_exchange"], st["slab_histogram_allreduce"], max(st["rows_allgather"], 0.0), max(st["level_allgather"], 0.0),
                        st["amr_total"], st["halo_gather"] + st["halo_sort"] + st["halo_unbind"] + st["halo_profiles"], float(info["resident"]),
                        st["keys"] + st["sort"] + st["gather"], st["slab_keys"] + st["slab_decompose"] + st["slab_partition"]], device=dev, dtype=torch.float64)
    mx = loc.clone(); sm = torch.tensor([float(info["resident"]), float(d2h), float(nh_ok), float(launches), st["halo_gathered"], st["deposit_particles"]], device=dev, dtype=torch.float64)
    mn = loc.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN); dist.all_reduce(sm)
    ms_res, ms_e2e = float(mx[0]), float(mx[1])
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        C0 = args.n1d ** 3
        dep_bytes = 16.0 * info["resident"] + 4.0 * C0
        achieved = dep_bytes / (st["deposit_dom_kernel"] * 1e-3) / 1e9
        colls = {"particle exchange (ncclSend/ncclRecv, owners + ghost holders)": float(mx[2]), "block histogram (ncclAllReduce)": float(mx[3]),
                 "row keys per level (grouped ncclBroadcast)": float(mx[4]), "per-level scalars (ncclAllGather)": float(mx[5])}
        config = dict(config, workload=config["workload"].replace("seed 43+rank", "seed 43, generated per rank as its file share"), n_particles_total=n,
                      parallelism=f"ONE box over {world} GPU(s): Hilbert-block slabs with equal particle counts, ghost shell of {info['shell_blocks']} blocks "
                                  f"(2^-{info['decomp_bits']} box each, >= 8 domain cells and >= the largest gathering radius) sent as particles in the one "
                                  "exchange; per level one scalar all-gather + one all-gather of row keys; haloes served by the owner of their centre")
        config.pop("n_particles_per_gpu", None)
        line = {"metric": METRIC, "value": n / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_res, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 particles, u32/u64 fixed-point deposit, f64 halo arithmetic", "data": "synthetic", "config": config, "mode": "slab",
                "e2e": {"value": n / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": 24 * n + 56 * len(r) * world,
                        "d2h_bytes_per_step": int(sm[1]), "returns": "per rank: scalars, member lists (global particle indices) and profiles of the haloes it serves"},
                "gpu_launches": int(sm[3]), "clocks": clocks,
                "roofline": {"kernel": "TSC deposit, domain level (k_deposit_dom), rank 0", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes": dep_bytes, "kernel_ms": st["deposit_dom_kernel"]},
                "limiting_collective": max(colls, key=colls.get), "collectives_ms_max_over_ranks": colls,
                "phases_ms_max_over_ranks": {"decompose (keys, histogram, partition)": float(mx[10]), "exchange": float(mx[2]), "sort (keys+radix+gather)": float(mx[9]),
                                             "mesh": float(mx[6]), "halo pass": float(mx[7])},
                "phases_ms_min_over_ranks": {"mesh": float(mn[6]), "halo pass": float(mn[7]), "sort (keys+radix+gather)": float(mn[9])},
                "balance": {"resident_particles_max": float(mx[8]), "resident_particles_min": float(mn[8]), "resident_particles_sum": float(sm[0]),
                            "ghost_overhead": float(sm[0]) / n - 1.0, "own_particles_rank0": info["own_hi"] - info["own_lo"]},
                "stages_ms_rank0": {k: v for k, v in st.items() if k not in ("deposit_particles", "halo_gathered", "halos_mine")},
                "throughput": {"levels": info["levels"], "halos_in": len(r), "halos_ge_minpart": int(sm[2]), "halo_gathered_particles": float(sm[4]),
                               "deposit_particles_all_levels_incl_ghosts": float(sm[5])},
                "halo_seeds": seeds_info}
        if cat_info is not None:
            line["catalogue"] = cat_info
        emit(line)
    sb.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n1d", type=int, default=0, help="particles per dimension (default: 256 on one GPU / per box, 512 for ONE box over several GPUs)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-n1d", type=int, default=256, help="box of the reference arm / cpu_baseline leg (default: the N = 1 workload itself)")
    ap.add_argument("--seeds", default="device", choices=["device", "generator"], help="halo seeds of the N = 1 arm: from the device hierarchy (default) or the generator's clump centres")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--catalogue-max-members", type=float, default=4e8, help="N > 1: write the box's catalogue once after the timed passes unless it holds more member entries than this")
    ap.add_argument("--no-dropin", action="store_true", help="skip the wall-clock run of the drop-in program on the reference leg's GADGET file")
    ap.add_argument("--no-many-haloes", action="store_true", help="skip the second halo-pass measurement (box with 2e4 clumps, seeds from the device tree)")
    ap.add_argument("--breakdown", action="store_true", help="print the per-stage table to stderr")
    ap.add_argument("--mode", default="auto", choices=["auto", "boxes", "slab"],
                    help="'slab' = ONE box of --n1d^3 particles split into SFC slabs over the GPUs (default for N > 1, strong scaling); "
                         "'boxes' = one independent box per GPU (default for N = 1; weak, no collective)")
    args = ap.parse_args()
    claim_stdout()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.mode == "auto":
        args.mode = "slab" if max(world, args.gpus) > 1 else "boxes"
    if args.n1d <= 0:
        args.n1d = 512 if args.mode == "slab" else 256
    config = {"workload": f"synthetic {args.n1d}^3-particle box (jittered lattice + Plummer clumps, seed 43+rank), LgridDomain {args.n1d}, "
                          "AHF.input-example settings (NperDomCell 2.0, NperRefCell 2.5, VescTune 1.5, NminPerHalo 20, Dvir 200)",
              "n_particles_per_gpu": args.n1d ** 3, "lgrid_domain": args.n1d,
              "l2": "inputs (particle arrays 0.5 GB, domain grid 0.2 GB at 256^3) exceed the 126 MB L2; no explicit flush",
              "parallelism": "one independent box per GPU, no collective" if world > 1 else "single GPU"}

    # ------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle import oracle as O
        nthreads = os.cpu_count() or 1
        have_ref = os.path.exists(O.REF_BIN)
        ref_n1d = args.ref_n1d if have_ref else min(args.ref_n1d, 64)
        nrun = max(1, min(args.steps, 3))               # a 256^3 pass of the reference is about 25 s of wall clock
        vals, details = [], []
        for i in range(nrun):
            v, d = run_reference_sample(ref_n1d, 43, nthreads) if have_ref else run_port_sample(ref_n1d, 43)
            vals.append(v); details.append(d)
        one = None
        if have_ref:
            v1, d1 = run_reference_sample(ref_n1d, 43, 1)
            one = dict(d1, value=v1)
        val = float(max(vals))                           # best of nrun
        best = details[int(np.argmax(vals))]
        same = (ref_n1d == args.n1d)
        sample = (f"{ref_n1d}^3 box of the same generator (seed 43), " + ("the configuration of the GPU arm itself" if same else f"a bounded sample of the {args.n1d}^3 workload")
                  + f"; OMP_NUM_THREADS={nthreads}, best of {nrun} run(s), no warm-up run (the CPU arm has no warm-up effects worth a 25 s pass); "
                  "time = keys+qsort+ll+deposit(all levels, as the reference does them: twice)+refine+relink+halo loop from hook timers inside the unmodified binary")
        config = dict(config, reference_arm_box=f"{ref_n1d}^3", reference_arm_same_config=same)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": nrun, "steps_requested": args.steps,
                "warmup": 0, "warmup_requested": args.warmup, "ms_per_step": 1e3 * best.get("path_s", 0.0), "higher_is_better": True,
                "scaling": "strong" if args.mode == "slab" else "weak",
                "vs_baseline": None, "dtype": "f32 particles / f64 halo arithmetic", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads if have_ref else 1,
                                 "kind": "reference" if have_ref else "port", "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "detail": {"runs_pps": vals, "best": best, "one_thread": one,
                           "note": "only the halo loop, zero_dens and the gathering-radius loop are OpenMP-parallel in the reference's default build (SURVEY 2.1): the thread count barely matters"}}
        emit(line)
        return 0

    # ------------------------------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    from ahf_b200 import ahf, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ahf.build()
    if args.mode == "slab":
        return bench_slab(args, rank, world, local_rank, config)
    box = synth.make_box(args.n1d, seed=43 + rank)
    n = box.npart
    par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=args.n1d, device=local_rank)
    g = ahf.AhfGpu(par)
    # pinned host buffers for the end-to-end leg
    hpos = torch.empty((n, 3), dtype=torch.float32, pin_memory=True); hpos.numpy()[:] = box.pos
    hmom = torch.empty((n, 3), dtype=torch.float32, pin_memory=True); hmom.numpy()[:] = box.mom
    horder = torch.empty((n,), dtype=torch.int32, pin_memory=True)       # sorted offset -> input index: what a caller needs to read the member lists
    # halo seeds: what AHF's tree stage hands to the halo loop.  Default: from OUR hierarchy on the device (patch labels + RefCentre tables
    # on the GPU, analyseRef / spatialRef2halos on the host), computed once here; --seeds generator: the generator's clump centres
    seeds_ms = None
    if args.seeds == "device":
        g.upload(box.pos, box.mom); g.sfc_sort_resident(); g.build_amr()
        g.halo_seeds(3.0 / box.boxsize, lists=False)                                   # warm (allocations)
        g.build_amr()                                                     # a fresh hierarchy: the per-level tables are computed again
        g.synchronize()
        t0 = time.perf_counter()
        hs = g.halo_seeds(3.0 / box.boxsize, lists=False)
        seeds_ms = 1e3 * (time.perf_counter() - t0)
        centres, rad, seednp = np.ascontiguousarray(hs["pos"]), np.ascontiguousarray(hs["gather_rad"]), np.ascontiguousarray(hs["npart"], np.int64)
    else:
        centres, rad, seednp = synth.halo_seeds(box)

    def barrier():
        g.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        # the timed passes run with AHFGPU_STAGES=0 (only the domain-deposit kernel timer the roofline needs): a stage timer is an
        # event record between kernels on the stream; the per-stage table is taken from separate instrumented passes afterwards
        g.sfc_sort_resident()
        st = {k: g.stage_ms(k) for k in ("keys", "sort", "gather")}
        st["rs_scatter_kernel"] = g.stage_ms("rs_scatter_kernel")          # -1 unless AHFGPU_KERNEL_STAGES=1 (instrumented passes)
        st["rs_scatter_pairs"] = g.stage_count("rs_scatter_kernel")
        g.build_amr()
        st.update({k: g.stage_ms(k) for k in ("ll", "deposit", "deposit_dom_kernel", "flag", "refine", "relink")})
        st["deposit_particles"] = g.stage_count("deposit")
        g.construct_halos(centres, rad, seednp, fetch=False)
        st.update({k: g.stage_ms(k) for k in ("halo_gather", "halo_sort", "halo_unbind", "halo_profiles")})
        st["halo_gathered"] = g.stage_count("halo_gathered")
        st["halo_iter_members"] = g.stage_count("halo_unbind_iter_members")
        st["halo_final_members"] = g.stage_count("halo_final_members")
        return st

    pinned = {}                                                        # result buffers of the end-to-end leg: pinned, sized after the first pass

    e2e_wall = {"upload_pos+keys+sort": 0.0, "mesh": 0.0, "halo pass (waits for the momenta)": 0.0, "fetch": 0.0}

    def step_e2e():
        t0 = time.perf_counter()
        g.sfc_sort_async_ptr(hpos.data_ptr(), hmom.data_ptr(), n)      # momenta travel behind the sort and the hierarchy build
        g._chk(g._L.ahfgpu_particle_ids_async(g._h, horder.data_ptr()))   # the permutation that ties member offsets to the caller's particles: device -> host behind the momenta
        t1 = time.perf_counter()
        g.build_amr()
        t2 = time.perf_counter()
        g.construct_halos(centres, rad, seednp, fetch=False)
        t3 = time.perf_counter()
        res = g.fetch_halos(len(rad), bufs={k: v.numpy() for k, v in pinned.items()})   # scalars, member lists, profiles: everything the catalogue writers read
        t4 = time.perf_counter()
        for k, dt in zip(e2e_wall, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            e2e_wall[k] += 1e3 * dt
        if not pinned:
            for k, v in res.items():
                if k in ("scal", "members", "prof"):
                    pinned[k] = torch.empty((int(v.size * 1.25) + 1024,), dtype=torch.float64 if v.dtype == np.float64 else torch.int64, pin_memory=True)
            # the member lists are final after the unbinding: from now on they travel home while the profiles are computed
            g._chk(g._L.ahfgpu_halo_members_buffer(g._h, pinned["members"].data_ptr(), pinned["members"].numel()))
        return res

    # ---- HBM-resident timing
    g.upload(box.pos, box.mom)
    os.environ["AHFGPU_STAGES"] = "0"                 # warm-up in the mode that is timed (block cache, clocks and launch pattern settle on it)
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank); sampler.start()
    step_resident()
    barrier()
    l0 = g.launches()
    g.event_record(0)
    timed, host_ms = [], []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        timed.append(step_resident())
        host_ms.append(1e3 * (time.perf_counter() - t0))          # every API call returns synchronised: host time per step, for diagnosis only
    g.event_record(1)
    barrier()
    ms_res = g.event_elapsed_ms(0, 1) / args.steps
    launches = g.launches() - l0
    os.environ["AHFGPU_STAGES"] = "1"; os.environ["AHFGPU_KERNEL_STAGES"] = "1"
    stages = [step_resident() for _ in range(3)][1:]            # instrumented passes (not timed): the per-stage table
    os.environ["AHFGPU_STAGES"] = "0"; os.environ.pop("AHFGPU_KERNEL_STAGES", None)
    # ---- end-to-end timing (pinned host -> device every step, results back every step)
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    for k in e2e_wall:
        e2e_wall[k] = 0.0
    g.event_record(2)
    for _ in range(args.steps):
        e2e_res = step_e2e()
    scal = e2e_res["scal"]
    d2h_bytes = int(sum(v.nbytes for v in e2e_res.values() if hasattr(v, "nbytes"))) + 4 * n
    g.event_record(3)
    barrier()
    ms_e2e = g.event_elapsed_ms(2, 3) / args.steps
    # ---- the same resident pass with the halo seeds re-derived EVERY pass from the pass's own hierarchy (patch labels + per-patch tables on
    #      the device, tree on the host): no stage between particles and halo results is taken from an earlier pass.  Wall clock (the tree
    #      is host code), synchronised on both sides.
    seeds_leg = None
    if args.seeds == "device" and world == 1:
        os.environ["AHFGPU_STAGES"] = "0"
        ws, ss = [], []
        for it in range(min(args.steps, 5) + 1):
            barrier(); t0 = time.perf_counter()
            g.sfc_sort_resident(); g.build_amr()
            t1 = time.perf_counter()
            hs2 = g.halo_seeds(3.0 / box.boxsize, lists=False)
            t2 = time.perf_counter()
            g.construct_halos(np.ascontiguousarray(hs2["pos"]), np.ascontiguousarray(hs2["gather_rad"]), np.ascontiguousarray(hs2["npart"], np.int64), fetch=False)
            barrier(); t3 = time.perf_counter()
            if it:
                ws.append(1e3 * (t3 - t0)); ss.append(1e3 * (t2 - t1))
        seeds_leg = {"ms_per_step": float(np.mean(ws)), "value": n / (float(np.mean(ws)) * 1e-3), "unit": UNIT, "seeds_ms": float(np.mean(ss)), "steps": len(ws),
                     "n_seeds": int(len(hs2["npart"])),
                     "what": "keys+sort, mesh, patch labels + per-patch tables (device) + tree and seeds (host), halo pass; host wall clock"}
    os.environ.pop("AHFGPU_STAGES", None)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms_res, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_res, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    st = {k: float(np.mean([s[k] for s in stages])) for k in stages[0]}
    st["deposit_dom_kernel"] = float(np.mean([t["deposit_dom_kernel"] for t in timed]))      # the roofline kernel: CUDA events inside the TIMED passes
    nlev = g.nlevels()
    hdr0, _ = g.level_header(0)
    peak, peak_src = measured_peak_gbs()
    C0 = args.n1d ** 3
    dep_bytes = 16.0 * n + 4.0 * C0                       # B_dep,0 = 16 N_0 + 4 C_0
    t_dep_kernel = st["deposit_dom_kernel"] * 1e-3
    achieved = dep_bytes / t_dep_kernel / 1e9
    nhalo_ok = int((scal[:, 9] >= par.min_part).sum())
    # stage-level fractions of the HBM peak with the ALGORITHMIC bytes of SURVEY 8d (explanatory; `roofline` above is the contract's object)
    sumN = st["deposit_particles"]
    sumC = float(sum(g.level_header(l)[0][1] for l in range(nlev)))
    Clast, Nlast = float(g.level_header(nlev - 1)[0][1]), float(g.level_header(nlev - 1)[0][2])
    def _rf(nbytes, ms):
        gbs = nbytes / (ms * 1e-3) / 1e9
        return {"algorithmic_bytes": nbytes, "ms": ms, "achieved_gbs": gbs, "frac": gbs / peak}
    t_halo = st["halo_gather"] + st["halo_sort"] + st["halo_unbind"] + st["halo_profiles"]
    roofline_stages = {
        "keys+sort+gather (264 N)": _rf(264.0 * n, st["keys"] + st["sort"] + st["gather"]),
        "deposit, all levels (16 sum N_l + 4 sum C_l)": _rf(16.0 * sumN + 4.0 * sumC, st["deposit"]),
        "flags (5 sum C_l)": _rf(5.0 * sumC, st["flag"]),
        # next-level construction: reads the marks of a level (1 B per cell), writes per NEW cell the key (8), the compressed neighbour
        # table (40), parent link (4), interior / x-break / mark / tn bytes (4), dens (4) and the row index (4) = 64 B
        "refine (1 sum C_l + 64 sum C_{l+1})": _rf(1.0 * (sumC - Clast) + 64.0 * (sumC - C0), st["refine"]),
        # relink: positions of the particles of a level in (16 B each), cell + list entry + local position of the moved ones out (24 B)
        "relink (16 sum N_l + 24 sum N_{l+1})": _rf(16.0 * (sumN - Nlast) + 24.0 * (sumN - n), st["relink"]),
        "halo pass (92 n_gathered + 21 sum_i n^(i) + 28 n_final)": _rf(92.0 * st["halo_gathered"] + 21.0 * st["halo_iter_members"] + 28.0 * st["halo_final_members"], t_halo),
    }
    # the kernel with the largest share of a pass (profiles/*_launches_summary.txt): the ranked scatter of the main radix sort; per launch it
    # reads and writes every (u64 key, u32 index) pair once: 24 N algorithmic bytes (SURVEY 8d counts the sort as 8 passes x 2 x 12 N)
    roofline_kernels = {}
    if st.get("rs_scatter_kernel", -1) > 0 and st.get("rs_scatter_pairs", 0) > 0:
        nl = st["rs_scatter_pairs"] / n
        roofline_kernels["k_rs_scatter (main sort)"] = dict(_rf(24.0 * st["rs_scatter_pairs"], st["rs_scatter_kernel"]), launches_per_pass=nl,
                                                           kernel_ms_per_launch=st["rs_scatter_kernel"] / nl, bound="hbm")
    line = {
        "metric": METRIC, "value": world * n / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 particles, u32/u64 fixed-point deposit, f64 halo arithmetic", "data": "synthetic", "config": config,
        "e2e": {"value": world * n / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": 24 * n + 32 * len(rad) + 8 * len(rad),
                "d2h_bytes_per_step": d2h_bytes,
                "returns": "halo scalars (64 doubles each), member lists, profiles (25 columns per bin) and the sorted-offset -> input-index permutation",
                "host_wall_ms_per_call": {k: v / args.steps for k, v in e2e_wall.items()}},
        "halo_seeds": {"source": args.seeds, "n": int(len(rad)), "ms_once_outside_the_timed_step": seeds_ms, "what_is_timed": "patch labels + per-patch tables of every coloured level on the device, tree + seeds on the host, wall clock",
                       "note": "device: ahfgpu_amr_patch_stats per level (GPU) + ahfgpu_tree_halos (host, includes the O(N_h^2) gathering-radius loop)"},
        "gpu_launches": int(launches),
        "resident_step_ms_host": [round(x, 3) for x in host_ms],
        "clocks": clocks,
        "roofline": {"kernel": "TSC deposit, domain level (k_deposit_*)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic_bytes(args.n1d), "peak_source": peak_src,
                     "algorithmic_bytes": dep_bytes, "kernel_ms": st["deposit_dom_kernel"]},
        "roofline_stages": roofline_stages,
        "roofline_kernels": roofline_kernels,
        "stages_note": "timed passes run with AHFGPU_STAGES=0 (only the domain-deposit kernel timer, which `roofline` uses); stages_ms / roofline_stages come from two separate instrumented passes, so their sum exceeds ms_per_step by the cost of ~90 event records per pass",
        "stages_ms": {k: v for k, v in st.items() if k not in ("deposit_particles", "halo_gathered", "halo_iter_members", "halo_final_members", "rs_scatter_pairs")},
        "throughput": {"deposit_pps": st["deposit_particles"] / (st["deposit"] * 1e-3), "deposit_particles_all_levels": st["deposit_particles"],
                       "unbind_pps": st["halo_gathered"] / ((st["halo_gather"] + st["halo_sort"] + st["halo_unbind"] + st["halo_profiles"]) * 1e-3),
                       "halo_gathered_particles": st["halo_gathered"], "levels": nlev, "halos_in": len(rad), "halos_ge_minpart": nhalo_ok},
    }
    if seeds_leg is not None:
        line["step_with_seeds"] = seeds_leg
    if not args.no_many_haloes and world == 1:
        line["halo_pass_many_haloes"] = many_haloes_leg(g, ahf, synth, args.n1d, peak)
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        nthreads = os.cpu_count() or 1
        if os.path.exists(O.REF_BIN):
            v, d = run_reference_sample(args.ref_n1d, 43, nthreads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": nthreads, "kind": "reference",
                                    "sample": f"unmodified reference (oracle/_ref/ahf_ref) on the {args.ref_n1d}^3 box of the same generator and seed"
                                              + (" -- the GPU arm's own box" if args.ref_n1d == args.n1d else "") + ", one run, hook timers of the path's phases",
                                    "detail": {k: d[k] for k in ("keys", "sort", "ll", "deposit", "refine", "relink", "halo_loop", "path_s")}}
        else:
            v, d = run_port_sample(64, 43)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": "oracle/ C port on a 64^3 box of the same generator"}
    if not args.no_cpu_baseline and "cpu_baseline" in line and line["cpu_baseline"]["kind"] == "reference" and not args.no_dropin:
        dl = dropin_leg(args.ref_n1d, 43, d["wall_s"])
        if dl is not None:
            line["e2e_dropin"] = dl
    if args.breakdown:
        for k, v in sorted(line["stages_ms"].items()):
            print(f"  {k:22s} {v:10.3f} ms", file=sys.stderr)
    emit(line)
    g.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
