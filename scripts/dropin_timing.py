"""Wall clock of the DROP-IN program against the reference program on the same GADGET snapshot (BASELINE.json configs[0], [1]):
ahf_b200/host/_build/AHF-b200 (the reference's own main / readers / ahf_gridinfo / tree / writers + libahfgpu.so for keys+sort, the
mesh and the halo loop) vs oracle/_ref/ahf_ref (the unmodified CPU reference), both from `AHF.input` to the written catalogues, plus
a check that the catalogues agree.  Prints one JSON object (kept under profiles/)."""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from ahf_b200 import synth          # noqa: E402
from oracle import oracle as O      # noqa: E402

DROPIN = os.path.join(ROOT, "ahf_b200", "host", "_build", "AHF-b200")
DROPIN_FULL = os.path.join(ROOT, "ahf_b200", "host", "_build", "AHF-b200-full")     # ahf_gridinfo + ahf_halos replaced as well (NEXT-1/2/3)


def run(exe, inp, cwd, threads):
    env = dict(os.environ); env.pop("AHF_DUMP_DIR", None); env["OMP_NUM_THREADS"] = str(threads); env["AHFB200_TIMING"] = "1"
    t0 = time.perf_counter()
    pr = subprocess.run([exe, inp], cwd=cwd, env=env, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if pr.returncode != 0:
        raise RuntimeError(pr.stderr[-2000:])
    t = {}
    for line in pr.stderr.splitlines():
        if line.startswith("REFHOOK_TIMING") or line.startswith("AHFB200_TIMING"):
            for tok in line.split()[1:]:
                k, v = tok.split("="); t[k] = float(v)
    return dt, t


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [128, 256]
    ncpu = os.cpu_count() or 1
    out = {"host_cores": ncpu, "runs": []}
    for n1d in sizes:
        box = synth.make_box(n1d, seed=43)
        work = tempfile.mkdtemp(prefix="ahf_dropin_time_")
        try:
            res = {"n1d": n1d, "particles": box.npart}
            dirs = {}
            for tag, exe in (("reference", O.REF_BIN), ("dropin", DROPIN), ("dropin_full", DROPIN_FULL)):
                d = os.path.join(work, tag); dirs[tag] = d
                inp = synth.write_reference_case(box, d)
                walls = []
                for rep in range(2):
                    for f in os.listdir(d):
                        if ".AHF_" in f or f.endswith(".log") or f.endswith(".parameter"):
                            os.remove(os.path.join(d, f))
                    w, t = run(exe, inp, d, ncpu)
                    walls.append(w)
                res[tag] = {"wall_s": walls, "best_s": min(walls), "hook_timing": t}
            pre = "ref.z0.000.AHF_"
            for tag in ("dropin", "dropin_full"):
                same = {}
                for f in ("particles", "substructure", "halos", "profiles"):
                    a = open(os.path.join(dirs["reference"], pre + f)).read(); b = open(os.path.join(dirs[tag], pre + f)).read()
                    same[f] = (a == b)
                res["catalogues_byte_identical" + ("" if tag == "dropin" else "_full")] = same
            res["halos"] = sum(1 for line in open(os.path.join(dirs["dropin"], pre + "halos")) if not line.startswith("#"))
            res["speedup_wall"] = res["reference"]["best_s"] / res["dropin"]["best_s"]
            res["speedup_wall_full"] = res["reference"]["best_s"] / res["dropin_full"]["best_s"]
            out["runs"].append(res)
            print(json.dumps(res), file=sys.stderr, flush=True)
        finally:
            shutil.rmtree(work, ignore_errors=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
