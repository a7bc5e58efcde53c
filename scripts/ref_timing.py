"""Time the unmodified reference (oracle/_ref/ahf_ref) on the box's host cores: 128^3 and 256^3 boxes of the bench generator,
OMP_NUM_THREADS = 1 and nproc.  Prints one JSON object (kept under profiles/ as the CPU baseline of the round)."""
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from ahf_b200 import synth          # noqa: E402
from oracle import oracle as O      # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [128, 256]
    ncpu = os.cpu_count() or 1
    out = {"host_cores": ncpu, "runs": []}
    try:
        out["mem_gb"] = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2 ** 30
    except Exception:
        pass
    for n1d in sizes:
        box = synth.make_box(n1d, seed=43)
        work = tempfile.mkdtemp(prefix="ahf_reftime_")
        try:
            inp = synth.write_reference_case(box, work)
            for thr in (ncpu, 1):
                t0 = time.perf_counter()
                t = O.run_reference(inp, dump_dir=None, threads=thr)
                wall = time.perf_counter() - t0
                path_s = t["keys"] + t["sort"] + t["ll"] + t["deposit"] + t["refine"] + t["relink"] + t["halo_loop"]
                out["runs"].append(dict(n1d=n1d, threads=thr, wall_s=wall, path_s=path_s, pps=box.npart / path_s, timing=t))
                print(json.dumps(out["runs"][-1]), file=sys.stderr, flush=True)
        finally:
            shutil.rmtree(work, ignore_errors=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
