"""Run under torchrun on >= 2 GPUs: one box split into SFC slabs over the ranks with the NCCL transport (ahf_b200/csrc/comm.cu, slab.cu)
must give exactly the single-GPU result (tests/slab_util.py): every rank checks its own part against the truth it computes itself, the
partition property is checked on rank 0 from the gathered reports."""
import os
import pickle
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ahf_b200 import ahf, multigpu, synth   # noqa: E402
import slab_util                             # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    n1d = int(os.environ.get("AHF_MGPU_N1D", "64"))
    cb = np.array([[0.5, 0.5, 0.5], [0.5, 0.25, 0.75], [0.999, 0.5, 0.3], [0.001, 0.002, 0.998], [0.25, 0.5, 0.5]])
    box = synth.make_box(n1d, seed=17, n_clumps=14, centres_box=cb)
    n = box.npart
    c, r, seed = synth.halo_seeds(box)
    b = (np.arange(world + 1) * n) // world
    par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d, device=lr)
    sb = multigpu.SlabRank(par, rank, world, lr, nccl_id=multigpu.nccl_id_via_torch(rank, torch.device("cuda", lr)))
    sb.distribute(box.pos[b[rank]:b[rank + 1]], box.mom[b[rank]:b[rank + 1]], id_base=int(b[rank]))
    rep = slab_util.rank_report(sb, c, r, seed)
    # seeds from the split box's own hierarchy (labels joined across the ranks over NCCL), halo pass on the owners, catalogue on rank 0
    maxg = 3.0 / box.boxsize
    sb.distribute(box.pos[b[rank]:b[rank + 1]], box.mom[b[rank]:b[rank + 1]], id_base=int(b[rank]), ghost_width=min(maxg, 0.25))
    sb.build_amr()
    sd = sb.g.halo_seeds(maxg)
    mine, res = sb.construct_halos(np.ascontiguousarray(sd["pos"]), np.ascontiguousarray(sd["gather_rad"]), np.ascontiguousarray(sd["npart"], np.int64))
    parts = multigpu.gather_parts_torch(mine, res)
    sb.close()
    blobs = [None] * world
    dist.all_gather_object(blobs, pickle.dumps(rep))
    if rank == 0:
        import tempfile
        T = slab_util.single_gpu_truth(ahf, box, n1d, c, r, seed, device=lr)
        out = slab_util.check_against_truth([pickle.loads(x) for x in blobs], T, n)
        with tempfile.TemporaryDirectory() as d, ahf.AhfGpu(par) as g:
            keys, order = g.sfc_sort(box.pos, box.mom); g.build_amr()
            ref = g.halo_seeds(maxg)
            for k in ("pos", "npart", "host", "host_level"):
                assert np.array_equal(sd[k], ref[k]), k
            assert np.allclose(sd["gather_rad"], ref["gather_rad"], rtol=1e-9, atol=0)
            r1 = dict(g.construct_halos(np.ascontiguousarray(ref["pos"]), np.ascontiguousarray(ref["gather_rad"]), np.ascontiguousarray(ref["npart"], np.int64)))
            r1["members"] = order.astype(np.int64)[r1["members"]]
            multigpu.catalogue_from_ranks(os.path.join(d, "one.z0.000"), par, ref, [(np.arange(len(ref["npart"])), r1)], box.ids)
            multigpu.catalogue_from_ranks(os.path.join(d, "split.z0.000"), par, sd, parts, box.ids)
            for ext in ("AHF_halos", "AHF_profiles", "AHF_substructure", "AHF_particles"):
                assert open(os.path.join(d, "one.z0.000." + ext), "rb").read() == open(os.path.join(d, "split.z0.000." + ext), "rb").read(), ext
            out["seeds"] = len(ref["npart"])
        print("MGPU_OK world=%d %s" % (world, out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
