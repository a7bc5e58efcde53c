"""Run under torchrun on >= 2 GPUs: one box split into SFC slabs over the ranks (ahf_b200/multigpu.py) must give exactly the
single-GPU result: same levels, bit-identical densities and refinement marks, same halo table."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from ahf_b200 import ahf, multigpu, synth   # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    n1d = int(os.environ.get("AHF_MGPU_N1D", "64"))
    box = synth.make_box(n1d, seed=17, n_clumps=10)
    n = box.npart
    b = (np.arange(world + 1) * n) // world
    par = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d, device=lr)
    sb = multigpu.SlabBox(par, rank, world, lr)
    sb.distribute(box.pos[b[rank]:b[rank + 1]], box.mom[b[rank]:b[rank + 1]])       # each rank "reads" a file-order slice
    nl = sb.build_amr()
    ntot = sb.gather_box()
    assert ntot == n, (ntot, n)
    c, r, seed = synth.halo_seeds(box)
    scal = sb.construct_halos(c, r, seed)
    # single-GPU truth on every rank
    par1 = ahf.make_params(boxsize=box.boxsize, pmass=box.pmass, lgrid_dom=n1d, device=lr)
    with ahf.AhfGpu(par1) as g:
        keys, _ = g.sfc_sort(box.pos, box.mom)
        nl1 = g.build_amr()
        assert nl == nl1, (nl, nl1)
        dep = torch.zeros(nl, dtype=torch.int64, device="cuda")
        for l in range(nl):
            A, B = sb.g.level(l), g.level(l)
            assert A.ncell == B.ncell and np.array_equal(A.lin(), B.lin()), "level %d cells" % l
            assert np.array_equal(A.dens, B.dens), "level %d densities are not bit identical" % l
            assert np.array_equal(A.mark, B.mark) and np.array_equal(A.runflags, B.runflags)
            dep[l] = A.npart_dep
        dist.all_reduce(dep)
        assert [int(v) for v in dep] == [g.level_header(l)[0][2] for l in range(nl)], "particles per level"
        res1 = g.construct_halos(c, r, seed)
        assert np.array_equal(scal[:, 5:10], res1["scal"][:, 5:10])
        assert np.allclose(scal, res1["scal"], rtol=1e-12, atol=0, equal_nan=True)
        for k, h in enumerate(sb.local_halos):
            assert np.array_equal(sb.gh.halo_members(sb.local_result, k), g.halo_members(res1, h))
    sb.close()
    dist.barrier()
    if rank == 0:
        print("MGPU_OK world=%d levels=%d haloes=%d" % (world, nl, len(r)))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
