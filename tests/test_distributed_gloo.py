"""The N>1 path on CPU: world_size-2 gloo.  Each rank owns a contiguous Morton range of the
coarse cells (the reference's p4est ownership, ms.tpp:52), computes its cells' (M, b) -- with
the CPU oracle standing in for the GPU stage, which tests may do -- and the ranks all_gather
the coarse contributions exactly as bench.py / the driver do over NCCL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, r, l, outdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mpi_parallel_multiscale_diffusion_fem_b200 import parallel, coarse_corners
    from oracle import oracle as O
    total = (1 << r) ** 2
    lo, hi = parallel.shard_range(total)
    cor = coarse_corners(r, lo, hi)
    res = O.run_cells(l, cor, O.coeff(O.COEFF_REFERENCE), keep_phi=False)
    M, b = parallel.gather_coarse_contributions(res["M"], res["b"], total)
    t = parallel.max_over_ranks(float(rank + 1))
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), M=M.numpy(), b=b.numpy(), lo=lo, hi=hi, t=t)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_and_gather(tmp_path, oracle):
    r, l, world = 2, 4, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, r, l, str(tmp_path)), nprocs=world, join=True)
    from mpi_parallel_multiscale_diffusion_fem_b200 import coarse_corners
    ref = oracle.run_cells(l, coarse_corners(r), oracle.coeff(oracle.COEFF_REFERENCE), keep_phi=False)
    got = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % k)) for k in range(world)]
    assert (int(got[0]["lo"]), int(got[0]["hi"])) == (0, 8) and (int(got[1]["lo"]), int(got[1]["hi"])) == (8, 16)
    for g in got:
        assert np.array_equal(g["M"], ref["M"]) and np.array_equal(g["b"], ref["b"])
        assert float(g["t"]) == 2.0     # max over ranks


def test_uneven_partition_covers_everything():
    from mpi_parallel_multiscale_diffusion_fem_b200 import morton_partition
    for total in (1, 7, 64, 65536):
        for world in (1, 2, 3, 4, 8):
            edges = [morton_partition(total, q, world) for q in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
