"""Multi-GPU plumbing: one process per GPU, torch.distributed for the only exchange the
path has -- gathering the per-cell coarse contributions (M, b), which is what the
reference's compress(VectorOperation::add) does for the coarse system after the basis
stage (/root/reference/include/base/diffusion_problem_ms.tpp:253-254).  The basis stage
itself has no collective: coarse cells are partitioned in contiguous Morton ranges
(ms.tpp:52) and every rank solves its own cells."""
import torch
import torch.distributed as dist

from .coarse import morton_partition


def shard_range(n_cells, rank=None, world=None):
    world = dist.get_world_size() if world is None else world
    rank = dist.get_rank() if rank is None else rank
    return morton_partition(n_cells, rank, world)


def gather_coarse_contributions(M_local, b_local, n_cells_total, device=None):
    """all_gather of the [n_local, nb*nb + nb] packed (M | b) blocks of every rank (nb = 2^dim
    bases per cell); returns (M [C,nb,nb], b [C,nb]) in Morton order on every rank.  NCCL with
    CUDA tensors on the GPU box, gloo with CPU tensors in the CPU tests."""
    world, rank = dist.get_world_size(), dist.get_rank()
    M_local = torch.as_tensor(M_local, dtype=torch.float64)
    b_local = torch.as_tensor(b_local, dtype=torch.float64)
    n_local, nb = M_local.shape[0], b_local.shape[-1]
    lo, hi = morton_partition(n_cells_total, rank, world)
    assert hi - lo == n_local, "shard does not match the Morton partition"
    pack = torch.cat([M_local.reshape(n_local, nb * nb), b_local.reshape(n_local, nb)], dim=1).contiguous()
    if device is not None:
        pack = pack.to(device)
    outs = []
    for q in range(world):
        a, b = morton_partition(n_cells_total, q, world)
        outs.append(torch.empty((b - a, nb * nb + nb), dtype=torch.float64, device=pack.device))
    dist.all_gather(outs, pack)
    full = torch.cat(outs, dim=0)
    return full[:, :nb * nb].reshape(-1, nb, nb), full[:, nb * nb:].reshape(-1, nb)


def max_over_ranks(value, device=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
