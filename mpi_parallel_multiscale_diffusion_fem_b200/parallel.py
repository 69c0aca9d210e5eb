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


class _DeviceArray:
    """A device buffer owned by the C library, exposed through __cuda_array_interface__ so that
    torch can wrap it WITHOUT a copy (torch.as_tensor)."""

    def __init__(self, addr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(addr), False),
                                         "version": 2, "strides": None}


def device_result_tensors(shard, device):
    """(M [n,nb,nb], b [n,nb], iters [n,nb]) of the last run as torch CUDA tensors that alias the
    library's device memory (msb_get_device_results); valid until the next set_cells / run."""
    dM, db, dit = shard.device_results()
    n, nb = shard.n_cells, shard.nb
    M = torch.as_tensor(_DeviceArray(dM, (n, nb, nb), "<f8"), device=device)
    b = torch.as_tensor(_DeviceArray(db, (n, nb), "<f8"), device=device)
    it = torch.as_tensor(_DeviceArray(dit, (n, nb), "<i4"), device=device)
    return M, b, it


def gather_coarse_contributions_device(shard, n_cells_total, device, out_M=None, out_b=None):
    """The compress(add) exchange (ms.tpp:253-254) straight from the library's device buffers: NCCL
    all_gather of M and b of every rank (no host round trip, no packing kernel); returns CUDA
    tensors (M [C,nb,nb], b [C,nb]) in Morton order on every rank."""
    world, rank = dist.get_world_size(), dist.get_rank()
    M, b, _ = device_result_tensors(shard, device)
    nb = shard.nb
    sizes = [morton_partition(n_cells_total, q, world) for q in range(world)]
    assert sizes[rank][1] - sizes[rank][0] == shard.n_cells, "shard does not match the Morton partition"
    if out_M is None:
        out_M = torch.empty((n_cells_total, nb, nb), dtype=torch.float64, device=device)
    if out_b is None:
        out_b = torch.empty((n_cells_total, nb), dtype=torch.float64, device=device)
    if all(hi - lo == shard.n_cells for lo, hi in sizes):
        dist.all_gather_into_tensor(out_M, M)
        dist.all_gather_into_tensor(out_b, b)
    else:
        dist.all_gather([out_M[lo:hi] for lo, hi in sizes], M)
        dist.all_gather([out_b[lo:hi] for lo, hi in sizes], b)
    return out_M, out_b


def max_over_ranks(value, device=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
