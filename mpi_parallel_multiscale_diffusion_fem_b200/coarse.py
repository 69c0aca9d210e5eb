"""Coarse mesh of the MsFEM problem: hyper_cube(0,1) refined r times
(/root/reference/include/base/diffusion_problem_ms.tpp:91-103), cells in Morton /
CellId / p4est order, and the p4est partition rule."""
import numpy as np


def _compact(v):
    v = v & np.uint64(0x5555555555555555)
    v = (v | (v >> np.uint64(1))) & np.uint64(0x3333333333333333)
    v = (v | (v >> np.uint64(2))) & np.uint64(0x0F0F0F0F0F0F0F0F)
    v = (v | (v >> np.uint64(4))) & np.uint64(0x00FF00FF00FF00FF)
    v = (v | (v >> np.uint64(8))) & np.uint64(0x0000FFFF0000FFFF)
    v = (v | (v >> np.uint64(16))) & np.uint64(0x00000000FFFFFFFF)
    return v


def coarse_corners(r, lo=0, hi=None):
    """Corner points [hi-lo, 4, 2] (deal.II vertex order) of the coarse cells with Morton
    index in [lo, hi) of the 2^r x 2^r mesh on the unit square."""
    nc = 1 << r
    hi = nc * nc if hi is None else hi
    m = np.arange(lo, hi, dtype=np.uint64)
    ix = _compact(m).astype(np.float64)
    iy = _compact(m >> np.uint64(1)).astype(np.float64)
    H = 1.0 / nc
    out = np.empty((m.size, 4, 2), dtype=np.float64)
    for v in range(4):
        out[:, v, 0] = (ix + (v & 1)) * H
        out[:, v, 1] = (iy + (v >> 1)) * H
    return out


def coarse_corners3(r, lo=0, hi=None):
    """Corner points [hi-lo, 8, 3] (deal.II hex vertex order, x fastest) of the coarse cells with
    3D Morton index in [lo, hi) of the (2^r)^3 mesh on the unit cube."""
    nc = 1 << r
    hi = nc ** 3 if hi is None else hi
    m = np.arange(lo, hi, dtype=np.uint64)
    idx = np.zeros((3, m.size), dtype=np.uint64)
    for bit in range(r):
        for a in range(3):
            idx[a] |= ((m >> np.uint64(3 * bit + a)) & np.uint64(1)) << np.uint64(bit)
    H = 1.0 / nc
    out = np.empty((m.size, 8, 3), dtype=np.float64)
    for v in range(8):
        out[:, v, 0] = (idx[0] + (v & 1)) * H
        out[:, v, 1] = (idx[1] + ((v >> 1) & 1)) * H
        out[:, v, 2] = (idx[2] + (v >> 2)) * H
    return out


def morton_partition(n_cells, rank, world):
    """Contiguous Z-curve range of rank `rank` of `world` (uniform weights): the p4est rule the
    reference inherits for cell ownership (ms.tpp:52)."""
    return (n_cells * rank) // world, (n_cells * (rank + 1)) // world


def cell_id_string(r, m):
    """CellId::to_string() of coarse cell m at depth r: '0_r:<child digits>' (SURVEY A.6)."""
    digits = "".join(str((m >> (2 * (r - 1 - k))) & 3) for k in range(r))
    return "0_%d:%s" % (r, digits)
