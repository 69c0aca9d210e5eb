"""B200-native multiscale-basis stage of konsim83/MPI-Parallel-Multiscale-Diffusion-FEM.

The product is the CUDA library `libmsfem_basis.so` (csrc/, C ABI in include/msfem_basis.h)
and the C++ host mirror of the reference interface in host/.  This Python package is
plumbing only: a ctypes binding of the C ABI used by tests/ and bench.py, the coarse-mesh
/ Morton-partition helpers, and the torch.distributed gather of the coarse contributions.
"""
from .binding import (BasisShard, MsbError, CoeffDesc, lib_path, load_library,
                      COEFF_REFERENCE, COEFF_PERIODIC, COEFF_INCLUSIONS, COEFF_CONSTANT,
                      COEFF_TABLE, TIER_AUTO, TIER_SMEM, TIER_STREAMED, EXPORTED_SYMBOLS)
from .coarse import coarse_corners, coarse_corners3, morton_partition, cell_id_string

__all__ = ["BasisShard", "MsbError", "CoeffDesc", "lib_path", "load_library",
           "COEFF_REFERENCE", "COEFF_PERIODIC", "COEFF_INCLUSIONS", "COEFF_CONSTANT",
           "COEFF_TABLE", "TIER_AUTO", "TIER_SMEM", "TIER_STREAMED", "EXPORTED_SYMBOLS",
           "coarse_corners", "coarse_corners3", "morton_partition", "cell_id_string"]
