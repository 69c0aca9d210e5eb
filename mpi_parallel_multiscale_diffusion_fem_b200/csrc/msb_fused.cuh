// msb_fused.cuh -- parameters of the fused basis-stage kernel (msb_solve_fused.cu)
#pragma once

#include "msb_coeff.cuh"
#include "msb_internal.cuh"

namespace msb
{
  struct FusedParams
  {
    const double *corners; // [C][8]
    const double *q1coef;  // [C][16]
    double       *phi;     // [C][4][N]
    double       *M;       // [C][16]
    double       *b;       // [C][4]
    int32_t      *iters;   // [C][4]
    double       *res;     // [C][4]
    int32_t      *fail;
    int           fail_base; // index of this launch's first solve in the shard (ranges of cells)
    double        tol2;
    int           max_iter;
    int           n_cells;
    double        rhs_value;
    int           flavor; // A/B: 0 default, 1 residual in tensor memory, 2 q through tensor memory (variants 10-12)
    int           split;  // 1: one CTA per (cell, pair of bases) instead of one per cell (the short last wave of a small shard)
    CoeffEval     coef;
  };

  // the whole stage (assembly, 4 solves, element matrices) of every cell in ONE launch; l = 5, 6 (32 x 32 and 64 x 64 local meshes)
  cudaError_t launch_solve_fused(const FusedParams &P, int l, cudaStream_t st);
} // namespace msb
