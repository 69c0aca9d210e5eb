// msb_coeff.cuh -- device twins of the reference's coefficient classes, shared by the stencil
// assembly kernel (msb_setup.cu) and the fused solve kernel (msb_solve_fused.cu).
#pragma once

#include <math.h>

#include "msb_internal.cuh"

namespace msb
{
  // ======================================================================================
  // Coefficient evaluation (device twins of include/coefficients/matrix_coeff.tpp and of
  // the BASELINE.md synthetic coefficients).
  // ======================================================================================
  __device__ inline uint64_t
  mix64(uint64_t z)
  {
    z ^= z >> 33;
    z *= 0xff51afd7ed558ccdULL;
    z ^= z >> 33;
    z *= 0xc4ceb9fe1a85ec53ULL;
    z ^= z >> 33;
    return z;
  }

  struct CoeffEval
  {
    int    kind, seed;
    double par[6];
    double rot00, rot01, rot10, rot11; // reference rotation (matrix_coeff.tpp:17-25)

    // REFERENCE and PERIODIC are a(x,y) = 1 - c (sin(kx)/2 + sin(ky)/2): separable sines
    __device__ inline bool
    separable() const
    {
      return kind == MSB_COEFF_REFERENCE || kind == MSB_COEFF_PERIODIC;
    }

    // the 1-D sine term of coordinate t, evaluated exactly as the reference writes it
    __device__ inline double
    sine_term(double t) const
    {
      if (kind == MSB_COEFF_REFERENCE)
        {
          // coefficients.h:21 (sic) and matrix_coeff.hpp:45: sin(2 * PI_D * k * p(d))
          const double PI_D = 3.14592653509793218403;
          return sin(2 * PI_D * 57 * t);
        }
      const double PI = 3.14159265358979323846;
      return sin(2 * PI * t / par[0]);
    }

    __device__ inline void
    from_sines(double sx, double sy, double &a00, double &a01, double &a10, double &a11) const
    {
      if (kind == MSB_COEFF_REFERENCE)
        {
          // matrix_coeff.hpp:46, matrix_coeff.tpp:78-89
          const double a = 1.0 * (1.0 - 0.9999 * (0.5 * sx + 0.5 * sy));
          // values = rot * (a I) * transpose(rot), evaluated in that order
          const double t00 = rot00 * a, t01 = rot01 * a, t10 = rot10 * a, t11 = rot11 * a;
          a00 = t00 * rot00 + t01 * rot01;
          a01 = t00 * rot10 + t01 * rot11;
          a10 = t10 * rot00 + t11 * rot01;
          a11 = t10 * rot10 + t11 * rot11;
        }
      else
        {
          const double a = 1.0 - par[1] * (0.5 * sx + 0.5 * sy);
          a00 = a, a01 = 0.0, a10 = 0.0, a11 = a;
        }
    }

    __device__ inline void
    operator()(double x, double y, double &a00, double &a01, double &a10, double &a11) const
    {
      if (separable())
        from_sines(sine_term(x), sine_term(y), a00, a01, a10, a11);
      else if (kind == MSB_COEFF_INCLUSIONS)
        {
          const long long bx = (long long)floor(x / par[0]), by = (long long)floor(y / par[0]);
          uint64_t        h  = (uint64_t)bx * 0x9E3779B97F4A7C15ULL;
          h ^= mix64((uint64_t)by + 0xC2B2AE3D27D4EB4FULL * (uint64_t)(uint32_t)seed);
          h = mix64(h);
          const bool   in = (double)(h >> 11) * (1.0 / 9007199254740992.0) < par[1];
          const double a  = in ? par[2] : par[3];
          a00 = a, a01 = 0.0, a10 = 0.0, a11 = a;
        }
      else
        {
          a00 = par[0], a01 = 0.0, a10 = 0.0, a11 = par[0];
        }
    }
  };


  // host side: descriptor -> evaluator (rotation of matrix_coeff.hpp:48, matrix_coeff.tpp:17-25: alpha = PI_D/3)
  inline CoeffEval
  make_coeff_eval(const msb_coeff_desc &d)
  {
    CoeffEval c;
    c.kind = d.kind;
    c.seed = d.seed;
    for (int i = 0; i < 6; ++i)
      c.par[i] = d.par[i];
    const double alpha = 3.14592653509793218403 / 3;
    c.rot00 = cos(alpha), c.rot01 = sin(alpha);
    c.rot10 = -sin(alpha), c.rot11 = cos(alpha);
    return c;
  }
} // namespace msb
