// coefficients.hpp -- host-side coefficient classes with the reference's names and call
// signatures (value / value_list), for dim = 2.  They exist so that host code written against
// /root/reference/include/coefficients/*.hpp keeps compiling; inside the basis stage the SAME
// functions are evaluated on the device (csrc/msb_setup.cu, CoeffEval) and a class whose
// formula the device does not know is passed through its value_list as a quadrature-point
// table (MSB_COEFF_TABLE).
//
//   MatrixCoeff     matrix_coeff.hpp:32-49, matrix_coeff.tpp:17-25,45-91
//   RightHandSide   right_hand_side.tpp:17-40
//   BasisQ1         basis_q1.hpp:33-101, basis_q1.tpp:26-47,86-133
//   DirichletBC     dirichlet_bc.tpp:17-43
//   NeumannBC       neumann_bc.tpp:17-41
#pragma once

#include <cassert>
#include <cmath>
#include <vector>

#include "msfem/shims.hpp"
#include "msfem_basis.h"

namespace Coefficients
{
  using namespace msfem;

  // coefficients.h:21 -- not pi; reproduced verbatim because it defines the reference problem
  const double PI_D = 3.14592653509793218403;
  const float  PI_F = 3.14159265358979f;

  // Anything that can hand the basis stage a diffusion tensor: either a formula the device
  // knows (descriptor) or values tabulated through value_list.
  template <int dim>
  class TensorCoefficient
  {
  public:
    virtual ~TensorCoefficient() = default;
    virtual Tensor<2, dim> value(const Point<dim> &p) const = 0;
    virtual void           value_list(const std::vector<Point<dim>> &points,
                                      std::vector<Tensor<2, dim>>   &values) const
    {
      assert(points.size() == values.size());
      for (std::size_t i = 0; i < points.size(); ++i)
        values[i] = value(points[i]);
    }
    // MSB_COEFF_TABLE unless a subclass knows better
    virtual msb_coeff_desc device_descriptor() const
    {
      msb_coeff_desc d{};
      d.kind = MSB_COEFF_TABLE;
      return d;
    }
  };

  // A(x) = R (a(x) I) R^T, a = 1 - 0.9999 (sin(2 PI_D 57 x)/2 + sin(2 PI_D 57 y)/2)
  template <int dim>
  class MatrixCoeff : public TensorCoefficient<dim>
  {
  public:
    MatrixCoeff()
    {
      static_assert(dim == 2 || dim == 3, "the reference instantiates dim 2 and 3");
      set_rotation(rot);
    }
    Tensor<2, dim> value(const Point<dim> &p) const override
    {
      Tensor<2, dim> v;
      const double   a = 1.0 * (1.0 - scale_factor * (0.5 * std::sin(2 * PI_D * k * p(0)) +
                                                    0.5 * std::sin(2 * PI_D * k * p(1))));
      for (int d = 0; d < dim; ++d)
        v[d][d] = a;
      return rot * v * transpose(rot);
    }
    msb_coeff_desc device_descriptor() const override
    {
      msb_coeff_desc d{};
      d.kind = MSB_COEFF_REFERENCE;
      return d;
    }

  private:
    // matrix_coeff.tpp:17-25 (dim 2) and :28-41 (dim 3, Euler angles alpha, beta, gamma)
    void set_rotation(Tensor<2, 2> &r) const
    {
      r[0][0] = std::cos(alpha), r[0][1] = std::sin(alpha);
      r[1][0] = -std::sin(alpha), r[1][1] = std::cos(alpha);
    }
    void set_rotation(Tensor<2, 3> &r) const
    {
      const double ca = std::cos(alpha), sa = std::sin(alpha), cb = std::cos(beta), sb = std::sin(beta),
                   cg = std::cos(gamma), sg = std::sin(gamma);
      r[0][0] = ca * cg - sa * cb * sg, r[0][1] = -ca * sg - sa * cb * cg, r[0][2] = sa * sb;
      r[1][0] = sa * cg + ca * cb * sg, r[1][1] = -sa * sg + ca * cb * cg, r[1][2] = -ca * sb;
      r[2][0] = sb * sg, r[2][1] = sb * cg, r[2][2] = cb;
    }

    const int      k            = 57;
    const double   scale_factor = 0.9999;
    const double   alpha        = PI_D / 3;
    const double   beta         = PI_D / 6;
    const double   gamma        = PI_D / 4;
    Tensor<2, dim> rot;
  };

  // BASELINE.md section 4 synthetic coefficients, same interface
  template <int dim>
  class PeriodicCoeff : public TensorCoefficient<dim>
  {
  public:
    explicit PeriodicCoeff(double eps, double scale = 0.9999)
      : eps(eps)
      , scale(scale)
    {}
    Tensor<2, dim> value(const Point<dim> &p) const override
    {
      const double   pi = 3.14159265358979323846;
      Tensor<2, dim> v;
      const double   a =
        1.0 - scale * (0.5 * std::sin(2 * pi * p(0) / eps) + 0.5 * std::sin(2 * pi * p(1) / eps));
      for (int d = 0; d < dim; ++d)
        v[d][d] = a;
      return v;
    }
    msb_coeff_desc device_descriptor() const override
    {
      msb_coeff_desc d{};
      d.kind   = MSB_COEFF_PERIODIC;
      d.par[0] = eps, d.par[1] = scale;
      return d;
    }

  private:
    double eps, scale;
  };

  template <int dim>
  class InclusionCoeff : public TensorCoefficient<dim>
  {
  public:
    InclusionCoeff(double block, double prob, double a_in, double a_out, int seed)
      : block(block)
      , prob(prob)
      , a_in(a_in)
      , a_out(a_out)
      , seed(seed)
    {}
    static std::uint64_t mix(std::uint64_t z)
    {
      z ^= z >> 33, z *= 0xff51afd7ed558ccdULL;
      z ^= z >> 33, z *= 0xc4ceb9fe1a85ec53ULL;
      return z ^ (z >> 33);
    }
    Tensor<2, dim> value(const Point<dim> &p) const override
    {
      const long long bx = (long long)std::floor(p(0) / block), by = (long long)std::floor(p(1) / block);
      std::uint64_t   h  = (std::uint64_t)bx * 0x9E3779B97F4A7C15ULL;
      h ^= mix((std::uint64_t)by + 0xC2B2AE3D27D4EB4FULL * (std::uint64_t)(std::uint32_t)seed);
      h = mix(h);
      const bool     in = (double)(h >> 11) * (1.0 / 9007199254740992.0) < prob;
      Tensor<2, dim> v;
      for (int d = 0; d < dim; ++d)
        v[d][d] = in ? a_in : a_out;
      return v;
    }
    msb_coeff_desc device_descriptor() const override
    {
      msb_coeff_desc d{};
      d.kind = MSB_COEFF_INCLUSIONS, d.seed = seed;
      d.par[0] = block, d.par[1] = prob, d.par[2] = a_in, d.par[3] = a_out;
      return d;
    }

  private:
    double block, prob, a_in, a_out;
    int    seed;
  };

  template <int dim>
  class RightHandSide
  {
  public:
    double value(const Point<dim> & /*p*/, const unsigned int /*component*/ = 0) const { return 2.0; }
    void   value_list(const std::vector<Point<dim>> &points, std::vector<double> &values,
                      const unsigned int = 0) const
    {
      assert(points.size() == values.size());
      for (auto &v : values)
        v = 2.0;
    }
  };

  // the coarse Q1 shape functions of a cell: Dirichlet data of the local problems
  template <int dim>
  class BasisQ1
  {
  public:
    BasisQ1() = delete;
    explicit BasisQ1(const CoarseCell<dim> &cell)
      : index_basis(0)
      , coeff_matrix(1 << dim, 1 << dim)
    {
      static_assert(dim == 2 || dim == 3, "the reference instantiates dim 2 and 3");
      // rows (1, x, y, xy) resp. (1, x, y, z, xy, yz, xz, xyz) at the vertices
      // (basis_q1.tpp:26-47, :50-75), inverted by Gauss-Jordan elimination; column i then
      // holds the monomial coefficients of the basis of vertex i
      constexpr int nb = 1 << dim;
      double        a[nb][2 * nb];
      for (int i = 0; i < nb; ++i)
        {
          double m[8];
          monomials(cell.vertex(i), m);
          for (int j = 0; j < nb; ++j)
            a[i][j] = m[j], a[i][nb + j] = (i == j);
        }
      for (int c = 0; c < nb; ++c)
        {
          int piv = c;
          for (int r = c + 1; r < nb; ++r)
            if (std::fabs(a[r][c]) > std::fabs(a[piv][c]))
              piv = r;
          for (int j = 0; j < 2 * nb; ++j)
            std::swap(a[c][j], a[piv][j]);
          const double inv = 1.0 / a[c][c];
          for (int j = 0; j < 2 * nb; ++j)
            a[c][j] *= inv;
          for (int r = 0; r < nb; ++r)
            if (r != c)
              {
                const double f = a[r][c];
                for (int j = 0; j < 2 * nb; ++j)
                  a[r][j] -= f * a[c][j];
              }
        }
      for (int i = 0; i < nb; ++i)
        for (int j = 0; j < nb; ++j)
          coeff_matrix(i, j) = a[i][nb + j];
    }
    void   set_index(unsigned int index) { index_basis = index; }
    // basis_q1.tpp:86-96 (dim 2), :99-113 (dim 3)
    double value(const Point<dim> &p, const unsigned int /*component*/ = 0) const
    {
      double m[8];
      monomials(p, m);
      double v = 0.0;
      for (int k = 0; k < (1 << dim); ++k)
        v += coeff_matrix(k, index_basis) * m[k];
      return v;
    }
    void value_list(const std::vector<Point<dim>> &points, std::vector<double> &values,
                    const unsigned int = 0) const
    {
      assert(points.size() == values.size());
      for (std::size_t i = 0; i < points.size(); ++i)
        values[i] = value(points[i]);
    }

  private:
    static void monomials(const Point<2> &p, double m[8])
    {
      m[0] = 1, m[1] = p(0), m[2] = p(1), m[3] = p(0) * p(1);
    }
    static void monomials(const Point<3> &p, double m[8])
    {
      m[0] = 1, m[1] = p(0), m[2] = p(1), m[3] = p(2), m[4] = p(0) * p(1), m[5] = p(1) * p(2);
      m[6] = p(0) * p(2), m[7] = p(0) * p(1) * p(2);
    }

    unsigned int       index_basis;
    FullMatrix<double> coeff_matrix;
  };

  template <int dim>
  class DirichletBC
  {
  public:
    double value(const Point<dim> &p, const unsigned int = 0) const
    {
      return (p(0) - 0.5) * (p(0) - 0.5) + (p(1) - 0.5) * (p(1) - 0.5);
    }
  };

  template <int dim>
  class NeumannBC
  {
  public:
    double value(const Point<dim> &p, const unsigned int = 0) const
    {
      return std::cos(2 * PI_D * p(0)) * std::cos(2 * PI_D * p(1));
    }
  };
} // namespace Coefficients
