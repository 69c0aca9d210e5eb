// shim_types.hpp -- deal.II-free restatements of the value types the reference's basis interface exposes
// (Point, Tensor<2,dim>, Vector, FullMatrix, CellId), in the namespace MSFEM_SHIM_NAMESPACE.  Included by
// shims.hpp as namespace msfem; tests/fake_dealii re-exports the same classes as namespace dealii to compile
// the MSFEM_WITH_DEALII branch of the mirror without deal.II.  No include guard on purpose (one inclusion per
// namespace).
#include <array>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace MSFEM_SHIM_NAMESPACE
{
  template <int dim>
  class Point
  {
  public:
    Point() { c_.fill(0.0); }
    Point(double x, double y)
    {
      static_assert(dim == 2, "2-argument Point is 2D");
      c_[0] = x, c_[1] = y;
    }
    Point(double x, double y, double z)
    {
      static_assert(dim == 3, "3-argument Point is 3D");
      c_[0] = x, c_[1] = y, c_[2] = z;
    }
    double  operator()(unsigned i) const { return c_[i]; }
    double &operator()(unsigned i) { return c_[i]; }
    double  operator[](unsigned i) const { return c_[i]; }
    double &operator[](unsigned i) { return c_[i]; }

  private:
    std::array<double, dim> c_;
  };

  // rank-2 tensor only (the diffusion coefficient)
  template <int rank, int dim>
  class Tensor;

  template <int dim>
  class Tensor<2, dim>
  {
  public:
    Tensor() { clear(); }
    void clear()
    {
      for (auto &row : a_)
        row.fill(0.0);
    }
    std::array<double, dim>       &operator[](unsigned i) { return a_[i]; }
    const std::array<double, dim> &operator[](unsigned i) const { return a_[i]; }

  private:
    std::array<std::array<double, dim>, dim> a_;
  };

  template <int dim>
  Tensor<2, dim>
  operator*(const Tensor<2, dim> &A, const Tensor<2, dim> &B)
  {
    Tensor<2, dim> C;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        {
          double s = 0.0;
          for (int k = 0; k < dim; ++k)
            s += A[i][k] * B[k][j];
          C[i][j] = s;
        }
    return C;
  }

  template <int dim>
  Tensor<2, dim>
  transpose(const Tensor<2, dim> &A)
  {
    Tensor<2, dim> T;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        T[i][j] = A[j][i];
    return T;
  }

  template <typename number>
  class Vector
  {
  public:
    Vector() = default;
    explicit Vector(std::size_t n)
      : v_(n, number(0))
    {}
    void        reinit(std::size_t n) { v_.assign(n, number(0)); }
    std::size_t size() const { return v_.size(); }
    number      operator()(std::size_t i) const { return v_[i]; }
    number     &operator()(std::size_t i) { return v_[i]; }
    number      operator[](std::size_t i) const { return v_[i]; }
    number     &operator[](std::size_t i) { return v_[i]; }
    number     *data() { return v_.data(); }
    const number *data() const { return v_.data(); }
    number      operator*(const Vector &o) const
    {
      number s = 0;
      for (std::size_t i = 0; i < v_.size(); ++i)
        s += v_[i] * o.v_[i];
      return s;
    }

  private:
    std::vector<number> v_;
  };

  // row-major dense matrix, FullMatrix<double>-like
  template <typename number>
  class FullMatrix
  {
  public:
    FullMatrix() = default;
    FullMatrix(std::size_t m, std::size_t n)
      : m_(m)
      , n_(n)
      , v_(m * n, number(0))
    {}
    std::size_t m() const { return m_; }
    std::size_t n() const { return n_; }
    number      operator()(std::size_t i, std::size_t j) const { return v_[i * n_ + j]; }
    number     &operator()(std::size_t i, std::size_t j) { return v_[i * n_ + j]; }
    number     *data() { return v_.data(); }
    const number *data() const { return v_.data(); }

  private:
    std::size_t         m_ = 0, n_ = 0;
    std::vector<number> v_;
  };

  // identifies a coarse cell of the refined hyper_cube: Morton index at a given depth;
  // to_string() follows deal.II's "coarse_depth:child digits" form
  class CellId
  {
  public:
    CellId() = default;
    CellId(unsigned depth, std::uint64_t morton, unsigned dim = 2)
      : depth_(depth)
      , dim_(dim)
      , morton_(morton)
    {}
    std::string to_string() const
    {
      std::string s = "0_" + std::to_string(depth_) + ":";
      for (unsigned k = 0; k < depth_; ++k)
        s += char('0' + ((morton_ >> (dim_ * (depth_ - 1 - k))) & ((1u << dim_) - 1u)));
      return s;
    }
    bool          operator<(const CellId &o) const { return morton_ < o.morton_; }
    bool          operator==(const CellId &o) const { return depth_ == o.depth_ && morton_ == o.morton_; }
    std::uint64_t morton() const { return morton_; }
    unsigned      depth() const { return depth_; }

  private:
    unsigned      depth_  = 0;
    unsigned      dim_    = 2;
    std::uint64_t morton_ = 0;
  };
} // namespace MSFEM_SHIM_NAMESPACE
