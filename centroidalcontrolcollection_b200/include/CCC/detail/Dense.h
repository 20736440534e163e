/* CCC/detail/Dense.h — the little dense linear algebra the host-side formulation needs (setup time:
 * discretisation, condensing, Riccati doubling, QP coefficient assembly).  Eigen is absent from this
 * image; this is a plain row-major double matrix with the handful of operations used by the drop-in
 * classes.  Nothing here is on the per-solve hot path (that is the CUDA engine behind ccc_b200.h).
 */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <vector>

namespace CCC
{
namespace detail
{
class Matrix
{
public:
  Matrix() {}
  Matrix(int rows, int cols, double v = 0.0) : r_(rows), c_(cols), d_(static_cast<size_t>(rows) * cols, v)
  {
    if(rows < 0 || cols < 0) throw std::invalid_argument("Matrix: negative dimension");
  }
  static Matrix Identity(int n)
  {
    Matrix m(n, n);
    for(int i = 0; i < n; i++) m(i, i) = 1.0;
    return m;
  }
  static Matrix Diagonal(const std::vector<double> & v)
  {
    Matrix m(static_cast<int>(v.size()), static_cast<int>(v.size()));
    for(size_t i = 0; i < v.size(); i++) m(static_cast<int>(i), static_cast<int>(i)) = v[i];
    return m;
  }
  static Matrix Column(const std::vector<double> & v)
  {
    Matrix m(static_cast<int>(v.size()), 1);
    m.d_ = v;
    return m;
  }

  int rows() const { return r_; }
  int cols() const { return c_; }
  double & operator()(int i, int j) { return d_[static_cast<size_t>(i) * c_ + j]; }
  double operator()(int i, int j) const { return d_[static_cast<size_t>(i) * c_ + j]; }
  double * data() { return d_.data(); }
  const double * data() const { return d_.data(); }
  const std::vector<double> & vec() const { return d_; }

  Matrix transpose() const
  {
    Matrix t(c_, r_);
    for(int i = 0; i < r_; i++)
      for(int j = 0; j < c_; j++) t(j, i) = (*this)(i, j);
    return t;
  }
  Matrix block(int i0, int j0, int nr, int nc) const
  {
    Matrix b(nr, nc);
    for(int i = 0; i < nr; i++)
      for(int j = 0; j < nc; j++) b(i, j) = (*this)(i0 + i, j0 + j);
    return b;
  }
  void setBlock(int i0, int j0, const Matrix & b)
  {
    for(int i = 0; i < b.r_; i++)
      for(int j = 0; j < b.c_; j++) (*this)(i0 + i, j0 + j) = b(i, j);
  }
  /** Frobenius norm (Eigen's MatrixBase::norm()). */
  double norm() const
  {
    double s = 0;
    for(double v : d_) s += v * v;
    return std::sqrt(s);
  }
  double norm1() const
  {
    double best = 0;
    for(int j = 0; j < c_; j++)
    {
      double s = 0;
      for(int i = 0; i < r_; i++) s += std::fabs((*this)(i, j));
      best = std::max(best, s);
    }
    return best;
  }

  Matrix operator*(const Matrix & o) const
  {
    if(c_ != o.r_) throw std::invalid_argument("Matrix product: inner dimensions differ");
    Matrix p(r_, o.c_);
    for(int i = 0; i < r_; i++)
      for(int k = 0; k < c_; k++)
      {
        const double a = (*this)(i, k);
        if(a == 0.0) continue;
        const double * brow = &o.d_[static_cast<size_t>(k) * o.c_];
        double * prow = &p.d_[static_cast<size_t>(i) * o.c_];
        for(int j = 0; j < o.c_; j++) prow[j] += a * brow[j];
      }
    return p;
  }
  Matrix operator+(const Matrix & o) const
  {
    check(o);
    Matrix s(*this);
    for(size_t i = 0; i < d_.size(); i++) s.d_[i] += o.d_[i];
    return s;
  }
  Matrix operator-(const Matrix & o) const
  {
    check(o);
    Matrix s(*this);
    for(size_t i = 0; i < d_.size(); i++) s.d_[i] -= o.d_[i];
    return s;
  }
  Matrix operator*(double a) const
  {
    Matrix s(*this);
    for(double & v : s.d_) v *= a;
    return s;
  }

  /** Inverse by Gauss-Jordan elimination with partial pivoting (small matrices only). */
  Matrix inverse() const
  {
    if(r_ != c_) throw std::invalid_argument("Matrix inverse: not square");
    const int n = r_;
    Matrix a(*this), inv = Identity(n);
    for(int col = 0; col < n; col++)
    {
      int piv = col;
      for(int i = col + 1; i < n; i++)
        if(std::fabs(a(i, col)) > std::fabs(a(piv, col))) piv = i;
      if(a(piv, col) == 0.0) throw std::runtime_error("Matrix inverse: singular");
      if(piv != col)
        for(int j = 0; j < n; j++)
        {
          std::swap(a(piv, j), a(col, j));
          std::swap(inv(piv, j), inv(col, j));
        }
      const double d = 1.0 / a(col, col);
      for(int j = 0; j < n; j++)
      {
        a(col, j) *= d;
        inv(col, j) *= d;
      }
      for(int i = 0; i < n; i++)
      {
        if(i == col) continue;
        const double f = a(i, col);
        if(f == 0.0) continue;
        for(int j = 0; j < n; j++)
        {
          a(i, j) -= f * a(col, j);
          inv(i, j) -= f * inv(col, j);
        }
      }
    }
    return inv;
  }

  /** Matrix exponential: scaling and squaring around a degree-18 Taylor polynomial (norm <= 1/2 after
   *  scaling: truncation error below 1e-22).  The discretisation matrices of this library are small and
   *  mostly nilpotent, for which the series terminates exactly. */
  Matrix exp() const
  {
    if(r_ != c_) throw std::invalid_argument("Matrix exp: not square");
    const int n = r_;
    int squarings = 0;
    double nrm = norm1();
    while(nrm > 0.5)
    {
      nrm *= 0.5;
      squarings++;
    }
    const Matrix a = (*this) * std::ldexp(1.0, -squarings);
    Matrix term = Identity(n), sum = Identity(n);
    for(int k = 1; k <= 18; k++)
    {
      term = (term * a) * (1.0 / k);
      sum = sum + term;
      if(term.norm1() == 0.0) break; // nilpotent
    }
    for(int s = 0; s < squarings; s++) sum = sum * sum;
    return sum;
  }

private:
  void check(const Matrix & o) const
  {
    if(r_ != o.r_ || c_ != o.c_) throw std::invalid_argument("Matrix sum: shapes differ");
  }
  int r_ = 0, c_ = 0;
  std::vector<double> d_;
};
} // namespace detail
} // namespace CCC
