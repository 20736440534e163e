/* CCC/StateSpaceModel.h — state-space model with zero-order-hold discretisation (host side, setup time).
 *
 * Mirrors reference include/CCC/StateSpaceModel.h: same members (A_, B_, C_, D_, E_, F_, dt_, Ad_, Bd_, Ed_),
 * same checks (:43-93), same methods (stateDim/inputDim/outputDim :101-116, stateEq :123, stateEqDisc :133,
 * observEq :142-155, calcDiscMatrix :164-216).  The reference's template dimensions become run-time
 * dimensions (Eigen is absent from this image; see detail/Dense.h).
 *   x' = A x + B u + E,   y = C x + D u + F
 */
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "detail/Dense.h"

namespace CCC
{
class StateSpaceModel
{
public:
  using Matrix = detail::Matrix;
  using Vector = std::vector<double>;

  StateSpaceModel(int state_dim, int input_dim, int output_dim)
  : state_dim_(state_dim), input_dim_(input_dim), output_dim_(output_dim)
  {
    if(state_dim_ <= 0) throw std::runtime_error("[StateSpaceModel] state_dim must be positive: " + std::to_string(state_dim_));
    if(input_dim_ < 0) throw std::runtime_error("[StateSpaceModel] input_dim must be non-negative: " + std::to_string(input_dim_));
    if(output_dim_ < 0) throw std::runtime_error("[StateSpaceModel] output_dim must be non-negative: " + std::to_string(output_dim_));
    A_ = Matrix(state_dim_, state_dim_);
    B_ = Matrix(state_dim_, input_dim_);
    C_ = Matrix(output_dim_, state_dim_);
    D_ = Matrix(output_dim_, input_dim_);
    E_.assign(state_dim_, 0.0);
    F_.assign(output_dim_, 0.0);
  }
  virtual ~StateSpaceModel() = default;

  int stateDim() const { return state_dim_; }
  int inputDim() const { return input_dim_; }
  int outputDim() const { return output_dim_; }

  /** Continuous state equation: A x + B u + E. */
  Vector stateEq(const Vector & x, const Vector & u) const { return affine(A_, B_, E_, x, u); }
  /** Discrete state equation: Ad x + Bd u + Ed (calcDiscMatrix must have been called). */
  Vector stateEqDisc(const Vector & x, const Vector & u) const
  {
    if(dt_ <= 0) throw std::runtime_error("[StateSpaceModel] dt is not positive: " + std::to_string(dt_));
    return affine(Ad_, Bd_, Ed_, x, u);
  }
  Vector observEq(const Vector & x) const { return affine(C_, Matrix(output_dim_, 0), F_, x, Vector()); }
  Vector observEq(const Vector & x, const Vector & u) const { return affine(C_, D_, F_, x, u); }

  /** Zero-order hold through exp([A B (E); 0] dt); the E column is appended only when E != 0
   *  (reference :164-216). */
  void calcDiscMatrix(double dt)
  {
    dt_ = dt;
    const int n = state_dim_, m = input_dim_;
    double e_norm = 0;
    for(double v : E_) e_norm += v * v;
    const bool with_e = e_norm != 0;
    const int dim = n + m + (with_e ? 1 : 0);
    Matrix big(dim, dim);
    for(int i = 0; i < n; i++)
    {
      for(int j = 0; j < n; j++) big(i, j) = dt * A_(i, j);
      for(int j = 0; j < m; j++) big(i, n + j) = dt * B_(i, j);
      if(with_e) big(i, n + m) = dt * E_[i];
    }
    const Matrix x = big.exp();
    Ad_ = x.block(0, 0, n, n);
    Bd_ = x.block(0, n, n, m);
    Ed_.assign(n, 0.0);
    if(with_e)
      for(int i = 0; i < n; i++) Ed_[i] = x(i, n + m);
  }

public:
  const int state_dim_ = 0;
  const int input_dim_ = 0;
  const int output_dim_ = 0;
  Matrix A_, B_, C_, D_;
  Vector E_, F_;
  double dt_ = 0;
  Matrix Ad_, Bd_;
  Vector Ed_;

private:
  static Vector affine(const Matrix & a, const Matrix & b, const Vector & e, const Vector & x, const Vector & u)
  {
    Vector y(e);
    for(int i = 0; i < a.rows(); i++)
    {
      double s = y[i];
      for(int j = 0; j < a.cols(); j++) s += a(i, j) * x[j];
      for(int j = 0; j < b.cols(); j++) s += b(i, j) * u[j];
      y[i] = s;
    }
    return y;
  }
};
} // namespace CCC
