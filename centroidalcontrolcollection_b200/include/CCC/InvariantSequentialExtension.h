/* CCC/InvariantSequentialExtension.h — sequential extension ("condensing") of a time-invariant model:
 *   x_seq = A_seq x_0 + B_seq u_seq + E_seq,   x_seq = (x_1 .. x_N), u_seq = (u_0 .. u_{N-1})
 * Mirrors reference include/CCC/InvariantSequentialExtension.h (constructor :74-81, totalStateDim /
 * totalInputDim / totalOutputDim :84-99, setup :103-181, members A_seq_, B_seq_, E_seq_, seq_len_).
 * B_seq is block lower-triangular Toeplitz: block row i of its first block column is Ad^i Bd and every
 * other block is a copy along the diagonal, which is how it is filled here.
 */
#pragma once
#include <memory>
#include <stdexcept>
#include <string>

#include "StateSpaceModel.h"

namespace CCC
{
class InvariantSequentialExtension
{
public:
  using Matrix = detail::Matrix;

  InvariantSequentialExtension(const std::shared_ptr<StateSpaceModel> & model, int seq_len, bool extend_for_output = false)
  : model_(model), seq_len_(seq_len)
  {
    setup(extend_for_output);
  }

  int totalStateDim() const { return seq_len_ * model_->stateDim(); }
  int totalInputDim() const { return seq_len_ * model_->inputDim(); }
  int totalOutputDim() const { return seq_len_ * model_->outputDim(); }

protected:
  void setup(bool extend_for_output)
  {
    if(seq_len_ <= 0) throw std::runtime_error("[InvariantSequentialExtension] seq_len must be positive: " + std::to_string(seq_len_));
    if(model_->dt_ <= 0) throw std::runtime_error("[InvariantSequentialExtension] model is not discretized");
    const int n = model_->stateDim(), m = model_->inputDim(), N = seq_len_;
    A_seq_ = Matrix(N * n, n);
    B_seq_ = Matrix(N * n, N * m);
    E_seq_.assign(static_cast<size_t>(N) * n, 0.0);
    Matrix a_pow = model_->Ad_; // Ad^(i+1)
    Matrix ab = model_->Bd_;    // Ad^i Bd
    std::vector<double> e = model_->Ed_;
    for(int i = 0; i < N; i++)
    {
      if(i > 0)
      {
        a_pow = model_->Ad_ * a_pow;
        ab = model_->Ad_ * ab;
        std::vector<double> en(model_->Ed_);
        for(int r = 0; r < n; r++)
          for(int c = 0; c < n; c++) en[r] += model_->Ad_(r, c) * e[c];
        e = en;
      }
      A_seq_.setBlock(i * n, 0, a_pow);
      for(int j = 0; j + i < N; j++) B_seq_.setBlock((i + j) * n, j * m, ab);
      for(int r = 0; r < n; r++) E_seq_[static_cast<size_t>(i) * n + r] = e[r];
    }
    if(extend_for_output)
    {
      // y_i = C x_i (the reference requires D = 0 here and drops F, :166-179)
      const int p = model_->outputDim();
      const Matrix & C = model_->C_;
      Matrix a_out(N * p, n), b_out(N * p, N * m);
      std::vector<double> e_out(static_cast<size_t>(N) * p, 0.0);
      for(int i = 0; i < N; i++)
      {
        a_out.setBlock(i * p, 0, C * A_seq_.block(i * n, 0, n, n));
        b_out.setBlock(i * p, 0, C * B_seq_.block(i * n, 0, n, N * m));
        for(int r = 0; r < p; r++)
          for(int c = 0; c < n; c++) e_out[static_cast<size_t>(i) * p + r] += C(r, c) * E_seq_[static_cast<size_t>(i) * n + c];
      }
      A_seq_ = a_out;
      B_seq_ = b_out;
      E_seq_ = e_out;
    }
  }

public:
  std::shared_ptr<StateSpaceModel> model_;
  int seq_len_ = 0;
  Matrix A_seq_, B_seq_;
  std::vector<double> E_seq_;
};
} // namespace CCC
