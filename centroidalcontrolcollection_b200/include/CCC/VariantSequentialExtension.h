/* CCC/VariantSequentialExtension.h — sequential extension of a time-variant model list whose input
 * dimension may change (and be zero) from stage to stage:
 *   x_seq = A_seq x_0 + B_seq u_seq + E_seq
 * Mirrors reference include/CCC/VariantSequentialExtension.h (constructor :80-88, totalStateDim /
 * totalInputDim / totalOutputDim :91-106, setup :110-208, members A_seq_, B_seq_, E_seq_, model_list_).
 */
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "StateSpaceModel.h"

namespace CCC
{
class VariantSequentialExtension
{
public:
  using Matrix = detail::Matrix;

  VariantSequentialExtension(const std::vector<std::shared_ptr<StateSpaceModel>> & model_list, bool extend_for_output = false)
  : model_list_(model_list)
  {
    setup(extend_for_output);
  }

  int totalStateDim() const { return total_state_dim_; }
  int totalInputDim() const { return total_input_dim_; }
  int totalOutputDim() const { return total_output_dim_; }

protected:
  void setup(bool extend_for_output)
  {
    const int L = static_cast<int>(model_list_.size());
    if(L == 0) throw std::runtime_error("[VariantSequentialExtension] model_list is empty");
    const int n = model_list_[0]->stateDim();
    total_state_dim_ = L * n;
    total_input_dim_ = 0;
    total_output_dim_ = 0;
    for(const auto & model : model_list_)
    {
      if(model->dt_ <= 0) throw std::runtime_error("[VariantSequentialExtension] model is not discretized");
      if(model->stateDim() != n) throw std::runtime_error("[VariantSequentialExtension] state dimensions differ");
      total_input_dim_ += model->inputDim();
      total_output_dim_ += model->outputDim();
    }
    A_seq_ = Matrix(total_state_dim_, n);
    B_seq_ = Matrix(total_state_dim_, total_input_dim_);
    E_seq_.assign(total_state_dim_, 0.0);
    int acc = 0;
    Matrix a_prod;
    std::vector<double> e(n, 0.0);
    for(int i = 0; i < L; i++)
    {
      const auto & mi = *model_list_[i];
      const int m = mi.inputDim();
      a_prod = i == 0 ? mi.Ad_ : mi.Ad_ * a_prod;
      A_seq_.setBlock(i * n, 0, a_prod);
      // block column of stage i's input: Bd_i in block row i, then propagated by Ad_j
      Matrix col = mi.Bd_;
      for(int j = i; j < L; j++)
      {
        if(j > i) col = model_list_[j]->Ad_ * col;
        B_seq_.setBlock(j * n, acc, col);
      }
      std::vector<double> en(mi.Ed_);
      if(i > 0)
        for(int r = 0; r < n; r++)
          for(int c = 0; c < n; c++) en[r] += mi.Ad_(r, c) * e[c];
      e = en;
      for(int r = 0; r < n; r++) E_seq_[static_cast<size_t>(i) * n + r] = e[r];
      acc += m;
    }
    if(extend_for_output)
    {
      // the reference's extension for outputs requires D = 0 and drops F (:188-206)
      Matrix a_out(total_output_dim_, n), b_out(total_output_dim_, total_input_dim_);
      std::vector<double> e_out(total_output_dim_, 0.0);
      int acc_o = 0;
      for(int i = 0; i < L; i++)
      {
        const auto & mi = *model_list_[i];
        const int p = mi.outputDim();
        a_out.setBlock(acc_o, 0, mi.C_ * A_seq_.block(i * n, 0, n, n));
        b_out.setBlock(acc_o, 0, mi.C_ * B_seq_.block(i * n, 0, n, total_input_dim_));
        for(int r = 0; r < p; r++)
          for(int c = 0; c < n; c++) e_out[acc_o + r] += mi.C_(r, c) * E_seq_[static_cast<size_t>(i) * n + c];
        acc_o += p;
      }
      A_seq_ = a_out;
      B_seq_ = b_out;
      E_seq_ = e_out;
    }
  }

public:
  std::vector<std::shared_ptr<StateSpaceModel>> model_list_;
  int total_state_dim_ = 0, total_input_dim_ = 0, total_output_dim_ = 0;
  Matrix A_seq_, B_seq_;
  std::vector<double> E_seq_;
};
} // namespace CCC
