/* CCC/PreviewControl.h — preview control (Kajita 2003) for a discretised state-space model: host side.
 *
 * Mirrors reference include/CCC/PreviewControl.h: WeightParam (:36-53), constructor (:64-81, with the
 * argument check and horizon_steps = ceil(duration / dt)), calcOptimalInput (:86-89), calcGain (:93-172:
 * discrete algebraic Riccati equation by structure-preserving doubling, feedback gain K, preview gains F),
 * public members model_, horizon_dt_, horizon_steps_, P_, K_, F_, riccati_error_.  Template dimensions
 * become run-time dimensions (Eigen is absent).  calcOptimalInput here is the single-problem host form
 * (K is 1x3 and F 1xN for the ZMP model: 203 multiply-adds); batches go through ccc_preview_input, see
 * PreviewControlZmp.h.
 */
#pragma once
#include <cmath>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "StateSpaceModel.h"

namespace CCC
{
class PreviewControl
{
public:
  using Matrix = detail::Matrix;

  struct WeightParam
  {
    //! Output weight (diagonal of Q)
    std::vector<double> output;
    //! Input weight (diagonal of R)
    std::vector<double> input;
  };

  PreviewControl(const std::shared_ptr<StateSpaceModel> & model,
                 double horizon_duration,
                 double horizon_dt,
                 const WeightParam & weight_param)
  : model_(model), horizon_dt_(horizon_dt), horizon_steps_(static_cast<int>(std::ceil(horizon_duration / horizon_dt)))
  {
    if(horizon_duration <= 0 || horizon_dt <= 0)
      throw std::runtime_error("[PreviewControl] Input arguments are invalid. horizon_duration: " + std::to_string(horizon_duration)
                               + ", horizon_dt: " + std::to_string(horizon_dt));
    calcGain(weight_param);
  }
  virtual ~PreviewControl() = default;

  /** u = -K x + F ref_output_seq */
  std::vector<double> calcOptimalInput(const std::vector<double> & x, const std::vector<double> & ref_output_seq) const
  {
    std::vector<double> u(K_.rows(), 0.0);
    for(int i = 0; i < K_.rows(); i++)
    {
      double s = 0;
      for(int j = 0; j < K_.cols(); j++) s -= K_(i, j) * x[j];
      for(int j = 0; j < F_.cols(); j++) s += F_(i, j) * ref_output_seq[j];
      u[i] = s;
    }
    return u;
  }

protected:
  void calcGain(const WeightParam & weight_param)
  {
    if(model_->dt_ != horizon_dt_) model_->calcDiscMatrix(horizon_dt_);
    const Matrix & A = model_->Ad_;
    const Matrix & B = model_->Bd_;
    const Matrix & C = model_->C_;
    const int n = A.rows();
    if(static_cast<int>(weight_param.output.size()) != C.rows() || static_cast<int>(weight_param.input.size()) != B.cols())
      throw std::runtime_error("[PreviewControl] weight dimensions do not match the model");
    const Matrix Q = Matrix::Diagonal(weight_param.output), R = Matrix::Diagonal(weight_param.input);
    std::vector<double> rinv(weight_param.input);
    for(double & v : rinv) v = 1.0 / v;
    const Matrix Rinv = Matrix::Diagonal(rinv);

    // 1. discrete algebraic Riccati equation by doubling (reference :113-144)
    Matrix A0 = A, G0 = B * Rinv * B.transpose(), H0 = C.transpose() * Q * C, H1;
    constexpr int max_iter = 10000;
    constexpr double rel_norm_thre = 1e-8;
    riccati_converged_ = false;
    for(int it = 0; it < max_iter; it++)
    {
      const Matrix W = (Matrix::Identity(n) + G0 * H0).inverse();
      const Matrix A1 = A0 * W * A0;
      const Matrix G1 = G0 + A0 * W * G0 * A0.transpose();
      H1 = H0 + A0.transpose() * H0 * W * A0;
      const double rel_norm = (H1 - H0).norm() / H1.norm();
      if(rel_norm < rel_norm_thre)
      {
        riccati_converged_ = true;
        break;
      }
      A0 = A1;
      G0 = G1;
      H0 = H1;
    }
    P_ = H1;
    const Matrix S = (R + B.transpose() * P_ * B).inverse();
    riccati_error_ = (P_ - (A.transpose() * P_ * A + C.transpose() * Q * C - A.transpose() * P_ * B * S * B.transpose() * P_ * A)).norm();

    // 2. gains (reference :152-171)
    K_ = S * B.transpose() * P_ * A;
    const int p = C.rows();
    F_ = Matrix(B.cols(), horizon_steps_ * p);
    const Matrix A_BK_T = (A - B * K_).transpose();
    Matrix f_sub = Matrix::Identity(n);
    for(int i = 0; i < horizon_steps_; i++)
    {
      const Matrix blk = i < horizon_steps_ - 1 ? S * B.transpose() * f_sub * C.transpose() * Q
                                                : S * B.transpose() * f_sub * P_ * C.transpose();
      F_.setBlock(0, i * p, blk);
      f_sub = f_sub * A_BK_T;
    }
  }

public:
  std::shared_ptr<StateSpaceModel> model_;
  double horizon_dt_ = 0;
  int horizon_steps_ = 0;
  //! Solution of the algebraic Riccati equation
  Matrix P_;
  //! Feedback gain
  Matrix K_;
  //! Preview gain
  Matrix F_;
  double riccati_error_ = 0;
  bool riccati_converged_ = false;
};
} // namespace CCC
