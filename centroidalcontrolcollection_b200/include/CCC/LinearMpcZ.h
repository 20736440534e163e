/* CCC/LinearMpcZ.h — drop-in host class for CCC::LinearMpcZ (linear MPC of the vertical CoM motion with the
 * contact-phase sequence given; decision variables = vertical force of every contact stage) on top of the
 * C-ABI QP engine.
 *
 * Mirrors reference include/CCC/LinearMpcZ.h and src/LinearMpcZ.cpp: InitialParam (pos, vel) (:30),
 * WeightParam (:33-47, defaults pos 1.0, force 1e-7), ModelContactPhase / ModelNoncontactPhase (:74-100,
 * src :10-24: state (m c_z, P_z), input f_z, output c_z, gravity as affine term), constructor (src :26-41,
 * force_range_ = (10, 10 m g)), planOnce (:152-155, src :43-66: zero force without contact at current_time),
 * procOnce (src :68-94: condensing for outputs, Q = w_pos B'B + w_force I, box bounds).
 * The box enters the engine as 2n inequality rows.  New: planBatch() — initial states sharing one contact /
 * reference schedule.  Header-only; link with libccc_b200.so; no CPU fallback.
 */
#pragma once
#include <array>
#include <functional>
#include <memory>
#include <utility>
#include <vector>

#include "Gravity.h"
#include "VariantSequentialExtension.h"
#include "detail/QpEngine.h"

namespace CCC
{
class LinearMpcZ
{
public:
  static constexpr int state_dim_ = 2;
  /** (CoM height, vertical velocity) */
  using InitialParam = std::array<double, 2>;

  struct WeightParam
  {
    double pos, force;
    WeightParam(double _pos = 1.0, double _force = 1e-7) : pos(_pos), force(_force) {}
  };

  class ModelContactPhase : public StateSpaceModel
  {
  public:
    explicit ModelContactPhase(double mass) : StateSpaceModel(LinearMpcZ::state_dim_, 1, 1)
    {
      A_(0, 1) = 1;
      B_(1, 0) = 1;
      C_(0, 0) = 1 / mass;
      E_ = {0, -1 * mass * constants::g};
    }
  };
  class ModelNoncontactPhase : public StateSpaceModel
  {
  public:
    explicit ModelNoncontactPhase(double mass) : StateSpaceModel(LinearMpcZ::state_dim_, 0, 1)
    {
      A_(0, 1) = 1;
      C_(0, 0) = 1 / mass;
      E_ = {0, -1 * mass * constants::g};
    }
  };

  LinearMpcZ(double mass,
             double horizon_dt,
             int horizon_steps,
             const WeightParam & weight_param = WeightParam(),
             QpSolverCollection::QpSolverType = QpSolverCollection::QpSolverType::Any)
  : mass_(mass), horizon_dt_(horizon_dt), horizon_steps_(horizon_steps), weight_param_(weight_param),
    force_range_(10.0, 10.0 * mass * constants::g)
  {
    model_contact_ = std::make_shared<ModelContactPhase>(mass_);
    model_noncontact_ = std::make_shared<ModelNoncontactPhase>(mass_);
    model_contact_->calcDiscMatrix(horizon_dt_);
    model_noncontact_->calcDiscMatrix(horizon_dt_);
  }

  /** Plan one step: planned vertical force. */
  double planOnce(const std::function<bool(double)> & contact_func,
                  const std::function<double(double)> & ref_pos_func,
                  const InitialParam & initial_param,
                  double current_time)
  {
    return planBatch(contact_func, ref_pos_func, {initial_param}, current_time)[0];
  }

  std::vector<double> planBatch(const std::function<bool(double)> & contact_func,
                                const std::function<double(double)> & ref_pos_func,
                                const std::vector<InitialParam> & initial_params,
                                double current_time)
  {
    const int B = static_cast<int>(initial_params.size());
    // planned force is always zero if there is no contact
    if(!contact_func(current_time)) return std::vector<double>(B, 0.0);
    std::vector<std::shared_ptr<StateSpaceModel>> model_list(horizon_steps_);
    std::vector<double> ref_pos_seq(horizon_steps_);
    for(int i = 0; i < horizon_steps_; i++)
    {
      const double t = current_time + i * horizon_dt_;
      model_list[i] = contact_func(t) ? std::static_pointer_cast<StateSpaceModel>(model_contact_)
                                      : std::static_pointer_cast<StateSpaceModel>(model_noncontact_);
      ref_pos_seq[i] = ref_pos_func(t);
    }
    VariantSequentialExtension seq_ext(model_list, true);
    const int n = seq_ext.totalInputDim(), rows = seq_ext.totalOutputDim();
    const detail::Matrix & Bs = seq_ext.B_seq_;
    const detail::Matrix Bt = Bs.transpose();
    detail::Matrix Q = (Bt * Bs) * weight_param_.pos;
    for(int j = 0; j < n; j++) Q(j, j) += weight_param_.force;
    detail::Matrix C(2 * n, n);
    for(int j = 0; j < n; j++)
    {
      C(j, j) = -1.0;
      C(n + j, j) = 1.0;
    }
    qp_.setup(Q, detail::Matrix(0, n), C);
    qp_.resize(B, true);
    for(int b = 0; b < B; b++)
    {
      const double x0[2] = {mass_ * initial_params[b][0], mass_ * initial_params[b][1]};
      std::vector<double> resid(rows);
      for(int r = 0; r < rows; r++)
        resid[r] = ref_pos_seq[r] - (seq_ext.A_seq_(r, 0) * x0[0] + seq_ext.A_seq_(r, 1) * x0[1]) - seq_ext.E_seq_[r];
      double * c = qp_.objVec(b);
      double * d = qp_.ineqVec(b);
      for(int j = 0; j < n; j++)
      {
        double s = 0;
        for(int r = 0; r < rows; r++) s += Bt(j, r) * resid[r];
        c[j] = -1 * weight_param_.pos * s;
        d[j] = -force_range_.first;
        d[n + j] = force_range_.second;
      }
    }
    qp_.solve();
    std::vector<double> force(B);
    for(int b = 0; b < B; b++) force[b] = qp_.x(b)[0];
    return force;
  }

  int lastStatus(int b = 0) const { return qp_.status(b); }

public:
  double mass_ = 0;
  double horizon_dt_ = 0;
  int horizon_steps_ = 0;
  WeightParam weight_param_;
  std::shared_ptr<ModelContactPhase> model_contact_;
  std::shared_ptr<ModelNoncontactPhase> model_noncontact_;
  //! Min/max vertical force [N]
  std::pair<double, double> force_range_;

protected:
  detail::QpEngine qp_;
};
} // namespace CCC
