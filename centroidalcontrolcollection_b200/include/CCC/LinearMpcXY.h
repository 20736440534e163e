/* CCC/LinearMpcXY.h — drop-in host class for CCC::LinearMpcXY (Audren et al. 2014 / Nagasaka et al. 2012:
 * linear MPC of the horizontal linear and angular momentum for a predefined vertical motion and contact
 * sequence, ridge force scales of every stage as decision variables) on top of the C-ABI QP engine.
 *
 * Mirrors reference include/CCC/LinearMpcXY.h and src/LinearMpcXY.cpp: MotionParam (:38-51), InitialParam
 * (:54-69, toState src :26-31), RefData (:72-101, toOutput src :33-38), WeightParam (:104-143, outputWeight
 * src :45-57), Model (:149-166, src :59-83), constructor (:177-181, src :85-94, force_range_ = (3, 3 m g)),
 * planOnce (:224-227, src :96-115), procOnce (src :117-182).  The box x_min <= x <= x_max of the
 * reference's QpCoeff enters the engine as 2n inequality rows (include/ccc_b200.h, QP section).
 * Eigen is absent: Vector2d = std::array<double,2>, VectorXd = std::vector<double>.
 * New: planBatch() — a batch of initial states sharing the sampled contact / reference schedule (the
 * condensing and the QP matrices are built once per call).  Header-only; link with libccc_b200.so; no CPU
 * fallback.
 */
#pragma once
#include <array>
#include <functional>
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

#include "Gravity.h"
#include "Contact.h"
#include "VariantSequentialExtension.h"
#include "detail/QpEngine.h"

namespace CCC
{
class LinearMpcXY
{
public:
  static constexpr int state_dim_ = 6;
  using Vector2d = std::array<double, 2>;
  using VectorXd = std::vector<double>;
  using StateDimVector = std::array<double, 6>;

  struct MotionParam
  {
    //! CoM height [m]
    double com_z = 0;
    //! Total vertical force [N]
    double total_force_z = 0;
    std::vector<std::shared_ptr<ForceColl::Contact>> contact_list;
  };

  struct InitialParam
  {
    Vector2d pos = {0, 0};
    Vector2d vel = {0, 0};
    Vector2d angular_momentum = {0, 0};
    StateDimVector toState(double mass) const
    {
      return {mass * pos[0], mass * vel[0], mass * pos[1], mass * vel[1], angular_momentum[0], angular_momentum[1]};
    }
  };

  struct RefData
  {
    Vector2d pos = {0, 0};
    Vector2d vel = {0, 0};
    Vector2d angular_momentum = {0, 0};
    static constexpr int outputDim() { return 6; }
    StateDimVector toOutput(double mass) const
    {
      return {mass * pos[0], mass * vel[0], mass * pos[1], mass * vel[1], angular_momentum[0], angular_momentum[1]};
    }
  };

  struct WeightParam
  {
    Vector2d linear_momentum_integral, linear_momentum, angular_momentum;
    double force;
    WeightParam(const Vector2d & _linear_momentum_integral = {1.0, 1.0},
                const Vector2d & _linear_momentum = {0.0, 0.0},
                const Vector2d & _angular_momentum = {1.0, 1.0},
                double _force = 1e-5)
    : linear_momentum_integral(_linear_momentum_integral), linear_momentum(_linear_momentum),
      angular_momentum(_angular_momentum), force(_force)
    {
    }
    VectorXd inputWeight(int total_input_dim) const { return VectorXd(total_input_dim, force); }
    VectorXd outputWeight(size_t seq_len) const
    {
      const StateDimVector one = {linear_momentum_integral[0], linear_momentum[0], linear_momentum_integral[1],
                                  linear_momentum[1], angular_momentum[0], angular_momentum[1]};
      VectorXd w(6 * seq_len);
      for(size_t i = 0; i < seq_len; i++)
        for(int j = 0; j < 6; j++) w[6 * i + j] = one[j];
      return w;
    }
  };

  /** State (m c_x, m v_x, m c_y, m v_y, L_x, L_y), input = ridge force scales of the stage's contacts. */
  class Model : public StateSpaceModel
  {
  public:
    Model(double mass, const MotionParam & motion_param, int output_dim = 0)
    : StateSpaceModel(LinearMpcXY::state_dim_, totalRidgeNum(motion_param.contact_list), output_dim), motion_param_(motion_param)
    {
      A_(0, 1) = 1;
      A_(2, 3) = 1;
      A_(4, 2) = -1 * motion_param_.total_force_z / mass;
      A_(5, 0) = motion_param_.total_force_z / mass;
      int ridge_idx = 0;
      for(const auto & contact : motion_param_.contact_list)
        for(const auto & vr : contact->vertexWithRidgeList_)
          for(const auto & ridge : vr.ridgeList)
          {
            const auto & vertex = vr.vertex;
            B_(1, ridge_idx) = ridge[0];
            B_(3, ridge_idx) = ridge[1];
            B_(4, ridge_idx) = -1 * (vertex[2] - motion_param_.com_z) * ridge[1] + vertex[1] * ridge[2];
            B_(5, ridge_idx) = (vertex[2] - motion_param_.com_z) * ridge[0] + -1 * vertex[0] * ridge[2];
            ridge_idx++;
          }
    }
    static int totalRidgeNum(const std::vector<std::shared_ptr<ForceColl::Contact>> & contact_list)
    {
      int n = 0;
      for(const auto & contact : contact_list) n += contact->ridgeNum();
      return n;
    }
    MotionParam motion_param_;
  };

public:
  LinearMpcXY(double mass,
              double horizon_dt,
              int horizon_steps,
              const WeightParam & weight_param = WeightParam(),
              QpSolverCollection::QpSolverType = QpSolverCollection::QpSolverType::Any)
  : mass_(mass), horizon_dt_(horizon_dt), horizon_steps_(horizon_steps), weight_param_(weight_param),
    force_range_(3.0, 3.0 * mass * constants::g)
  {
  }

  /** Plan one step: planned force scales of the first stage. */
  VectorXd planOnce(const std::function<MotionParam(double)> & motion_param_func,
                    const std::function<RefData(double)> & ref_data_func,
                    const InitialParam & initial_param,
                    double current_time)
  {
    return planBatch(motion_param_func, ref_data_func, {initial_param}, current_time)[0];
  }

  /** Batched planOnce for initial states sharing one contact / reference schedule. */
  std::vector<VectorXd> planBatch(const std::function<MotionParam(double)> & motion_param_func,
                                  const std::function<RefData(double)> & ref_data_func,
                                  const std::vector<InitialParam> & initial_params,
                                  double current_time)
  {
    std::vector<std::shared_ptr<StateSpaceModel>> model_list(horizon_steps_);
    VectorXd ref_output_seq(static_cast<size_t>(horizon_steps_) * RefData::outputDim());
    for(int i = 0; i < horizon_steps_; i++)
    {
      const double t = current_time + i * horizon_dt_;
      model_list[i] = std::make_shared<Model>(mass_, motion_param_func(t));
      model_list[i]->calcDiscMatrix(horizon_dt_);
      const StateDimVector out = ref_data_func(t).toOutput(mass_);
      for(int j = 0; j < 6; j++) ref_output_seq[static_cast<size_t>(i) * 6 + j] = out[j];
    }
    std::vector<StateDimVector> xs(initial_params.size());
    for(size_t b = 0; b < initial_params.size(); b++) xs[b] = initial_params[b].toState(mass_);
    return procBatch(model_list, xs, ref_output_seq);
  }

  int lastStatus(int b = 0) const { return qp_.status(b); }
  int lastIter(int b = 0) const { return qp_.iters(b); }

protected:
  std::vector<VectorXd> procBatch(const std::vector<std::shared_ptr<StateSpaceModel>> & model_list,
                                  const std::vector<StateDimVector> & current_xs,
                                  const VectorXd & ref_output_seq)
  {
    VariantSequentialExtension seq_ext(model_list, false);
    const int n = seq_ext.totalInputDim(), rows = seq_ext.totalStateDim(), B = static_cast<int>(current_xs.size());
    if(n == 0) throw std::runtime_error("[LinearMpcXY] no contact in the whole horizon");
    int dim_eq = 0;
    for(const auto & model : model_list)
      if(model->inputDim() > 0) dim_eq++; // no total_force_z constraint on stages without contact
    const VectorXd output_weight = weight_param_.outputWeight(model_list.size());
    const detail::Matrix & Bs = seq_ext.B_seq_;
    // obj_mat = B_seq' W B_seq + w_force I
    detail::Matrix WB(rows, n);
    for(int r = 0; r < rows; r++)
      for(int j = 0; j < n; j++) WB(r, j) = output_weight[r] * Bs(r, j);
    const detail::Matrix BtW = WB.transpose();
    detail::Matrix Q = BtW * Bs;
    const VectorXd input_weight = weight_param_.inputWeight(n);
    for(int j = 0; j < n; j++) Q(j, j) += input_weight[j];
    // equalities: total vertical force of every contact stage
    detail::Matrix A(dim_eq, n);
    VectorXd eq_vec(dim_eq, 0.0);
    int accum_eq_dim = 0, accum_input_dim = 0;
    for(const auto & _model : model_list)
    {
      const auto model = std::dynamic_pointer_cast<Model>(_model);
      if(!model) throw std::runtime_error("[LinearMpcXY] model_list must hold LinearMpcXY::Model");
      if(model->inputDim() == 0) continue;
      int ridge_idx = 0;
      for(const auto & contact : model->motion_param_.contact_list)
        for(const auto & vr : contact->vertexWithRidgeList_)
          for(const auto & ridge : vr.ridgeList)
          {
            A(accum_eq_dim, accum_input_dim + ridge_idx) = ridge[2];
            ridge_idx++;
          }
      eq_vec[accum_eq_dim] = model->motion_param_.total_force_z;
      accum_eq_dim++;
      accum_input_dim += model->inputDim();
    }
    // x_min <= x <= x_max as inequality rows: -x <= -x_min, x <= x_max
    detail::Matrix C(2 * n, n);
    for(int j = 0; j < n; j++)
    {
      C(j, j) = -1.0;
      C(n + j, j) = 1.0;
    }
    qp_.setup(Q, A, C);
    qp_.resize(B, true);
    for(int b = 0; b < B; b++)
    {
      // obj_vec = -B_seq' W (ref_output_seq - A_seq x - E_seq)
      VectorXd resid(rows);
      for(int r = 0; r < rows; r++)
      {
        double ax = 0;
        for(int c = 0; c < 6; c++) ax += seq_ext.A_seq_(r, c) * current_xs[b][c];
        resid[r] = ref_output_seq[r] - ax - seq_ext.E_seq_[r];
      }
      double * c = qp_.objVec(b);
      for(int j = 0; j < n; j++)
      {
        double s = 0;
        for(int r = 0; r < rows; r++) s += BtW(j, r) * resid[r];
        c[j] = -1 * s;
      }
      for(int e = 0; e < dim_eq; e++) qp_.eqVec(b)[e] = eq_vec[e];
      double * d = qp_.ineqVec(b);
      for(int j = 0; j < n; j++)
      {
        d[j] = -force_range_.first;
        d[n + j] = force_range_.second;
      }
    }
    qp_.solve();
    const int m0 = model_list[0]->inputDim();
    std::vector<VectorXd> out(B);
    for(int b = 0; b < B; b++) out[b].assign(qp_.x(b), qp_.x(b) + m0);
    return out;
  }

public:
  double mass_ = 0;
  double horizon_dt_ = 0;
  int horizon_steps_ = 0;
  WeightParam weight_param_;
  //! Min/max ridge force [N]
  std::pair<double, double> force_range_;

protected:
  detail::QpEngine qp_;
};
} // namespace CCC
