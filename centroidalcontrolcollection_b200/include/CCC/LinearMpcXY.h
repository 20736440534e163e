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
 * planOnce samples the two callbacks on the horizon grid (src :104-110) and hands the flat stage tables to
 * ccc_linear_mpc_xy_solve: stage models, closed-form discretisation, condensing, B_seq' W B_seq (FP64 tensor
 * cores), the QP vectors and the QP all run on the device.
 * New: planBatch() — a batch of initial states sharing one sampled schedule; planSweep() — initial states over a
 * sweep of schedules, one QP factorisation per schedule.  Header-only; link with libccc_b200.so; no CPU fallback.
 */
#pragma once
#include <array>
#include <functional>
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

#include <cstdio>
#include <string>

#include "Gravity.h"
#include "Contact.h"
#include "StateSpaceModel.h"
#include "detail/QpEngine.h"
#include "detail/RidgeTables.h"

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

  ~LinearMpcXY()
  {
    if(ws_) ccc_linear_mpc_xy_destroy(ws_);
  }
  LinearMpcXY(const LinearMpcXY &) = delete;
  LinearMpcXY & operator=(const LinearMpcXY &) = delete;

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
    return planSweep({Schedule{motion_param_func, ref_data_func, current_time}}, initial_params,
                     std::vector<int>(initial_params.size(), 0));
  }

  /** One contact / reference schedule of a sweep: the callbacks of planOnce and the time they are sampled from. */
  struct Schedule
  {
    std::function<MotionParam(double)> motion_param_func;
    std::function<RefData(double)> ref_data_func;
    double current_time = 0;
  };

  /** planOnce for initial_params[b] on schedules[sched_id[b]], all in one call: the callbacks are sampled on the
   *  horizon grid here (src/LinearMpcXY.cpp:104-110); stage models, discretisation, condensing, B_seq' W B_seq, the QP
   *  vectors and the QPs (one factorisation per schedule) run on the device (ccc_linear_mpc_xy_solve).  The
   *  schedules of one call must agree in total input dimension and number of contact stages. */
  std::vector<VectorXd> planSweep(const std::vector<Schedule> & schedules,
                                  const std::vector<InitialParam> & initial_params,
                                  const std::vector<int> & sched_id)
  {
    const int S = static_cast<int>(schedules.size()), B = static_cast<int>(initial_params.size()), N = horizon_steps_;
    if(S == 0 || B == 0 || sched_id.size() != initial_params.size()) throw std::invalid_argument("[LinearMpcXY] empty sweep");
    tables_.reset(S, N);
    std::vector<double> com_z(static_cast<size_t>(S) * N), fz(static_cast<size_t>(S) * N), ref(static_cast<size_t>(S) * N * 6);
    int n = 0, n_eq = 0;
    for(int s = 0; s < S; s++)
    {
      int ns = 0, es = 0;
      for(int i = 0; i < N; i++)
      {
        const double t = schedules[s].current_time + i * horizon_dt_;
        const MotionParam mp = schedules[s].motion_param_func(t);
        const int mk = tables_.setStage(s, i, mp.contact_list);
        ns += mk;
        es += mk > 0 ? 1 : 0; // no total_force_z constraint on stages without contact (src :128-132)
        com_z[static_cast<size_t>(s) * N + i] = mp.com_z;
        fz[static_cast<size_t>(s) * N + i] = mp.total_force_z;
        const StateDimVector out = schedules[s].ref_data_func(t).toOutput(mass_);
        for(int j = 0; j < 6; j++) ref[(static_cast<size_t>(s) * N + i) * 6 + j] = out[j];
      }
      if(s == 0)
      {
        n = ns;
        n_eq = es;
      }
      else if(ns != n || es != n_eq)
        throw std::invalid_argument("[LinearMpcXY] the schedules of one sweep must agree in total input dimension and contact stages");
    }
    if(n == 0) throw std::runtime_error("[LinearMpcXY] no contact in the whole horizon");
    if(!ws_ || n != ws_n_ || n_eq != ws_eq_ || B > ws_batch_ || S > ws_sched_)
    {
      if(ws_) ccc_linear_mpc_xy_destroy(ws_);
      ws_batch_ = B > ws_batch_ ? B : ws_batch_;
      ws_sched_ = S > ws_sched_ ? S : ws_sched_;
      ws_n_ = n;
      ws_eq_ = n_eq;
      ws_ = ccc_linear_mpc_xy_create(N, n, n_eq, ws_batch_, ws_sched_);
      if(!ws_) throw std::runtime_error(std::string("[LinearMpcXY] ") + ccc_last_error());
    }
    std::vector<double> x0(static_cast<size_t>(B) * 6);
    for(int b = 0; b < B; b++)
    {
      const StateDimVector x = initial_params[b].toState(mass_);
      for(int j = 0; j < 6; j++) x0[static_cast<size_t>(b) * 6 + j] = x[j];
    }
    std::vector<int32_t> sid(sched_id.begin(), sched_id.end());
    ccc_linear_mpc_xy_batch_t bt{};
    bt.horizon_steps = N;
    bt.batch = B;
    bt.n_sched = S;
    bt.m_max = tables_.M;
    bt.dt = horizon_dt_;
    bt.mass = mass_;
    bt.sched_id = sid.data();
    bt.m = tables_.m.data();
    bt.ridge = tables_.ridge.data();
    bt.vertex = tables_.vertex.data();
    bt.com_z = com_z.data();
    bt.total_force_z = fz.data();
    bt.ref_output = ref.data();
    const VectorXd w = weight_param_.outputWeight(1);
    for(int j = 0; j < 6; j++) bt.w_output[j] = w[j];
    bt.w_force = weight_param_.force;
    bt.force_lo = force_range_.first;
    bt.force_hi = force_range_.second;
    bt.x0 = x0.data();
    u_.assign(static_cast<size_t>(B) * n, 0.0);
    status_.assign(B, 0);
    iters_.assign(B, 0);
    ccc_linear_mpc_xy_result_t rs{};
    rs.u = u_.data();
    rs.status = status_.data();
    rs.iters = iters_.data();
    if(ccc_linear_mpc_xy_solve(ws_, &bt, &rs, CCC_MEM_HOST, nullptr) != CCC_OK)
      throw std::runtime_error(std::string("[LinearMpcXY] ") + ccc_last_error());
    // a failed QP is reported the way QpSolverCollection does it (a message, the last iterate is returned)
    failed_ = 0;
    for(int b = 0; b < B; b++)
      if(status_[b] != 0) failed_++;
    if(failed_ > 0) std::fprintf(stderr, "[LinearMpcXY] %d of %d QPs failed to solve (lastStatus())\n", failed_, B);
    std::vector<VectorXd> out(B);
    for(int b = 0; b < B; b++)
    {
      const int m0 = tables_.inputDim(sid[b], 0);
      out[b].assign(u_.begin() + static_cast<size_t>(b) * n, u_.begin() + static_cast<size_t>(b) * n + m0);
    }
    return out;
  }

  int lastStatus(int b = 0) const { return status_.at(b); }
  int lastIter(int b = 0) const { return iters_.at(b); }
  //! QPs of the last call that did not reach status 0
  int lastFailed() const { return failed_; }

public:
  double mass_ = 0;
  double horizon_dt_ = 0;
  int horizon_steps_ = 0;
  WeightParam weight_param_;
  //! Min/max ridge force [N]
  std::pair<double, double> force_range_;

protected:
  detail::RidgeTables tables_;
  ccc_linear_mpc_xy_ws_t * ws_ = nullptr;
  int ws_n_ = 0, ws_eq_ = 0, ws_batch_ = 0, ws_sched_ = 0, failed_ = 0;
  std::vector<double> u_;
  std::vector<int32_t> status_, iters_;
};
} // namespace CCC
