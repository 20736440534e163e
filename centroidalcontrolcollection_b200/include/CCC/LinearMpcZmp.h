/* CCC/LinearMpcZmp.h — drop-in host classes for CCC::LinearMpcZmp1d / CCC::LinearMpcZmp (Wieber 2006:
 * minimal-jerk sequence that keeps the ZMP inside its limits over the horizon) on top of the C-ABI QP engine.
 *
 * Mirrors reference include/CCC/LinearMpcZmp.h and src/LinearMpcZmp.cpp: RefData / InitialParam of the 1-D
 * class (:27-34), constructor (:44-48, src :9-28: horizon_steps = ceil(duration / dt), Q = I, C = [-B_seq; B_seq]),
 * planOnce / procOnce (src :30-81), the 2-D class with Vector2d data (:104-160, src :83-112).
 * Eigen is absent: Vector2d = std::array<double,2>, the 1-D InitialParam (Vector3d pos/vel/acc) = std::array<double,3>.
 * New: procBatch / planBatch — many problems in one engine call; planOnce itself sends both axes as one
 * batch of two QPs.  Header-only; link with libccc_b200.so; no CPU fallback.
 */
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <functional>
#include <memory>
#include <stdexcept>
#include <vector>

#include "CommonModels.h"
#include "InvariantSequentialExtension.h"
#include "detail/QpEngine.h"

namespace CCC
{
class LinearMpcZmp1d
{
public:
  struct RefData
  {
    //! Min/max limits of ZMP [m]
    std::array<double, 2> zmp_limits = {0, 0};
  };
  /** (CoM position, velocity, acceleration) */
  using InitialParam = std::array<double, 3>;

  LinearMpcZmp1d(double com_height,
                 double horizon_duration,
                 double horizon_dt,
                 QpSolverCollection::QpSolverType = QpSolverCollection::QpSolverType::Any)
  : horizon_dt_(horizon_dt), horizon_steps_(static_cast<int>(std::ceil(horizon_duration / horizon_dt))),
    model_(std::make_shared<ComZmpModelJerkInput>(com_height))
  {
    model_->calcDiscMatrix(horizon_dt_);
    seq_ext_ = std::make_shared<InvariantSequentialExtension>(model_, horizon_steps_, true);
    const int N = horizon_steps_;
    detail::Matrix C(2 * N, N);
    for(int i = 0; i < N; i++)
      for(int j = 0; j < N; j++)
      {
        C(i, j) = -1 * seq_ext_->B_seq_(i, j);
        C(N + i, j) = seq_ext_->B_seq_(i, j);
      }
    qp_.setup(detail::Matrix::Identity(N), detail::Matrix(0, N), C);
  }

  double planOnce(const std::function<RefData(double)> & ref_data_func,
                  const InitialParam & initial_param,
                  double current_time,
                  double control_dt = -1)
  {
    std::vector<RefData> ref_data_seq(horizon_steps_);
    for(int i = 0; i < horizon_steps_; i++) ref_data_seq[i] = ref_data_func(current_time + i * horizon_dt_);
    return procOnce(ref_data_seq, initial_param, current_time, control_dt);
  }

  double procOnce(const std::vector<RefData> & ref_data_seq, const InitialParam & initial_param, double current_time, double control_dt)
  {
    return procBatch({&ref_data_seq}, {initial_param}, {0}, current_time, control_dt)[0];
  }

  /** Problem b uses the limits sequence ref_data_seqs[seq_id[b]] and the initial state initial_params[b]. */
  std::vector<double> procBatch(const std::vector<const std::vector<RefData> *> & ref_data_seqs,
                                const std::vector<InitialParam> & initial_params,
                                const std::vector<int> & seq_id,
                                double, // current_time
                                double control_dt)
  {
    const int N = horizon_steps_, B = static_cast<int>(initial_params.size());
    qp_.resize(B, false);
    for(int b = 0; b < B; b++)
    {
      const auto & seq = *ref_data_seqs[seq_id[b]];
      const InitialParam & ip = initial_params[b];
      double * d = qp_.ineqVec(b);
      for(int i = 0; i < N; i++)
      {
        const double ax0 = seq_ext_->A_seq_(i, 0) * ip[0] + seq_ext_->A_seq_(i, 1) * ip[1] + seq_ext_->A_seq_(i, 2) * ip[2];
        d[i] = ax0 - seq[i].zmp_limits[0];
        d[N + i] = -1 * ax0 + seq[i].zmp_limits[1];
      }
    }
    qp_.solve();
    if(control_dt < 0) control_dt = horizon_dt_;
    std::vector<double> zmp(B);
    for(int b = 0; b < B; b++)
    {
      const auto & lim = (*ref_data_seqs[seq_id[b]])[0].zmp_limits;
      const InitialParam & ip = initial_params[b];
      const double com_jerk = qp_.x(b)[0];
      const double com_acc = ip[2] + control_dt * com_jerk;
      const double com_pos = ip[0] + control_dt * ip[1] + 0.5 * std::pow(control_dt, 2) * ip[2];
      zmp[b] = std::clamp(com_pos + model_->C_(0, 2) * com_acc, lim[0], lim[1]);
    }
    return zmp;
  }

  /** QP outcome of problem b of the last call: 0 solved (include/ccc_b200.h ccc_qp_result_t::status). */
  int lastStatus(int b = 0) const { return qp_.status(b); }
  int lastIter(int b = 0) const { return qp_.iters(b); }

public:
  double horizon_dt_ = 0;
  int horizon_steps_ = 0;
  std::shared_ptr<ComZmpModelJerkInput> model_;
  std::shared_ptr<InvariantSequentialExtension> seq_ext_;

protected:
  detail::QpEngine qp_;
};

class LinearMpcZmp
{
public:
  using Vector2d = std::array<double, 2>;

  struct RefData
  {
    //! Min/max limits of ZMP [m]
    std::array<Vector2d, 2> zmp_limits = {Vector2d{0, 0}, Vector2d{0, 0}};
  };
  struct InitialParam
  {
    Vector2d pos = {0, 0};
    Vector2d vel = {0, 0};
    Vector2d acc = {0, 0};
  };
  struct BatchItem
  {
    int schedule = 0;
    InitialParam initial_param;
  };

  LinearMpcZmp(double com_height,
               double horizon_duration,
               double horizon_dt,
               QpSolverCollection::QpSolverType qp_solver_type = QpSolverCollection::QpSolverType::Any)
  : mpc_1d_(std::make_shared<LinearMpcZmp1d>(com_height, horizon_duration, horizon_dt, qp_solver_type))
  {
  }

  Vector2d planOnce(const std::function<RefData(double)> & ref_data_func,
                    const InitialParam & initial_param,
                    double current_time,
                    double control_dt = -1)
  {
    BatchItem item;
    item.initial_param = initial_param;
    return planBatch({ref_data_func}, {item}, current_time, control_dt)[0];
  }

  /** Batched planOnce: items[b] follows the limits schedule ref_data_funcs[items[b].schedule].  Both axes
   *  of every problem go to the engine as one batch of 2B one-dimensional QPs. */
  std::vector<Vector2d> planBatch(const std::vector<std::function<RefData(double)>> & ref_data_funcs,
                                  const std::vector<BatchItem> & items,
                                  double current_time,
                                  double control_dt = -1)
  {
    const int N = mpc_1d_->horizon_steps_, S = static_cast<int>(ref_data_funcs.size()), B = static_cast<int>(items.size());
    if(S == 0 || B == 0) throw std::runtime_error("planBatch: empty schedule list or batch");
    for(const auto & item : items)
      if(item.schedule < 0 || item.schedule >= S) throw std::runtime_error("planBatch: schedule index out of range");
    std::vector<std::vector<LinearMpcZmp1d::RefData>> seqs(2 * S, std::vector<LinearMpcZmp1d::RefData>(N));
    for(int s = 0; s < S; s++)
      for(int i = 0; i < N; i++)
      {
        const RefData rd = ref_data_funcs[s](current_time + i * mpc_1d_->horizon_dt_);
        for(int j = 0; j < 2; j++)
        {
          seqs[s][i].zmp_limits[j] = rd.zmp_limits[j][0];
          seqs[S + s][i].zmp_limits[j] = rd.zmp_limits[j][1];
        }
      }
    std::vector<const std::vector<LinearMpcZmp1d::RefData> *> seq_ptrs(2 * S);
    for(int s = 0; s < 2 * S; s++) seq_ptrs[s] = &seqs[s];
    std::vector<LinearMpcZmp1d::InitialParam> ips(2 * B);
    std::vector<int> seq_id(2 * B);
    for(int b = 0; b < B; b++)
      for(int a = 0; a < 2; a++)
      {
        const InitialParam & ip = items[b].initial_param;
        ips[a * B + b] = {ip.pos[a], ip.vel[a], ip.acc[a]};
        seq_id[a * B + b] = a * S + items[b].schedule;
      }
    const std::vector<double> z = mpc_1d_->procBatch(seq_ptrs, ips, seq_id, current_time, control_dt);
    std::vector<Vector2d> out(B);
    for(int b = 0; b < B; b++) out[b] = {z[b], z[B + b]};
    return out;
  }

public:
  std::shared_ptr<LinearMpcZmp1d> mpc_1d_;
};
} // namespace CCC
