/* CCC/IntrinsicallyStableMpc.h — drop-in host classes for CCC::IntrinsicallyStableMpc1d /
 * CCC::IntrinsicallyStableMpc (Scianca et al. 2016) on top of the C-ABI QP engine.
 *
 * Mirrors reference include/CCC/IntrinsicallyStableMpc.h and src/IntrinsicallyStableMpc.cpp: RefData /
 * InitialParam / WeightParam (:22-55), constructor (:65-69, src :8-45: P lower-triangular dt, Q = w_vel I +
 * w_zmp P'P, stability equality (14), C = [-P; P]), planOnce / procOnce (src :47-104), 2-D class (:126-194,
 * src :106-139).  Eigen is absent: Vector2d = std::array<double,2>.  New: procBatch / planBatch.
 * Header-only; link with libccc_b200.so; no CPU fallback.
 */
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <functional>
#include <memory>
#include <stdexcept>
#include <vector>

#include "Gravity.h"
#include "detail/QpEngine.h"

namespace CCC
{
class IntrinsicallyStableMpc1d
{
public:
  struct RefData
  {
    //! ZMP [m]
    double zmp = 0;
    //! Min/max limits of ZMP [m]
    std::array<double, 2> zmp_limits = {0, 0};
  };
  struct InitialParam
  {
    //! Current capture point [m]
    double capture_point = 0;
    //! Current ZMP planned in previous step [m]
    double planned_zmp = 0;
  };
  struct WeightParam
  {
    double zmp, zmp_vel;
    WeightParam(double _zmp = 1.0, double _zmp_vel = 1e-3) : zmp(_zmp), zmp_vel(_zmp_vel) {}
  };

  IntrinsicallyStableMpc1d(double com_height,
                           double horizon_duration,
                           double horizon_dt,
                           QpSolverCollection::QpSolverType = QpSolverCollection::QpSolverType::Any,
                           const WeightParam & weight_param = WeightParam())
  : weight_param_(weight_param), horizon_dt_(horizon_dt),
    horizon_steps_(static_cast<int>(std::ceil(horizon_duration / horizon_dt))),
    omega_(std::sqrt(constants::g / com_height)), lambda_(std::exp(-1 * omega_ * horizon_dt_))
  {
    const int N = horizon_steps_;
    detail::Matrix Q(N, N), A(1, N), C(2 * N, N);
    // P(i, j) = dt for j <= i, so (P'P)(i, j) = dt^2 (N - max(i, j))
    for(int i = 0; i < N; i++)
      for(int j = 0; j < N; j++)
      {
        double ptp = 0;
        for(int k = std::max(i, j); k < N; k++) ptp += horizon_dt_ * horizon_dt_;
        Q(i, j) = (i == j ? weight_param_.zmp_vel : 0.0) + weight_param_.zmp * ptp;
        if(j <= i)
        {
          C(i, j) = -1 * horizon_dt_;
          C(N + i, j) = horizon_dt_;
        }
      }
    A(0, 0) = (1 - lambda_) / (omega_ * (1 - std::pow(lambda_, N)));
    for(int i = 1; i < N; i++) A(0, i) = lambda_ * A(0, i - 1);
    qp_.setup(Q, A, C);
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

  std::vector<double> procBatch(const std::vector<const std::vector<RefData> *> & ref_data_seqs,
                                const std::vector<InitialParam> & initial_params,
                                const std::vector<int> & seq_id,
                                double, // current_time
                                double control_dt)
  {
    const int N = horizon_steps_, B = static_cast<int>(initial_params.size());
    qp_.resize(B, true);
    for(int b = 0; b < B; b++)
    {
      const auto & seq = *ref_data_seqs[seq_id[b]];
      const InitialParam & ip = initial_params[b];
      qp_.eqVec(b)[0] = ip.capture_point - ip.planned_zmp;
      double * c = qp_.objVec(b);
      double * d = qp_.ineqVec(b);
      // obj_vec = w_zmp P' (planned_zmp 1 - ref_zmp): a suffix sum times dt
      double suffix = 0;
      for(int i = N - 1; i >= 0; i--)
      {
        suffix += ip.planned_zmp - seq[i].zmp;
        c[i] = weight_param_.zmp * horizon_dt_ * suffix;
        d[i] = -1 * seq[i].zmp_limits[0] + ip.planned_zmp;
        d[N + i] = seq[i].zmp_limits[1] - ip.planned_zmp;
      }
    }
    qp_.solve();
    if(control_dt < 0) control_dt = horizon_dt_;
    std::vector<double> zmp(B);
    for(int b = 0; b < B; b++)
    {
      const auto & lim = (*ref_data_seqs[seq_id[b]])[0].zmp_limits;
      zmp[b] = std::clamp(initial_params[b].planned_zmp + control_dt * qp_.x(b)[0], lim[0], lim[1]);
    }
    return zmp;
  }

  int lastStatus(int b = 0) const { return qp_.status(b); }
  int lastIter(int b = 0) const { return qp_.iters(b); }

public:
  WeightParam weight_param_;
  double horizon_dt_ = 0;
  int horizon_steps_ = 0;
  double omega_ = 0;
  double lambda_ = 0;

protected:
  detail::QpEngine qp_;
};

class IntrinsicallyStableMpc
{
public:
  using Vector2d = std::array<double, 2>;
  using WeightParam = IntrinsicallyStableMpc1d::WeightParam;

  struct RefData
  {
    Vector2d zmp = {0, 0};
    std::array<Vector2d, 2> zmp_limits = {Vector2d{0, 0}, Vector2d{0, 0}};
  };
  struct InitialParam
  {
    Vector2d capture_point = {0, 0};
    Vector2d planned_zmp = {0, 0};
  };
  struct BatchItem
  {
    int schedule = 0;
    InitialParam initial_param;
  };

  IntrinsicallyStableMpc(double com_height,
                         double horizon_duration,
                         double horizon_dt,
                         QpSolverCollection::QpSolverType qp_solver_type = QpSolverCollection::QpSolverType::Any,
                         const WeightParam & weight_param = WeightParam())
  : mpc_1d_(std::make_shared<IntrinsicallyStableMpc1d>(com_height, horizon_duration, horizon_dt, qp_solver_type, weight_param))
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

  std::vector<Vector2d> planBatch(const std::vector<std::function<RefData(double)>> & ref_data_funcs,
                                  const std::vector<BatchItem> & items,
                                  double current_time,
                                  double control_dt = -1)
  {
    const int N = mpc_1d_->horizon_steps_, S = static_cast<int>(ref_data_funcs.size()), B = static_cast<int>(items.size());
    if(S == 0 || B == 0) throw std::runtime_error("planBatch: empty schedule list or batch");
    for(const auto & item : items)
      if(item.schedule < 0 || item.schedule >= S) throw std::runtime_error("planBatch: schedule index out of range");
    std::vector<std::vector<IntrinsicallyStableMpc1d::RefData>> seqs(2 * S, std::vector<IntrinsicallyStableMpc1d::RefData>(N));
    for(int s = 0; s < S; s++)
      for(int i = 0; i < N; i++)
      {
        const RefData rd = ref_data_funcs[s](current_time + i * mpc_1d_->horizon_dt_);
        for(int a = 0; a < 2; a++)
        {
          seqs[a * S + s][i].zmp = rd.zmp[a];
          for(int j = 0; j < 2; j++) seqs[a * S + s][i].zmp_limits[j] = rd.zmp_limits[j][a];
        }
      }
    std::vector<const std::vector<IntrinsicallyStableMpc1d::RefData> *> seq_ptrs(2 * S);
    for(int s = 0; s < 2 * S; s++) seq_ptrs[s] = &seqs[s];
    std::vector<IntrinsicallyStableMpc1d::InitialParam> ips(2 * B);
    std::vector<int> seq_id(2 * B);
    for(int b = 0; b < B; b++)
      for(int a = 0; a < 2; a++)
      {
        ips[a * B + b].capture_point = items[b].initial_param.capture_point[a];
        ips[a * B + b].planned_zmp = items[b].initial_param.planned_zmp[a];
        seq_id[a * B + b] = a * S + items[b].schedule;
      }
    const std::vector<double> z = mpc_1d_->procBatch(seq_ptrs, ips, seq_id, current_time, control_dt);
    std::vector<Vector2d> out(B);
    for(int b = 0; b < B; b++) out[b] = {z[b], z[B + b]};
    return out;
  }

public:
  std::shared_ptr<IntrinsicallyStableMpc1d> mpc_1d_;
};
} // namespace CCC
