/* CCC/SingularPreviewControlZmp.h — drop-in host classes for CCC::SingularPreviewControlZmp1d / CCC::SingularPreviewControlZmp
 * (Urata et al. 2011, singular LQ preview regulation) on the C-ABI engine.
 *
 * Mirrors reference include/CCC/SingularPreviewControlZmp.h: 1-D InitialParam (:24-34), constructor (:46-51), planOnce (:60-63,
 * src/SingularPreviewControlZmp.cpp:7-21, procOnce :23-57); 2-D InitialParam (:82-94), constructor (:102-106), planOnce (:114-117,
 * src :59-88).  Eigen is absent: Vector2d = std::array<double, 2>.  planOnce samples the reference ZMP on the horizon grid as
 * the reference does and runs a batch of one through ccc_singular_preview_plan; new: planBatch for B initial parameters on P
 * sampled reference sequences.  Header-only; link with libccc_b200.so; no CPU fallback.
 */
#pragma once
#include <array>
#include <cmath>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/ccc_b200.h"
#include "Gravity.h"

namespace CCC
{
class SingularPreviewControlZmp
{
public:
  using Vector2d = std::array<double, 2>;

  /** reference :82-94 */
  struct InitialParam
  {
    //! CoM position [m]
    Vector2d pos = {0, 0};
    //! CoM velocity [m/s]
    Vector2d vel = {0, 0};
    //! Current ZMP planned in previous step [m]
    Vector2d planned_zmp = {0, 0};
  };

  /** reference :102-106 (and the 1-D constructor :46-51) */
  SingularPreviewControlZmp(double com_height, double horizon_duration, double horizon_dt)
  : horizon_dt_(horizon_dt), horizon_steps_(static_cast<int>(std::ceil(horizon_duration / horizon_dt))),
    omega_(std::sqrt(constants::g / com_height))
  {
  }

  /** reference :114-117: planned ZMP; control_dt < 0 means horizon_dt. */
  Vector2d planOnce(const std::function<Vector2d(double)> & ref_zmp_func,
                    const InitialParam & initial_param,
                    double current_time,
                    double control_dt = -1) const
  {
    return planBatch({sample(ref_zmp_func, current_time)}, {initial_param}, {0}, control_dt)[0];
  }

  /** The reference's sampling of ref_zmp_func on the horizon grid (src/SingularPreviewControlZmp.cpp:66-71). */
  std::vector<Vector2d> sample(const std::function<Vector2d(double)> & ref_zmp_func, double current_time) const
  {
    std::vector<Vector2d> seq(static_cast<size_t>(horizon_steps_));
    for(int i = 0; i < horizon_steps_; i++) seq[static_cast<size_t>(i)] = ref_zmp_func(current_time + i * horizon_dt_);
    return seq;
  }

  /** planOnce for initial_params[b] on the sampled sequence ref_zmp_seqs[plan_id[b]]. */
  std::vector<Vector2d> planBatch(const std::vector<std::vector<Vector2d>> & ref_zmp_seqs,
                                  const std::vector<InitialParam> & initial_params,
                                  const std::vector<int> & plan_id,
                                  double control_dt = -1) const
  {
    const size_t P = ref_zmp_seqs.size(), B = initial_params.size(), N = static_cast<size_t>(horizon_steps_);
    if(P == 0 || B == 0 || plan_id.size() != B) throw std::invalid_argument("[SingularPreviewControlZmp] planBatch: sizes");
    std::vector<double> ref(P * N * 2), state(B * 6);
    for(size_t p = 0; p < P; p++)
    {
      if(ref_zmp_seqs[p].size() != N) throw std::invalid_argument("[SingularPreviewControlZmp] planBatch: sequence length");
      for(size_t i = 0; i < N; i++)
        for(size_t a = 0; a < 2; a++) ref[(p * N + i) * 2 + a] = ref_zmp_seqs[p][i][a];
    }
    std::vector<int32_t> pid(B);
    for(size_t b = 0; b < B; b++)
    {
      pid[b] = plan_id[b];
      for(size_t a = 0; a < 2; a++)
      {
        state[(b * 2 + a) * 3 + 0] = initial_params[b].planned_zmp[a];
        state[(b * 2 + a) * 3 + 1] = initial_params[b].pos[a];
        state[(b * 2 + a) * 3 + 2] = initial_params[b].vel[a];
      }
    }
    ccc_singular_preview_batch_t bt{};
    bt.batch = static_cast<int32_t>(B);
    bt.n_plans = static_cast<int32_t>(P);
    bt.horizon_steps = horizon_steps_;
    bt.omega = omega_;
    bt.horizon_dt = horizon_dt_;
    bt.control_dt = control_dt < 0 ? horizon_dt_ : control_dt;
    bt.plan_id = pid.data();
    bt.state = state.data();
    bt.ref_zmp = ref.data();
    std::vector<double> out(B * 2);
    if(ccc_singular_preview_plan(&bt, out.data(), CCC_MEM_HOST, nullptr) != CCC_OK)
      throw std::runtime_error(std::string("[SingularPreviewControlZmp] ") + ccc_last_error());
    std::vector<Vector2d> z(B);
    for(size_t b = 0; b < B; b++) z[b] = {out[2 * b], out[2 * b + 1]};
    return z;
  }

  //! Discretization timestep in horizon [sec]
  double horizon_dt_ = 0;
  //! Number of steps in horizon
  int horizon_steps_ = -1;
  //! Time constant for inverted pendulum dynamics
  double omega_ = 0;
};

/** 1-D controller (reference :17-76): the same engine call with the second axis left at zero. */
class SingularPreviewControlZmp1d
{
public:
  struct InitialParam
  {
    double pos = 0;
    double vel = 0;
    double planned_zmp = 0;
  };

  SingularPreviewControlZmp1d(double com_height, double horizon_duration, double horizon_dt) : spc_(com_height, horizon_duration, horizon_dt) {}

  double planOnce(const std::function<double(double)> & ref_zmp_func, const InitialParam & initial_param, double current_time, double control_dt = -1) const
  {
    SingularPreviewControlZmp::InitialParam ip;
    ip.pos = {initial_param.pos, 0.0};
    ip.vel = {initial_param.vel, 0.0};
    ip.planned_zmp = {initial_param.planned_zmp, 0.0};
    return spc_.planOnce([&](double t) { return SingularPreviewControlZmp::Vector2d{ref_zmp_func(t), 0.0}; }, ip, current_time, control_dt)[0];
  }

protected:
  SingularPreviewControlZmp spc_;
};
} // namespace CCC
