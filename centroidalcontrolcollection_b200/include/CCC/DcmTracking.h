/* CCC/DcmTracking.h — drop-in host class for CCC::DcmTracking (Englsberger et al. 2013) on the C-ABI engine.
 *
 * Mirrors reference include/CCC/DcmTracking.h: RefData (:24-37: current_zmp, time_zmp_list), InitialParam = current
 * DCM (:43), constructor (:50-53), planOnce (:64, src/DcmTracking.cpp:7-48), public feedback_gain_ (:68).
 * Eigen is absent: Vector2d = std::array<double, 2>.  planOnce is a batch of one through ccc_dcm_tracking_plan;
 * new: planBatch — B initial DCMs over P reference-data records.  The reference's std::runtime_error for a switching
 * time in the past is raised from the engine's CCC_ERR_INVALID.  Header-only; no CPU fallback.
 */
#pragma once
#include <array>
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/ccc_b200.h"
#include "Gravity.h"

namespace CCC
{
class DcmTracking
{
public:
  using Vector2d = std::array<double, 2>;

  struct RefData
  {
    //! Current ZMP [m]
    Vector2d current_zmp = {0, 0};
    //! List of pairs of future ZMP and switching time
    std::map<double, Vector2d> time_zmp_list;
  };

  using InitialParam = Vector2d;

public:
  DcmTracking(double com_height, double feedback_gain = 2.0) : feedback_gain_(feedback_gain), omega_(std::sqrt(constants::g / com_height)) {}

  Vector2d planOnce(const RefData & ref_data, const InitialParam & initial_param, double current_time) const
  {
    return planBatch({ref_data}, {current_time}, {initial_param}, {0})[0];
  }

  /** planOnce for initial_params[b] on ref_data[plan_id[b]] (sampled at current_times[plan_id[b]]). */
  std::vector<Vector2d> planBatch(const std::vector<RefData> & ref_data,
                                  const std::vector<double> & current_times,
                                  const std::vector<InitialParam> & initial_params,
                                  const std::vector<int> & plan_id) const
  {
    const int P = static_cast<int>(ref_data.size()), B = static_cast<int>(initial_params.size());
    if(P == 0 || B == 0 || current_times.size() != ref_data.size() || plan_id.size() != initial_params.size())
      throw std::invalid_argument("[DcmTracking] planBatch: sizes");
    int K = 1;
    for(const auto & r : ref_data) K = std::max(K, static_cast<int>(r.time_zmp_list.size()));
    std::vector<double> cz(static_cast<size_t>(P) * 2), kt(static_cast<size_t>(P) * K, 0.0), kz(static_cast<size_t>(P) * K * 2, 0.0), dcm(static_cast<size_t>(B) * 2);
    std::vector<int32_t> nk(P), pid(plan_id.begin(), plan_id.end());
    for(int p = 0; p < P; p++)
    {
      cz[2 * p] = ref_data[p].current_zmp[0];
      cz[2 * p + 1] = ref_data[p].current_zmp[1];
      int i = 0;
      for(const auto & kv : ref_data[p].time_zmp_list)
      {
        kt[static_cast<size_t>(p) * K + i] = kv.first;
        kz[(static_cast<size_t>(p) * K + i) * 2] = kv.second[0];
        kz[(static_cast<size_t>(p) * K + i) * 2 + 1] = kv.second[1];
        i++;
      }
      nk[p] = i;
    }
    for(int b = 0; b < B; b++)
    {
      dcm[2 * b] = initial_params[b][0];
      dcm[2 * b + 1] = initial_params[b][1];
    }
    ccc_dcm_tracking_batch_t bt{};
    bt.batch = B;
    bt.n_plans = P;
    bt.max_knots = K;
    bt.omega = omega_;
    bt.feedback_gain = feedback_gain_;
    bt.plan_id = pid.data();
    bt.dcm = dcm.data();
    bt.current_time = current_times.data();
    bt.current_zmp = cz.data();
    bt.n_knots = nk.data();
    bt.knot_time = kt.data();
    bt.knot_zmp = kz.data();
    std::vector<double> out(static_cast<size_t>(B) * 2);
    if(ccc_dcm_tracking_plan(&bt, out.data(), CCC_MEM_HOST, nullptr) != CCC_OK) throw std::runtime_error(std::string("[DcmTracking] ") + ccc_last_error());
    std::vector<Vector2d> res(B);
    for(int b = 0; b < B; b++) res[b] = {out[2 * b], out[2 * b + 1]};
    return res;
  }

public:
  //! Feedback gain to calculate control ZMP
  double feedback_gain_ = 0;

protected:
  //! Time constant for inverted pendulum dynamics
  double omega_ = 0;
};
} // namespace CCC
