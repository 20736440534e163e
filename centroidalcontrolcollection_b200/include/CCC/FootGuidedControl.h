/* CCC/FootGuidedControl.h — drop-in host classes for CCC::FootGuidedControl1d / CCC::FootGuidedControl (Sugihara
 * 2017, Kojio 2019) on the C-ABI engine.
 *
 * Mirrors reference include/CCC/FootGuidedControl.h: 1-D RefData (:22-35) / InitialParam = capture point (:41),
 * constructor (:48-51), planOnce (:60, src/FootGuidedControl.cpp:11-69); 2-D RefData (:76-92), planOnce (:110,
 * src :71-92).  Eigen is absent: Vector2d = std::array<double, 2>.  planOnce is a batch of one through
 * ccc_foot_guided_plan; new: planBatch.  The reference's two std::runtime_error cases (negative transition duration,
 * transition end not in the future) are raised from the engine's CCC_ERR_INVALID.  Header-only; no CPU fallback.
 */
#pragma once
#include <array>
#include <cmath>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/ccc_b200.h"
#include "Gravity.h"

namespace CCC
{
class FootGuidedControl
{
public:
  using Vector2d = std::array<double, 2>;

  struct RefData
  {
    //! Transition start ZMP [m]
    Vector2d transit_start_zmp = {0, 0};
    //! Transition end ZMP [m]
    Vector2d transit_end_zmp = {0, 0};
    //! Transition start time [s]
    double transit_start_time = 0;
    //! Transition duration [s]
    double transit_duration = 0;
  };

  using InitialParam = Vector2d;

public:
  FootGuidedControl(double com_height) : omega_(std::sqrt(constants::g / com_height)) {}

  Vector2d planOnce(const RefData & ref_data, const InitialParam & initial_param, double current_time) const
  {
    return planBatch({ref_data}, {current_time}, {initial_param}, {0})[0];
  }

  /** planOnce for initial_params[b] on ref_data[plan_id[b]] (at current_times[plan_id[b]]). */
  std::vector<Vector2d> planBatch(const std::vector<RefData> & ref_data,
                                  const std::vector<double> & current_times,
                                  const std::vector<InitialParam> & initial_params,
                                  const std::vector<int> & plan_id) const
  {
    const int P = static_cast<int>(ref_data.size()), B = static_cast<int>(initial_params.size());
    if(P == 0 || B == 0 || current_times.size() != ref_data.size() || plan_id.size() != initial_params.size())
      throw std::invalid_argument("[FootGuidedControl] planBatch: sizes");
    std::vector<double> zs(static_cast<size_t>(P) * 2), ze(static_cast<size_t>(P) * 2), ts(P), td(P), cp(static_cast<size_t>(B) * 2);
    std::vector<int32_t> pid(plan_id.begin(), plan_id.end());
    for(int p = 0; p < P; p++)
    {
      for(int a = 0; a < 2; a++)
      {
        zs[2 * p + a] = ref_data[p].transit_start_zmp[a];
        ze[2 * p + a] = ref_data[p].transit_end_zmp[a];
      }
      ts[p] = ref_data[p].transit_start_time;
      td[p] = ref_data[p].transit_duration;
    }
    for(int b = 0; b < B; b++)
    {
      cp[2 * b] = initial_params[b][0];
      cp[2 * b + 1] = initial_params[b][1];
    }
    ccc_foot_guided_batch_t bt{};
    bt.batch = B;
    bt.n_plans = P;
    bt.omega = omega_;
    bt.plan_id = pid.data();
    bt.capture_point = cp.data();
    bt.current_time = current_times.data();
    bt.transit_start_zmp = zs.data();
    bt.transit_end_zmp = ze.data();
    bt.transit_start_time = ts.data();
    bt.transit_duration = td.data();
    std::vector<double> out(static_cast<size_t>(B) * 2);
    if(ccc_foot_guided_plan(&bt, out.data(), CCC_MEM_HOST, nullptr) != CCC_OK) throw std::runtime_error(std::string("[FootGuidedControl] ") + ccc_last_error());
    std::vector<Vector2d> res(B);
    for(int b = 0; b < B; b++) res[b] = {out[2 * b], out[2 * b + 1]};
    return res;
  }

protected:
  double omega_ = 0;
};

/** One-dimensional form (reference include/CCC/FootGuidedControl.h:17-64): the two-dimensional engine call with both
 *  axes carrying the same data. */
class FootGuidedControl1d
{
public:
  struct RefData
  {
    double transit_start_zmp = 0, transit_end_zmp = 0, transit_start_time = 0, transit_duration = 0;
  };
  using InitialParam = double;

  FootGuidedControl1d(double com_height) : fgc_(com_height) {}

  double planOnce(const RefData & ref_data, const InitialParam & initial_param, double current_time) const
  {
    FootGuidedControl::RefData rd;
    rd.transit_start_zmp = {ref_data.transit_start_zmp, ref_data.transit_start_zmp};
    rd.transit_end_zmp = {ref_data.transit_end_zmp, ref_data.transit_end_zmp};
    rd.transit_start_time = ref_data.transit_start_time;
    rd.transit_duration = ref_data.transit_duration;
    return fgc_.planOnce(rd, {initial_param, initial_param}, current_time)[0];
  }

protected:
  FootGuidedControl fgc_;
};
} // namespace CCC
