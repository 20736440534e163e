/* CCC/StepMpc.h — drop-in host classes for CCC::StepMpc / CCC::StepMpc1d (Xin et al. 2019, step-to-step MPC with online
 * footstep adaptation) on the C-ABI engine.
 *
 * Mirrors reference include/CCC/StepMpc.h: RefData with its element list (:36-55, 2-D :155-178), PlannedData with the optional
 * next-foot ZMP (:58-68, :181-194), InitialParam (:75, :197-206), WeightParam and its defaults (:78-117), constructors (:124-127,
 * :214-217), planOnce (:135, :225; src/StepMpc.cpp:27-193, :195-247).  Eigen is absent: Vector2d = std::array<double, 2>.
 * planOnce is a batch of one through ccc_step_mpc_plan; new: planBatch.  Header-only; link with libccc_b200.so; no CPU fallback.
 */
#pragma once
#include <array>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/ccc_b200.h"

namespace CCC
{
class StepMpc1d
{
public:
  /** One support phase per element (reference :36-55): at least one; two double-support phases never follow each other. */
  struct RefData
  {
    struct Element
    {
      bool is_single_support = true; // single (true) or double (false) support
      double zmp = 0;                // reference ZMP of the phase [m]
      double end_time = 0;           // [sec]
    };
    std::vector<Element> element_list;
  };

  /** reference :58-68 */
  struct PlannedData
  {
    double current_zmp = 0;              // ZMP to apply now [m]
    std::optional<double> next_foot_zmp; // landing ZMP of the next foot [m]; empty if its single support is beyond the horizon
  };

  /** (CoM position, CoM velocity), reference :75 */
  using InitialParam = std::array<double, 2>;

  /** Objective weights with the reference's defaults and argument order (:78-117). */
  struct WeightParam
  {
    double free_zmp = 1e-2;          // ZMP of future contacts
    double fixed_zmp = 1e0;          // ZMP of existing contacts
    double double_support = 1e0;     // smoothness of the ZMP through a double-support phase
    double pos = 0.0;                // CoM position
    double vel = 0.0;                // CoM velocity
    double capture_point_abs = 1e1;  // capture point at the phase ends, absolute
    double capture_point_rel = 1e1;  // capture point relative to the next ZMP
    WeightParam() {}
    WeightParam(double w_free, double w_fixed = 1e0, double w_ds = 1e0, double w_pos = 0.0, double w_vel = 0.0, double w_cp_abs = 1e1, double w_cp_rel = 1e1)
    {
      free_zmp = w_free;
      fixed_zmp = w_fixed;
      double_support = w_ds;
      pos = w_pos;
      vel = w_vel;
      capture_point_abs = w_cp_abs;
      capture_point_rel = w_cp_rel;
    }
  };

  StepMpc1d(double com_height, const WeightParam & weight_param = WeightParam()) : com_height_(com_height), weight_param_(weight_param) {}

  /** reference :135 — the second axis of the engine call carries a copy of the problem. */
  PlannedData planOnce(const RefData & ref_data, const InitialParam & initial_param, double current_time);

  double com_height_ = 0;
  WeightParam weight_param_;
};

class StepMpc
{
public:
  using Vector2d = std::array<double, 2>;

  struct RefData
  {
    struct Element
    {
      bool is_single_support = true;
      Vector2d zmp = {0, 0};
      double end_time = 0;
    };
    std::vector<Element> element_list;
  };

  struct PlannedData
  {
    Vector2d current_zmp = {0, 0};
    std::optional<Vector2d> next_foot_zmp;
  };

  struct InitialParam
  {
    Vector2d pos = {0, 0};
    Vector2d vel = {0, 0};
  };

  StepMpc(double com_height, const StepMpc1d::WeightParam & weight_param = StepMpc1d::WeightParam())
  : mpc_1d_(std::make_shared<StepMpc1d>(com_height, weight_param))
  {
  }

  /** reference :225 */
  PlannedData planOnce(const RefData & ref_data, const InitialParam & initial_param, double current_time)
  {
    return planBatch({ref_data}, {current_time}, {initial_param}, {0})[0];
  }

  /** planOnce for initial_params[b] on ref_data[plan_id[b]] (at current_times[plan_id[b]]). */
  std::vector<PlannedData> planBatch(const std::vector<RefData> & ref_data,
                                     const std::vector<double> & current_times,
                                     const std::vector<InitialParam> & initial_params,
                                     const std::vector<int> & plan_id)
  {
    const size_t P = ref_data.size(), B = initial_params.size(), K = CCC_STEP_MPC_MAX_ELEMENTS;
    if(P == 0 || B == 0 || current_times.size() != P || plan_id.size() != B) throw std::invalid_argument("[StepMpc] planBatch: sizes");
    std::vector<int32_t> n_el(P), single(P * K, 0), pid(B);
    std::vector<double> zmp(P * K * 2, 0.0), end_time(P * K, 0.0), pos(B * 2), vel(B * 2);
    for(size_t p = 0; p < P; p++)
    {
      const auto & el = ref_data[p].element_list;
      if(el.empty() || el.size() > K) throw std::invalid_argument("[StepMpc] 1 .. " + std::to_string(K) + " elements per reference");
      n_el[p] = static_cast<int32_t>(el.size());
      for(size_t i = 0; i < el.size(); i++)
      {
        single[p * K + i] = el[i].is_single_support ? 1 : 0;
        zmp[(p * K + i) * 2] = el[i].zmp[0];
        zmp[(p * K + i) * 2 + 1] = el[i].zmp[1];
        end_time[p * K + i] = el[i].end_time;
      }
    }
    for(size_t b = 0; b < B; b++)
    {
      pid[b] = plan_id[b];
      for(size_t a = 0; a < 2; a++)
      {
        pos[2 * b + a] = initial_params[b].pos[a];
        vel[2 * b + a] = initial_params[b].vel[a];
      }
    }
    const auto & w = mpc_1d_->weight_param_;
    ccc_step_mpc_batch_t bt{};
    bt.batch = static_cast<int32_t>(B);
    bt.n_plans = static_cast<int32_t>(P);
    bt.max_elements = static_cast<int32_t>(K);
    bt.com_height = mpc_1d_->com_height_;
    bt.w_free_zmp = w.free_zmp;
    bt.w_fixed_zmp = w.fixed_zmp;
    bt.w_double_support = w.double_support;
    bt.w_pos = w.pos;
    bt.w_vel = w.vel;
    bt.w_capture_point_abs = w.capture_point_abs;
    bt.w_capture_point_rel = w.capture_point_rel;
    bt.plan_id = pid.data();
    bt.x_pos = pos.data();
    bt.x_vel = vel.data();
    bt.current_time = current_times.data();
    bt.n_elements = n_el.data();
    bt.single = single.data();
    bt.zmp = zmp.data();
    bt.end_time = end_time.data();
    std::vector<double> cur(B * 2), nxt(B * 2);
    std::vector<int32_t> has(B);
    ccc_step_mpc_result_t rs{cur.data(), nxt.data(), has.data()};
    if(ccc_step_mpc_plan(&bt, &rs, CCC_MEM_HOST, nullptr) != CCC_OK) throw std::runtime_error(std::string("[StepMpc] ") + ccc_last_error());
    std::vector<PlannedData> out(B);
    for(size_t b = 0; b < B; b++)
    {
      out[b].current_zmp = {cur[2 * b], cur[2 * b + 1]};
      if(has[b]) out[b].next_foot_zmp = Vector2d{nxt[2 * b], nxt[2 * b + 1]};
    }
    return out;
  }

  std::shared_ptr<StepMpc1d> mpc_1d_;
};

inline StepMpc1d::PlannedData StepMpc1d::planOnce(const RefData & ref_data, const InitialParam & initial_param, double current_time)
{
  StepMpc mpc(com_height_, weight_param_);
  StepMpc::RefData rd;
  for(const auto & e : ref_data.element_list) rd.element_list.push_back({e.is_single_support, {e.zmp, e.zmp}, e.end_time});
  StepMpc::InitialParam ip;
  ip.pos = {initial_param[0], initial_param[0]};
  ip.vel = {initial_param[1], initial_param[1]};
  const auto pd = mpc.planOnce(rd, ip, current_time);
  PlannedData out;
  out.current_zmp = pd.current_zmp[0];
  if(pd.next_foot_zmp) out.next_foot_zmp = (*pd.next_foot_zmp)[0];
  return out;
}
} // namespace CCC
