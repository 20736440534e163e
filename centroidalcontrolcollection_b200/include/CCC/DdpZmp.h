/* CCC/DdpZmp.h — drop-in host class for CCC::DdpZmp on top of the C-ABI engine.
 *
 * Mirrors reference include/CCC/DdpZmp.h: RefData (:20-28), PlannedData (:31-39), WeightParam (:42-85, same
 * defaults), InitialParam (:233-254), constructor (:265-270), planOnce (:278-280, src/DdpZmp.cpp:152-174).
 * State (pos_x, vel_x, pos_y, vel_y, pos_z, vel_z), input (zmp_x, zmp_y, force_z), no input limits
 * (nmpc_ddp defaults: with_input_constraint = false).  Eigen is absent: Vector2d/3d are std::array.
 * New: planBatch().  Header-only; link with libccc_b200.so; no CPU fallback.
 */
#pragma once
#include <array>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include <memory>

#include "../../../include/ccc_b200.h"
#include "detail/DdpFacade.h"

namespace CCC
{
class DdpZmp
{
public:
  using Vector2d = std::array<double, 2>;
  using Vector3d = std::array<double, 3>;
  using InputDimVector = std::array<double, 3>;
  using StateDimVector = std::array<double, 6>;

  struct RefData
  {
    Vector3d zmp = {0, 0, 0};
    double com_z = 0;
  };

  struct PlannedData
  {
    Vector2d zmp = {0, 0};
    double force_z = 0;
  };

  struct WeightParam
  {
    double running_com_pos_z, running_zmp, running_force_z, terminal_com_pos_xy, terminal_com_pos_z, terminal_com_vel;
    WeightParam(double _running_com_pos_z = 1e2,
                double _running_zmp = 1e-1,
                double _running_force_z = 1e-4,
                double _terminal_com_pos_xy = 1.0,
                double _terminal_com_pos_z = 1e2,
                double _terminal_com_vel = 1.0)
    : running_com_pos_z(_running_com_pos_z), running_zmp(_running_zmp), running_force_z(_running_force_z),
      terminal_com_pos_xy(_terminal_com_pos_xy), terminal_com_pos_z(_terminal_com_pos_z), terminal_com_vel(_terminal_com_vel)
    {
    }
  };

  struct InitialParam
  {
    Vector3d pos = {0, 0, 0};
    Vector3d vel = {0, 0, 0};
    /** Initial guess of the input sequence (length horizon_steps); empty = all zeros. */
    std::vector<InputDimVector> u_list = {};
    /** reference src/DdpZmp.cpp:145-150 */
    StateDimVector toState() const { return {pos[0], vel[0], pos[1], vel[1], pos[2], vel[2]}; }
  };

  struct BatchItem
  {
    int schedule = 0;
    InitialParam initial_param;
  };

public:
  DdpZmp(double mass, double horizon_dt, int horizon_steps, const WeightParam & weight_param = WeightParam())
  : ddp_solver_(std::make_shared<detail::DdpSolverFacade<InputDimVector>>()), ddp_problem_(std::make_shared<detail::DdpProblemFacade>()),
    mass_(mass), dt_(horizon_dt), horizon_steps_(horizon_steps), weight_param_(weight_param), config_(ddp_solver_->config_)
  {
    ccc_ddp_config_default(&config_); // nmpc_ddp defaults; the reference only sets horizon_steps (:269)
    config_.horizon_steps = horizon_steps;
    ddp_problem_->dt_ = horizon_dt;
    ddp_problem_->fixed_input_dim_ = 3;
  }
  ~DdpZmp()
  {
    if(ws_) ccc_ddp_zmp_destroy(ws_);
  }
  DdpZmp(const DdpZmp &) = delete;
  DdpZmp & operator=(const DdpZmp &) = delete;

  PlannedData planOnce(const std::function<RefData(double)> & ref_data_func, const InitialParam & initial_param, double current_time)
  {
    BatchItem item;
    item.initial_param = initial_param;
    return planBatch({ref_data_func}, {item}, current_time)[0];
  }

  std::vector<PlannedData> planBatch(const std::vector<std::function<RefData(double)>> & ref_data_funcs,
                                     const std::vector<BatchItem> & items,
                                     double current_time)
  {
    const int N = horizon_steps_, S = static_cast<int>(ref_data_funcs.size()), B = static_cast<int>(items.size());
    if(S == 0) throw std::runtime_error("planBatch: no reference schedule");
    ref_zmp_.assign(static_cast<size_t>(S) * (N + 1) * 3, 0.0);
    com_z_.assign(static_cast<size_t>(S) * (N + 1), 0.0);
    for(int s = 0; s < S; s++)
      for(int k = 0; k <= N; k++)
      {
        const RefData rd = ref_data_funcs[s](current_time + k * dt_);
        for(int a = 0; a < 3; a++) ref_zmp_[(static_cast<size_t>(s) * (N + 1) + k) * 3 + a] = rd.zmp[a];
        com_z_[static_cast<size_t>(s) * (N + 1) + k] = rd.com_z;
      }
    sched_id_.resize(B);
    x0_.resize(static_cast<size_t>(B) * 6);
    bool warm = false;
    for(int b = 0; b < B; b++) warm = warm || !items[b].initial_param.u_list.empty();
    u_init_.assign(warm ? static_cast<size_t>(B) * N * 3 : 0, 0.0);
    for(int b = 0; b < B; b++)
    {
      if(items[b].schedule < 0 || items[b].schedule >= S) throw std::runtime_error("planBatch: schedule index out of range");
      sched_id_[b] = items[b].schedule;
      const auto st = items[b].initial_param.toState();
      for(int i = 0; i < 6; i++) x0_[static_cast<size_t>(b) * 6 + i] = st[i];
      const auto & ul = items[b].initial_param.u_list;
      if(!ul.empty())
      {
        if(static_cast<int>(ul.size()) != N) throw std::runtime_error("planBatch: u_list length != horizon_steps");
        for(int k = 0; k < N; k++)
          for(int j = 0; j < 3; j++) u_init_[(static_cast<size_t>(b) * N + k) * 3 + j] = ul[k][j];
      }
    }
    if(!ws_ || B > ws_batch_ || S > ws_sched_)
    {
      if(ws_) ccc_ddp_zmp_destroy(ws_);
      ws_ = ccc_ddp_zmp_create(N, B, S);
      if(!ws_) throw std::runtime_error(std::string("ccc_ddp_zmp_create: ") + ccc_last_error());
      ws_batch_ = B;
      ws_sched_ = S;
    }
    x_.assign(static_cast<size_t>(B) * (N + 1) * 6, 0.0);
    u_.assign(static_cast<size_t>(B) * N * 3, 0.0);
    cost_.assign(B, 0.0);
    iters_.assign(B, 0);
    status_.assign(B, 0);

    ccc_ddp_zmp_batch_t bt{};
    bt.horizon_steps = N;
    bt.batch = B;
    bt.n_sched = S;
    bt.dt = dt_;
    bt.mass = mass_;
    bt.sched_id = sched_id_.data();
    bt.ref_zmp = ref_zmp_.data();
    bt.com_z = com_z_.data();
    bt.w[0] = weight_param_.running_com_pos_z;
    bt.w[1] = weight_param_.running_zmp;
    bt.w[2] = weight_param_.running_force_z;
    bt.w[3] = weight_param_.terminal_com_pos_xy;
    bt.w[4] = weight_param_.terminal_com_pos_z;
    bt.w[5] = weight_param_.terminal_com_vel;
    bt.x0 = x0_.data();
    bt.u_init = warm ? u_init_.data() : nullptr;
    ccc_ddp_result_t rs{};
    rs.x = x_.data();
    rs.u = u_.data();
    rs.cost = cost_.data();
    rs.iters = iters_.data();
    rs.status = status_.data();
    const int rc = ccc_ddp_zmp_solve(ws_, &bt, &config_, &rs, CCC_MEM_HOST, nullptr);
    if(rc != CCC_OK) throw std::runtime_error(std::string("ccc_ddp_zmp_solve: ") + ccc_last_error());
    batch_ = B;
    std::vector<PlannedData> out(B);
    for(int b = 0; b < B; b++)
    {
      const double * u0 = &u_[static_cast<size_t>(b) * N * 3];
      out[b].zmp = {u0[0], u0[1]};
      out[b].force_z = u0[2];
    }
    ddp_solver_->control_data_.u_list = u_list(0);
    ddp_solver_->trace_data_list_.assign(1, {iters_[0]});
    return out;
  }

  detail::DdpConfiguration & config() { return config_; }
  /** ddp_solver_->controlData().u_list of problem b of the last call */
  std::vector<InputDimVector> u_list(int b = 0) const
  {
    const int N = horizon_steps_;
    std::vector<InputDimVector> out(N);
    for(int k = 0; k < N; k++)
      for(int j = 0; j < 3; j++) out[k][j] = u_[(static_cast<size_t>(b) * N + k) * 3 + j];
    return out;
  }
  std::vector<StateDimVector> x_list(int b = 0) const
  {
    const int N = horizon_steps_;
    std::vector<StateDimVector> out(N + 1);
    for(int k = 0; k <= N; k++)
      for(int i = 0; i < 6; i++) out[k][i] = x_[(static_cast<size_t>(b) * (N + 1) + k) * 6 + i];
    return out;
  }
  int lastIter(int b = 0) const { return iters_[b]; }
  int lastStatus(int b = 0) const { return status_[b]; }
  bool hasSolution() const { return batch_ > 0; }
  double dt() const { return dt_; }
  int horizonSteps() const { return horizon_steps_; }

public:
  //! DDP solver / problem as the reference's callers see them (reference include/CCC/DdpZmp.h:295-299)
  std::shared_ptr<detail::DdpSolverFacade<InputDimVector>> ddp_solver_;
  std::shared_ptr<detail::DdpProblemFacade> ddp_problem_;
  double mass_ = 0;

private:
  double dt_;
  int horizon_steps_;
  WeightParam weight_param_;
  detail::DdpConfiguration & config_; // lives in ddp_solver_
  ccc_ddp_zmp_ws_t * ws_ = nullptr;
  int ws_batch_ = 0, ws_sched_ = 0, batch_ = 0;
  std::vector<int32_t> sched_id_, iters_, status_;
  std::vector<double> ref_zmp_, com_z_, x0_, u_init_, x_, u_, cost_;
};
} // namespace CCC
