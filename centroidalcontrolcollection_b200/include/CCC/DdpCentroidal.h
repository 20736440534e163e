/* CCC/DdpCentroidal.h — drop-in host class for CCC::DdpCentroidal on top of the C-ABI engine.
 *
 * Mirrors reference include/CCC/DdpCentroidal.h: same nested types (MotionParam, RefData,
 * WeightParam, InitialParam), same constructor (:342) and planOnce signature (:351-354), same
 * defaults, same public force_scale_limits_ (:364).  Differences, all forced by this image:
 *  - Eigen is absent, so Vector3d = std::array<double,3> and VectorXd = std::vector<double>;
 *  - ddp_solver_ / ddp_problem_ are views (CCC/detail/DdpFacade.h) with the members the reference's callers use:
 *    ddp_solver_->config() (max_iter, lambdas, horizon_steps, ...), ->controlData().u_list,
 *    ->traceDataList().back().iter, ddp_problem_->dt(), ->inputDim(t); config(), u_list(b), lastIter(b) reach
 *    the same state per problem of a batch;
 *  - new: planBatch() solves many (schedule, initial state) pairs in one engine call.
 * planOnce() = sample callbacks -> batch of one -> ccc_ddp_centroidal_solve -> u_list[0].
 * Header-only; link with libccc_b200.so.  No CPU fallback: throws std::runtime_error without a GPU.
 */
#pragma once
#include <cmath>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/ccc_b200.h"
#include "Contact.h"
#include "detail/DdpFacade.h"

namespace CCC
{
using Vector3d = std::array<double, 3>;
using VectorXd = std::vector<double>;

class DdpCentroidal
{
public:
  /** reference include/CCC/DdpCentroidal.h:19-25 */
  struct MotionParam
  {
    std::vector<std::shared_ptr<ForceColl::Contact>> contact_list;
  };

  /** reference :28-34 */
  struct RefData
  {
    Vector3d pos = {0.0, 0.0, 0.0};
  };

  /** reference :37-82 (same defaults) */
  struct WeightParam
  {
    Vector3d running_pos;
    Vector3d running_linear_momentum;
    Vector3d running_angular_momentum;
    double running_force;
    Vector3d terminal_pos;
    Vector3d terminal_linear_momentum;
    Vector3d terminal_angular_momentum;

    WeightParam(const Vector3d & _running_pos = {1.0, 1.0, 1.0},
                const Vector3d & _running_linear_momentum = {0.0, 0.0, 0.0},
                const Vector3d & _running_angular_momentum = {1.0, 1.0, 1.0},
                double _running_force = 1e-6,
                const Vector3d & _terminal_pos = {1.0, 1.0, 1.0},
                const Vector3d & _terminal_linear_momentum = {0.0, 0.0, 0.0},
                const Vector3d & _terminal_angular_momentum = {1.0, 1.0, 1.0})
    : running_pos(_running_pos), running_linear_momentum(_running_linear_momentum),
      running_angular_momentum(_running_angular_momentum), running_force(_running_force), terminal_pos(_terminal_pos),
      terminal_linear_momentum(_terminal_linear_momentum), terminal_angular_momentum(_terminal_angular_momentum)
    {
    }
  };

  /** reference :298-331 */
  struct InitialParam
  {
    Vector3d pos = {0.0, 0.0, 0.0};
    Vector3d vel = {0.0, 0.0, 0.0};
    Vector3d angular_momentum = {0.0, 0.0, 0.0};
    /** Initial guess of the input sequence (length horizon_steps); empty = all zeros. */
    std::vector<VectorXd> u_list = {};

    /** reference src/DdpCentroidal.cpp:186-191: [pos, mass * vel, angular_momentum] */
    std::array<double, 9> toState(double mass) const
    {
      return {pos[0], pos[1], pos[2], mass * vel[0], mass * vel[1], mass * vel[2],
              angular_momentum[0], angular_momentum[1], angular_momentum[2]};
    }
  };

  /** One problem of a batch: which schedule it follows and where it starts. */
  struct BatchItem
  {
    int schedule = 0;
    InitialParam initial_param;
  };

public:
  /** reference :342 and src/DdpCentroidal.cpp:193-211 (solver configuration). */
  DdpCentroidal(double mass, double horizon_dt, int horizon_steps, const WeightParam & weight_param = WeightParam())
  : ddp_solver_(std::make_shared<detail::DdpSolverFacade<VectorXd>>()), ddp_problem_(std::make_shared<detail::DdpProblemFacade>()),
    mass_(mass), dt_(horizon_dt), horizon_steps_(horizon_steps), weight_param_(weight_param), config_(ddp_solver_->config_)
  {
    ccc_ddp_config_default(&config_);
    config_.with_input_constraint = 1;
    config_.horizon_steps = horizon_steps;
    config_.initial_lambda = 1e-6;
    config_.lambda_min = 1e-8;
    config_.lambda_thre = 1e-7;
    ddp_problem_->dt_ = horizon_dt;
  }

  ~DdpCentroidal()
  {
    if(ws_) ccc_ddp_centroidal_destroy(ws_);
  }
  DdpCentroidal(const DdpCentroidal &) = delete;
  DdpCentroidal & operator=(const DdpCentroidal &) = delete;

  /** Plan one step (reference :351-354).  Returns the planned force scales of the first stage. */
  VectorXd planOnce(const std::function<MotionParam(double)> & motion_param_func,
                    const std::function<RefData(double)> & ref_data_func,
                    const InitialParam & initial_param,
                    double current_time)
  {
    BatchItem item;
    item.schedule = 0;
    item.initial_param = initial_param;
    // ddp_problem_->setMotionParamFunc (src/DdpCentroidal.cpp:218): inputDim(t) evaluates it
    ddp_problem_->input_dim_func_ = [motion_param_func](double t) {
      int n = 0;
      for(const auto & contact : motion_param_func(t).contact_list) n += contact->ridgeNum();
      return n;
    };
    return planBatch({motion_param_func}, {ref_data_func}, {item}, current_time)[0];
  }

  /** Batched planOnce: `items[b]` follows schedule `items[b].schedule` (an index into the two
   *  callback lists).  Returns u_list[0] of every problem; full results stay readable through
   *  u_list(b) / x_list(b) / lastIter(b). */
  std::vector<VectorXd> planBatch(const std::vector<std::function<MotionParam(double)>> & motion_param_funcs,
                                  const std::vector<std::function<RefData(double)>> & ref_data_funcs,
                                  const std::vector<BatchItem> & items,
                                  double current_time)
  {
    const int N = horizon_steps_, S = static_cast<int>(motion_param_funcs.size()), B = static_cast<int>(items.size());
    if(S == 0 || ref_data_funcs.size() != motion_param_funcs.size()) throw std::runtime_error("planBatch: schedule lists");
    const int M = CCC_DDP_M_MAX;
    // sample the callbacks at t_k = current_time + k dt (reference src/DdpCentroidal.cpp:21-30, :36, :225)
    m_.assign(static_cast<size_t>(S) * N, 0);
    ridge_.assign(static_cast<size_t>(S) * N * M * 3, 0.0);
    vertex_.assign(static_cast<size_t>(S) * N * M * 3, 0.0);
    ref_.assign(static_cast<size_t>(S) * (N + 1) * 3, 0.0);
    for(int s = 0; s < S; s++)
      for(int k = 0; k <= N; k++)
      {
        const double t = current_time + k * dt_;
        const RefData rd = ref_data_funcs[s](t);
        for(int a = 0; a < 3; a++) ref_[(static_cast<size_t>(s) * (N + 1) + k) * 3 + a] = rd.pos[a];
        if(k == N) break;
        const MotionParam mp = motion_param_funcs[s](t);
        int j = 0;
        for(const auto & contact : mp.contact_list)
          for(const auto & vr : contact->vertexWithRidgeList_)
            for(const auto & ridge : vr.ridgeList)
            {
              if(j >= M) throw std::runtime_error("planBatch: more than CCC_DDP_M_MAX inputs in a stage");
              const size_t o = ((static_cast<size_t>(s) * N + k) * M + j) * 3;
              for(int a = 0; a < 3; a++)
              {
                ridge_[o + a] = ridge[a];
                vertex_[o + a] = vr.vertex[a];
              }
              j++;
            }
        m_[static_cast<size_t>(s) * N + k] = j;
      }
    sched_id_.resize(B);
    x0_.resize(static_cast<size_t>(B) * 9);
    bool warm = false;
    for(int b = 0; b < B; b++) warm = warm || !items[b].initial_param.u_list.empty();
    u_init_.assign(warm ? static_cast<size_t>(B) * N * M : 0, 0.0);
    for(int b = 0; b < B; b++)
    {
      if(items[b].schedule < 0 || items[b].schedule >= S) throw std::runtime_error("planBatch: schedule index out of range");
      sched_id_[b] = items[b].schedule;
      const auto st = items[b].initial_param.toState(mass_);
      for(int i = 0; i < 9; i++) x0_[static_cast<size_t>(b) * 9 + i] = st[i];
      const auto & ul = items[b].initial_param.u_list;
      if(!ul.empty())
      {
        if(static_cast<int>(ul.size()) != N) throw std::runtime_error("planBatch: u_list length != horizon_steps");
        for(int k = 0; k < N; k++)
          for(size_t j = 0; j < ul[k].size() && j < static_cast<size_t>(M); j++)
            u_init_[(static_cast<size_t>(b) * N + k) * M + j] = ul[k][j];
      }
    }
    ensureWorkspace(B, S);
    x_.assign(static_cast<size_t>(B) * (N + 1) * 9, 0.0);
    u_.assign(static_cast<size_t>(B) * N * M, 0.0);
    cost_.assign(B, 0.0);
    iters_.assign(B, 0);
    status_.assign(B, 0);

    ccc_ddp_centroidal_batch_t bt{};
    bt.horizon_steps = N;
    bt.batch = B;
    bt.n_sched = S;
    bt.m_max = M;
    bt.dt = dt_;
    bt.mass = mass_;
    bt.sched_id = sched_id_.data();
    bt.m = m_.data();
    bt.ridge = ridge_.data();
    bt.vertex = vertex_.data();
    bt.ref_pos = ref_.data();
    for(int a = 0; a < 3; a++)
    {
      bt.w_run[a] = weight_param_.running_pos[a];
      bt.w_run[3 + a] = weight_param_.running_linear_momentum[a];
      bt.w_run[6 + a] = weight_param_.running_angular_momentum[a];
      bt.w_term[a] = weight_param_.terminal_pos[a];
      bt.w_term[3 + a] = weight_param_.terminal_linear_momentum[a];
      bt.w_term[6 + a] = weight_param_.terminal_angular_momentum[a];
    }
    bt.w_run[9] = weight_param_.running_force;
    bt.u_lo = force_scale_limits_[0];
    bt.u_hi = force_scale_limits_[1];
    bt.x0 = x0_.data();
    bt.u_init = warm ? u_init_.data() : nullptr;
    ccc_ddp_result_t rs{};
    rs.x = x_.data();
    rs.u = u_.data();
    rs.cost = cost_.data();
    rs.iters = iters_.data();
    rs.status = status_.data();
    const int rc = ccc_ddp_centroidal_solve(ws_, &bt, &config_, &rs, CCC_MEM_HOST, nullptr);
    if(rc != CCC_OK) throw std::runtime_error(std::string("ccc_ddp_centroidal_solve: ") + ccc_last_error());

    batch_ = B;
    std::vector<VectorXd> first(B);
    for(int b = 0; b < B; b++) first[b] = u_list(b)[0];
    // what the reference's callers read from the solver object after planOnce (problem 0 of a batch)
    ddp_solver_->control_data_.u_list = u_list(0);
    ddp_solver_->trace_data_list_.assign(1, {iters_[0]});
    return first;
  }

  /** Result of runClosedLoopBatch: plant state (position, velocity, angular momentum) of every problem at the
   *  start of every control cycle and after the last one, and the DDP iterations of every cycle. */
  struct ClosedLoopResult
  {
    int batch = 0, ticks = 0;
    std::vector<double> plant; // [batch][ticks + 1][9]
    std::vector<int32_t> iters; // [batch][ticks]
    const double * state(int b, int tick) const { return &plant[(static_cast<size_t>(b) * (ticks + 1) + tick) * 9]; }
    int iter(int b, int tick) const { return iters[static_cast<size_t>(b) * ticks + tick]; }
  };

  /** The receding-horizon loop of the reference's own test (tests/src/TestDdpCentroidal.cpp:94-150) for a batch of
   *  plants, resident on the device (ccc_ddp_centroidal_closed_loop): every cycle plans from the plant's state with
   *  the previous plan as warm start (config().max_iter iterations in the first cycle, max_iter_later afterwards),
   *  applies the first stage's wrench to the plant (point mass + angular momentum, exact zero-order hold over
   *  sim_dt) and moves on.  The callbacks are sampled at start_time + entry * sim_dt; horizon_dt must be an integer
   *  multiple of sim_dt.  items[b].initial_param gives the initial plant state (u_list is ignored). */
  ClosedLoopResult runClosedLoopBatch(const std::vector<std::function<MotionParam(double)>> & motion_param_funcs,
                                      const std::vector<std::function<RefData(double)>> & ref_data_funcs,
                                      const std::vector<BatchItem> & items,
                                      double start_time,
                                      double sim_dt,
                                      int ticks,
                                      int max_iter_later = 1,
                                      int disturb_tick = -1,
                                      const Vector3d & disturb_vel = {0.0, 0.0, 0.0})
  {
    const int N = horizon_steps_, S = static_cast<int>(motion_param_funcs.size()), B = static_cast<int>(items.size());
    if(S == 0 || ref_data_funcs.size() != motion_param_funcs.size() || B == 0 || ticks <= 0)
      throw std::runtime_error("runClosedLoopBatch: schedule lists / batch / ticks");
    const int stride = static_cast<int>(dt_ / sim_dt + 0.5);
    if(stride < 1 || std::fabs(stride * sim_dt - dt_) > 1e-12 * dt_)
      throw std::runtime_error("runClosedLoopBatch: horizon_dt must be an integer multiple of sim_dt");
    const int M = CCC_DDP_M_MAX, G = ticks - 1 + N * stride + 1;
    std::vector<int32_t> m(static_cast<size_t>(S) * G, 0), sched_id(B);
    std::vector<double> ridge(static_cast<size_t>(S) * G * M * 3, 0.0), vertex(ridge.size(), 0.0), ref(static_cast<size_t>(S) * G * 3, 0.0);
    for(int s = 0; s < S; s++)
      for(int e = 0; e < G; e++)
      {
        const double t = start_time + e * sim_dt;
        const RefData rd = ref_data_funcs[s](t);
        for(int a = 0; a < 3; a++) ref[(static_cast<size_t>(s) * G + e) * 3 + a] = rd.pos[a];
        const MotionParam mp = motion_param_funcs[s](t);
        int j = 0;
        for(const auto & contact : mp.contact_list)
          for(const auto & vr : contact->vertexWithRidgeList_)
            for(const auto & r : vr.ridgeList)
            {
              if(j >= M) throw std::runtime_error("runClosedLoopBatch: more than CCC_DDP_M_MAX inputs in a stage");
              const size_t o = ((static_cast<size_t>(s) * G + e) * M + j) * 3;
              for(int a = 0; a < 3; a++)
              {
                ridge[o + a] = r[a];
                vertex[o + a] = vr.vertex[a];
              }
              j++;
            }
        m[static_cast<size_t>(s) * G + e] = j;
      }
    std::vector<double> plant0(static_cast<size_t>(B) * 9);
    for(int b = 0; b < B; b++)
    {
      if(items[b].schedule < 0 || items[b].schedule >= S) throw std::runtime_error("runClosedLoopBatch: schedule index out of range");
      sched_id[b] = items[b].schedule;
      const InitialParam & ip = items[b].initial_param;
      for(int a = 0; a < 3; a++)
      {
        plant0[static_cast<size_t>(b) * 9 + a] = ip.pos[a];
        plant0[static_cast<size_t>(b) * 9 + 3 + a] = ip.vel[a];
        plant0[static_cast<size_t>(b) * 9 + 6 + a] = ip.angular_momentum[a];
      }
    }
    ensureWorkspace(B, S);
    ClosedLoopResult out;
    out.batch = B;
    out.ticks = ticks;
    out.plant.assign(static_cast<size_t>(B) * (ticks + 1) * 9, 0.0);
    out.iters.assign(static_cast<size_t>(B) * ticks, 0);
    ccc_ddp_centroidal_loop_t lp{};
    lp.horizon_steps = N;
    lp.batch = B;
    lp.n_sched = S;
    lp.m_max = M;
    lp.dt = dt_;
    lp.mass = mass_;
    lp.sim_dt = sim_dt;
    lp.ticks = ticks;
    lp.stride = stride;
    lp.grid_len = G;
    lp.max_iter_later = max_iter_later;
    lp.sched_id = sched_id.data();
    lp.m = m.data();
    lp.ridge = ridge.data();
    lp.vertex = vertex.data();
    lp.ref_pos = ref.data();
    for(int a = 0; a < 3; a++)
    {
      lp.w_run[a] = weight_param_.running_pos[a];
      lp.w_run[3 + a] = weight_param_.running_linear_momentum[a];
      lp.w_run[6 + a] = weight_param_.running_angular_momentum[a];
      lp.w_term[a] = weight_param_.terminal_pos[a];
      lp.w_term[3 + a] = weight_param_.terminal_linear_momentum[a];
      lp.w_term[6 + a] = weight_param_.terminal_angular_momentum[a];
      lp.disturb_vel[a] = disturb_vel[a];
    }
    lp.w_run[9] = weight_param_.running_force;
    lp.u_lo = force_scale_limits_[0];
    lp.u_hi = force_scale_limits_[1];
    lp.plant0 = plant0.data();
    lp.disturb_tick = disturb_tick;
    ccc_ddp_centroidal_loop_result_t rs{};
    rs.plant = out.plant.data();
    rs.iters = out.iters.data();
    const int rc = ccc_ddp_centroidal_closed_loop(ws_, &lp, &config_, &rs, CCC_MEM_HOST, nullptr);
    if(rc != CCC_OK) throw std::runtime_error(std::string("ccc_ddp_centroidal_closed_loop: ") + ccc_last_error());
    return out;
  }

  /** ddp_solver_->config() of the reference (max_iter, lambdas, ...). */
  detail::DdpConfiguration & config() { return config_; }

  /** ddp_solver_->controlData().u_list of problem b of the last call (stage vectors sized inputDim). */
  std::vector<VectorXd> u_list(int b = 0) const
  {
    const int N = horizon_steps_, M = CCC_DDP_M_MAX;
    std::vector<VectorXd> out(N);
    for(int k = 0; k < N; k++)
    {
      const int m = m_[static_cast<size_t>(sched_id_[b]) * N + k];
      out[k].assign(u_.begin() + (static_cast<size_t>(b) * N + k) * M, u_.begin() + (static_cast<size_t>(b) * N + k) * M + m);
    }
    return out;
  }
  /** ddp_solver_->controlData().x_list */
  std::vector<std::array<double, 9>> x_list(int b = 0) const
  {
    const int N = horizon_steps_;
    std::vector<std::array<double, 9>> out(N + 1);
    for(int k = 0; k <= N; k++)
      for(int i = 0; i < 9; i++) out[k][i] = x_[(static_cast<size_t>(b) * (N + 1) + k) * 9 + i];
    return out;
  }
  /** ddp_solver_->traceDataList().back().iter */
  int lastIter(int b = 0) const { return iters_[b]; }
  int lastStatus(int b = 0) const { return status_[b]; }
  bool hasSolution() const { return batch_ > 0; }
  /** ddp_problem_->inputDim(t) for stage k of the last sampled schedule of problem b. */
  int inputDim(int k, int b = 0) const { return m_[static_cast<size_t>(sched_id_[b]) * horizon_steps_ + k]; }
  double dt() const { return dt_; }
  int horizonSteps() const { return horizon_steps_; }

public:
  //! DDP solver as the reference's callers see it (reference include/CCC/DdpCentroidal.h:361)
  std::shared_ptr<detail::DdpSolverFacade<VectorXd>> ddp_solver_;
  //! DDP problem as the reference's callers see it (:358)
  std::shared_ptr<detail::DdpProblemFacade> ddp_problem_;
  //! Robot mass [kg]
  double mass_ = 0;
  //! Force scale limits (lower, upper), reference include/CCC/DdpCentroidal.h:364
  std::array<double, 2> force_scale_limits_ = {0.0, 1e6};

private:
  void ensureWorkspace(int B, int S)
  {
    if(ws_ && B <= ws_batch_ && S <= ws_sched_) return;
    if(ws_) ccc_ddp_centroidal_destroy(ws_);
    ws_ = ccc_ddp_centroidal_create(horizon_steps_, B, S);
    if(!ws_) throw std::runtime_error(std::string("ccc_ddp_centroidal_create: ") + ccc_last_error());
    ws_batch_ = B;
    ws_sched_ = S;
  }

  double dt_;
  int horizon_steps_;
  WeightParam weight_param_;
  detail::DdpConfiguration & config_; // lives in ddp_solver_
  ccc_ddp_centroidal_ws_t * ws_ = nullptr;
  int ws_batch_ = 0, ws_sched_ = 0, batch_ = 0;
  std::vector<int32_t> m_, sched_id_, iters_, status_;
  std::vector<double> ridge_, vertex_, ref_, x0_, u_init_, x_, u_, cost_;
};
} // namespace CCC
