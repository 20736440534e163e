/* CCC/DdpSingleRigidBody.h — drop-in host class for CCC::DdpSingleRigidBody on top of the C-ABI engine.
 *
 * Mirrors reference include/CCC/DdpSingleRigidBody.h: nested types MotionParam (:23-34), RefData (:37-46),
 * WeightParam (:49-104, same defaults), InitialParam (:335-365); constructor (:375-378) with the solver
 * configuration of src/DdpSingleRigidBody.cpp:261-281; planOnce (:391-394, src :283-307); public
 * force_scale_limits_ (:405).  Eigen is absent from this image: Vector3d = std::array<double,3>,
 * Matrix3d = row-major std::array<double,9>, VectorXd = std::vector<double>.
 * ddp_solver_ / ddp_problem_ are views (CCC/detail/DdpFacade.h) with the members the reference's callers use
 * (config().max_iter / horizon_steps, controlData().u_list, traceDataList().back().iter, dt(), inputDim(t));
 * config(), u_list(b), lastIter(b) reach the same state per problem of a batch.
 * New: planBatch() — many (schedule, initial state) pairs in one engine call.
 * Header-only; link with libccc_b200.so.  No CPU fallback: throws std::runtime_error without a GPU.
 */
#pragma once
#include <array>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "detail/DdpFacade.h"
#include "detail/RidgeTables.h"

namespace CCC
{
class DdpSingleRigidBody
{
public:
  using Vector3d = std::array<double, 3>;
  using Matrix3d = std::array<double, 9>;
  using VectorXd = std::vector<double>;
  using StateDimVector = std::array<double, 12>;

  struct MotionParam
  {
    std::vector<std::shared_ptr<ForceColl::Contact>> contact_list;
    Matrix3d inertia_mat = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  };

  struct RefData
  {
    Vector3d pos = {0, 0, 0};
    Vector3d ori = {0, 0, 0};
  };

  struct WeightParam
  {
    Vector3d running_pos, running_ori, running_linear_vel, running_angular_vel;
    double running_force;
    Vector3d terminal_pos, terminal_ori, terminal_linear_vel, terminal_angular_vel;

    WeightParam(const Vector3d & _running_pos = {1.0, 1.0, 1.0},
                const Vector3d & _running_ori = {1.0, 1.0, 1.0},
                const Vector3d & _running_linear_vel = {0.01, 0.01, 0.01},
                const Vector3d & _running_angular_vel = {0.01, 0.01, 0.01},
                double _running_force = 1e-6,
                const Vector3d & _terminal_pos = {1.0, 1.0, 1.0},
                const Vector3d & _terminal_ori = {1.0, 1.0, 1.0},
                const Vector3d & _terminal_linear_vel = {0.01, 0.01, 0.01},
                const Vector3d & _terminal_angular_vel = {0.01, 0.01, 0.01})
    : running_pos(_running_pos), running_ori(_running_ori), running_linear_vel(_running_linear_vel),
      running_angular_vel(_running_angular_vel), running_force(_running_force), terminal_pos(_terminal_pos),
      terminal_ori(_terminal_ori), terminal_linear_vel(_terminal_linear_vel), terminal_angular_vel(_terminal_angular_vel)
    {
    }
  };

  struct InitialParam
  {
    Vector3d pos = {0, 0, 0};
    Vector3d ori = {0, 0, 0}; // ZYX Euler angles
    Vector3d linear_vel = {0, 0, 0};
    Vector3d angular_vel = {0, 0, 0};
    /** Initial guess of the input sequence (length horizon_steps); empty = all zeros. */
    std::vector<VectorXd> u_list = {};

    InitialParam() {}
    /** reference src/DdpSingleRigidBody.cpp:247-253 */
    explicit InitialParam(const StateDimVector & state)
    {
      for(int a = 0; a < 3; a++)
      {
        pos[a] = state[a];
        ori[a] = state[3 + a];
        linear_vel[a] = state[6 + a];
        angular_vel[a] = state[9 + a];
      }
    }
    /** reference src/DdpSingleRigidBody.cpp:255-260 */
    StateDimVector toState() const
    {
      return {pos[0], pos[1], pos[2], ori[0], ori[1], ori[2], linear_vel[0], linear_vel[1], linear_vel[2],
              angular_vel[0], angular_vel[1], angular_vel[2]};
    }
  };

  struct BatchItem
  {
    int schedule = 0;
    InitialParam initial_param;
  };

public:
  DdpSingleRigidBody(double mass, double horizon_dt, int horizon_steps, const WeightParam & weight_param = WeightParam())
  : ddp_solver_(std::make_shared<detail::DdpSolverFacade<VectorXd>>()), ddp_problem_(std::make_shared<detail::DdpProblemFacade>()),
    mass_(mass), dt_(horizon_dt), horizon_steps_(horizon_steps), weight_param_(weight_param), config_(ddp_solver_->config_)
  {
    ccc_ddp_config_default(&config_);
    config_.with_input_constraint = 1;
    config_.horizon_steps = horizon_steps;
    ddp_problem_->dt_ = horizon_dt;
    config_.initial_lambda = 1e-6;
    config_.lambda_min = 1e-8;
    config_.lambda_thre = 1e-7;
  }
  ~DdpSingleRigidBody()
  {
    if(ws_) ccc_ddp_srb_destroy(ws_);
  }
  DdpSingleRigidBody(const DdpSingleRigidBody &) = delete;
  DdpSingleRigidBody & operator=(const DdpSingleRigidBody &) = delete;

  /** Plan one step (reference :391-394): planned force scales of the first stage. */
  VectorXd planOnce(const std::function<MotionParam(double)> & motion_param_func,
                    const std::function<RefData(double)> & ref_data_func,
                    const InitialParam & initial_param,
                    double current_time)
  {
    BatchItem item;
    item.initial_param = initial_param;
    // ddp_problem_->setMotionParamFunc (src/DdpSingleRigidBody.cpp:288): inputDim(t) evaluates it
    ddp_problem_->input_dim_func_ = [motion_param_func](double t) {
      int n = 0;
      for(const auto & contact : motion_param_func(t).contact_list) n += contact->ridgeNum();
      return n;
    };
    return planBatch({motion_param_func}, {ref_data_func}, {item}, current_time)[0];
  }

  /** Batched planOnce: items[b] follows schedule items[b].schedule (index into the callback lists). */
  std::vector<VectorXd> planBatch(const std::vector<std::function<MotionParam(double)>> & motion_param_funcs,
                                  const std::vector<std::function<RefData(double)>> & ref_data_funcs,
                                  const std::vector<BatchItem> & items,
                                  double current_time)
  {
    const int N = horizon_steps_, S = static_cast<int>(motion_param_funcs.size()), B = static_cast<int>(items.size());
    if(S == 0 || ref_data_funcs.size() != motion_param_funcs.size()) throw std::runtime_error("planBatch: schedule lists");
    const int M = CCC_DDP_M_MAX;
    tab_.reset(S, N);
    inertia_.assign(static_cast<size_t>(S) * N * 9, 0.0);
    ref_.assign(static_cast<size_t>(S) * (N + 1) * 6, 0.0);
    for(int s = 0; s < S; s++)
      for(int k = 0; k <= N; k++)
      {
        const double t = current_time + k * dt_;
        const RefData rd = ref_data_funcs[s](t);
        for(int a = 0; a < 3; a++)
        {
          ref_[(static_cast<size_t>(s) * (N + 1) + k) * 6 + a] = rd.pos[a];
          ref_[(static_cast<size_t>(s) * (N + 1) + k) * 6 + 3 + a] = rd.ori[a];
        }
        if(k == N) break;
        const MotionParam mp = motion_param_funcs[s](t);
        tab_.setStage(s, k, mp.contact_list);
        for(int i = 0; i < 9; i++) inertia_[(static_cast<size_t>(s) * N + k) * 9 + i] = mp.inertia_mat[i];
      }
    sched_id_.resize(B);
    x0_.resize(static_cast<size_t>(B) * 12);
    bool warm = false;
    for(int b = 0; b < B; b++) warm = warm || !items[b].initial_param.u_list.empty();
    u_init_.assign(warm ? static_cast<size_t>(B) * N * M : 0, 0.0);
    for(int b = 0; b < B; b++)
    {
      if(items[b].schedule < 0 || items[b].schedule >= S) throw std::runtime_error("planBatch: schedule index out of range");
      sched_id_[b] = items[b].schedule;
      const auto st = items[b].initial_param.toState();
      for(int i = 0; i < 12; i++) x0_[static_cast<size_t>(b) * 12 + i] = st[i];
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
    x_.assign(static_cast<size_t>(B) * (N + 1) * 12, 0.0);
    u_.assign(static_cast<size_t>(B) * N * M, 0.0);
    cost_.assign(B, 0.0);
    iters_.assign(B, 0);
    status_.assign(B, 0);

    ccc_ddp_srb_batch_t bt{};
    bt.horizon_steps = N;
    bt.batch = B;
    bt.n_sched = S;
    bt.m_max = M;
    bt.dt = dt_;
    bt.mass = mass_;
    bt.sched_id = sched_id_.data();
    bt.m = tab_.m.data();
    bt.ridge = tab_.ridge.data();
    bt.vertex = tab_.vertex.data();
    bt.inertia = inertia_.data();
    bt.ref = ref_.data();
    for(int a = 0; a < 3; a++)
    {
      bt.w_run[a] = weight_param_.running_pos[a];
      bt.w_run[3 + a] = weight_param_.running_ori[a];
      bt.w_run[6 + a] = weight_param_.running_linear_vel[a];
      bt.w_run[9 + a] = weight_param_.running_angular_vel[a];
      bt.w_term[a] = weight_param_.terminal_pos[a];
      bt.w_term[3 + a] = weight_param_.terminal_ori[a];
      bt.w_term[6 + a] = weight_param_.terminal_linear_vel[a];
      bt.w_term[9 + a] = weight_param_.terminal_angular_vel[a];
    }
    bt.w_run[12] = weight_param_.running_force;
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
    const int rc = ccc_ddp_srb_solve(ws_, &bt, &config_, &rs, CCC_MEM_HOST, nullptr);
    if(rc != CCC_OK) throw std::runtime_error(std::string("ccc_ddp_srb_solve: ") + ccc_last_error());
    batch_ = B;
    std::vector<VectorXd> first(B);
    for(int b = 0; b < B; b++) first[b] = u_list(b)[0];
    ddp_solver_->control_data_.u_list = u_list(0);
    ddp_solver_->trace_data_list_.assign(1, {iters_[0]});
    return first;
  }

  detail::DdpConfiguration & config() { return config_; }

  std::vector<VectorXd> u_list(int b = 0) const
  {
    const int N = horizon_steps_, M = CCC_DDP_M_MAX;
    std::vector<VectorXd> out(N);
    for(int k = 0; k < N; k++)
    {
      const size_t o = (static_cast<size_t>(b) * N + k) * M;
      out[k].assign(u_.begin() + o, u_.begin() + o + tab_.inputDim(sched_id_[b], k));
    }
    return out;
  }
  std::vector<StateDimVector> x_list(int b = 0) const
  {
    const int N = horizon_steps_;
    std::vector<StateDimVector> out(N + 1);
    for(int k = 0; k <= N; k++)
      for(int i = 0; i < 12; i++) out[k][i] = x_[(static_cast<size_t>(b) * (N + 1) + k) * 12 + i];
    return out;
  }
  int lastIter(int b = 0) const { return iters_[b]; }
  int lastStatus(int b = 0) const { return status_[b]; }
  bool hasSolution() const { return batch_ > 0; }
  int inputDim(int k, int b = 0) const { return tab_.inputDim(sched_id_[b], k); }
  double dt() const { return dt_; }
  int horizonSteps() const { return horizon_steps_; }

public:
  //! DDP solver / problem as the reference's callers see them (reference include/CCC/DdpSingleRigidBody.h:398-402)
  std::shared_ptr<detail::DdpSolverFacade<VectorXd>> ddp_solver_;
  std::shared_ptr<detail::DdpProblemFacade> ddp_problem_;
  double mass_ = 0;
  //! Force scale limits (lower, upper), reference include/CCC/DdpSingleRigidBody.h:405
  std::array<double, 2> force_scale_limits_ = {0.0, 1e6};

private:
  void ensureWorkspace(int B, int S)
  {
    if(ws_ && B <= ws_batch_ && S <= ws_sched_) return;
    if(ws_) ccc_ddp_srb_destroy(ws_);
    ws_ = ccc_ddp_srb_create(horizon_steps_, B, S);
    if(!ws_) throw std::runtime_error(std::string("ccc_ddp_srb_create: ") + ccc_last_error());
    ws_batch_ = B;
    ws_sched_ = S;
  }

  double dt_;
  int horizon_steps_;
  WeightParam weight_param_;
  detail::DdpConfiguration & config_; // lives in ddp_solver_
  ccc_ddp_srb_ws_t * ws_ = nullptr;
  int ws_batch_ = 0, ws_sched_ = 0, batch_ = 0;
  detail::RidgeTables tab_;
  std::vector<int32_t> sched_id_, iters_, status_;
  std::vector<double> inertia_, ref_, x0_, u_init_, x_, u_, cost_;
};
} // namespace CCC
