/* ccc_b200.h — C ABI of the B200-native batched centroidal-MPC solve engine.
 *
 * This is the drop-in boundary.  The reference
 * (isri-aist/CentroidalControlCollection) has no FFI: its boundary is the C++
 * `CCC::<Method>::planOnce()` member functions of libCCC.so.  Every entry point below
 * replaces the arithmetic that one of those member functions triggers, with a batch
 * axis of independent problem instances added.  The C++ classes of the same names in
 * `centroidalcontrolcollection_b200/include/CCC/` sample the user's callbacks on the host
 * into the flat stage tables declared here and call these functions.
 *
 * Conventions
 *  - all arrays are row-major, `double` unless stated, batch index outermost;
 *  - the caller owns every buffer; the library allocates only inside *_create();
 *  - `CCC_MEM_HOST`: pointers are host memory, the call copies H2D/D2H itself and is
 *    synchronous; `CCC_MEM_DEVICE`: pointers are device memory on the current CUDA
 *    device, the call only enqueues work on `stream` (a cudaStream_t passed as void*);
 *  - return value: 0 = ok, <0 = CCC_ERR_*; never throws; per-problem solver outcome
 *    is reported in `status[]`, mirroring nmpc_ddp's procOnce return codes;
 *  - table contents (stage input dimensions m <= m_max, sched_id / plan_id / group_id ranges) are validated for
 *    CCC_MEM_HOST calls, which return CCC_ERR_INVALID; CCC_MEM_DEVICE calls cannot read them without a
 *    synchronisation and do NOT check them: out-of-range values there are undefined behaviour, as for any kernel;
 *  - there is NO CPU fallback: with no CUDA device every solve returns CCC_ERR_CUDA.
 */
#ifndef CCC_B200_H
#define CCC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCC_B200_ABI_VERSION 1

enum {
  CCC_OK = 0,
  CCC_ERR_INVALID = -1, /* bad argument (size mismatch, null pointer, m > m_max ...) */
  CCC_ERR_CUDA = -2,    /* CUDA runtime error / no device; see ccc_last_error() */
  CCC_ERR_ALLOC = -3    /* workspace too small for this batch */
};

enum { CCC_MEM_HOST = 0, CCC_MEM_DEVICE = 1 };

/* Largest per-stage input dimension the DDP kernels hold in one warp (one ridge per lane). */
#define CCC_DDP_M_MAX 32
#define CCC_DDP_MAX_ALPHA 16

/* ---- nmpc_ddp::DDPSolver<>::Configuration + nmpc_ddp::BoxQP<>::Configuration ------------
 * Replaces: ddp_solver_->config() as set at reference src/DdpCentroidal.cpp:197-201,
 * src/DdpSingleRigidBody.cpp:267-271, include/CCC/DdpZmp.h:279-281 and
 * tests/src/TestDdpCentroidal.cpp:116 (max_iter).  Defaults: ccc_ddp_config_default(). */
typedef struct
{
  int32_t with_input_constraint; /* 1: BoxQP per stage (Centroidal, SRB); 0: plain LLT (Zmp) */
  int32_t max_iter;              /* DDP iterations (procOnce calls) */
  int32_t reg_type;              /* only 1 (Quu + lambda I) is implemented */
  int32_t n_alpha;               /* entries of alpha[] used by the line search */
  double initial_lambda, initial_dlambda, lambda_factor, lambda_min, lambda_max;
  double k_rel_norm_thre, lambda_thre;
  double cost_update_ratio_thre, cost_update_thre;
  double alpha[CCC_DDP_MAX_ALPHA]; /* 10^linspace(0,-3,11), evaluated on the host */
  int32_t boxqp_max_iter;
  int32_t reserved0;
  double boxqp_grad_thre, boxqp_rel_improve_thre, boxqp_step_factor, boxqp_min_step, boxqp_armijo;
} ccc_ddp_config_t;

/* Fill `cfg` with nmpc_ddp's defaults. */
void ccc_ddp_config_default(ccc_ddp_config_t * cfg);

/* ---- per-problem outputs of a DDP solve ---------------------------------------------------
 * Replaces: ddp_solver_->controlData().{x_list,u_list,cost_list} and
 * ddp_solver_->traceDataList() (reference src/DdpCentroidal.cpp:236,
 * tests/src/TestDdpCentroidal.cpp:102,129).  Any pointer may be NULL = not wanted. */
typedef struct
{
  double * x;      /* [B][N+1][nx]   optimal state trajectory */
  double * u;      /* [B][N][m_max]  optimal inputs, entries j >= m_k are 0 */
  double * cost;   /* [B]            sum of the accepted trajectory's cost_list */
  int32_t * iters; /* [B]            traceDataList().back().iter */
  int32_t * status; /* [B]           last procOnce return: 1 converged, 0 max_iter hit, -1 lambda > lambda_max */
  /* parity trace (optional) */
  int32_t trace_len;   /* slots per problem in alpha_idx[] / lambda_trace[] */
  int32_t reserved0;
  int8_t * alpha_idx;  /* [B][trace_len] accepted alpha index of iteration i+1; -1 = line search
                          failed; -2 = terminated on small gradient before the line search;
                          -3 = backward pass gave up (lambda_max); -4 = slot not reached */
  double * lambda_trace; /* [B][trace_len] lambda after iteration i+1 */
  uint32_t * clamped;  /* [B][N] bit j set = input j clamped by BoxQP in the last completed backward pass */
} ccc_ddp_result_t;

/* ---- CCC::DdpCentroidal ---------------------------------------------------------------------
 * Flat form of what DdpCentroidal::planOnce (reference src/DdpCentroidal.cpp:213-237) reads
 * through its two std::function callbacks, sampled at t_k = current_time + k*dt:
 *  MotionParam.contact_list -> m / ridge / vertex tables (order of
 *    src/DdpCentroidal.cpp:49-60: contact, vertex, ridge);
 *  RefData.pos              -> ref_pos (k = 0..N; the N-th entry feeds terminalCost).
 * `n_sched` distinct schedules are shared by the batch; problem b uses sched_id[b]. */
typedef struct
{
  int32_t horizon_steps; /* N */
  int32_t batch;         /* B */
  int32_t n_sched;       /* S */
  int32_t m_max;         /* row stride of ridge/vertex/u tables, <= CCC_DDP_M_MAX */
  double dt;             /* horizon_dt [s] */
  double mass;           /* [kg] */
  const int32_t * sched_id; /* [B] */
  const int32_t * m;        /* [S][N]   input dimension of each stage (0 = flight) */
  const double * ridge;     /* [S][N][m_max][3] friction-pyramid ridge of input j */
  const double * vertex;    /* [S][N][m_max][3] contact vertex input j acts on */
  const double * ref_pos;   /* [S][N+1][3] */
  double w_run[10];  /* WeightParam: running_pos(3), running_linear_momentum(3), running_angular_momentum(3), running_force */
  double w_term[9];  /* terminal_pos(3), terminal_linear_momentum(3), terminal_angular_momentum(3) */
  double u_lo, u_hi; /* force_scale_limits_ (reference include/CCC/DdpCentroidal.h:364) */
  const double * x0;     /* [B][9]  InitialParam::toState: pos, mass*vel, angular_momentum */
  const double * u_init; /* [B][N][m_max] warm start, or NULL = zeros (src/DdpCentroidal.cpp:221-229) */
} ccc_ddp_centroidal_batch_t;

typedef struct ccc_ddp_centroidal_ws ccc_ddp_centroidal_ws_t;

/* Allocate the device workspace (gain lists, candidate trajectories, staging buffers) for up
 * to `max_batch` problems of `horizon_steps` stages and `max_sched` schedules on the current
 * CUDA device.  Returns NULL on failure (see ccc_last_error()). */
ccc_ddp_centroidal_ws_t * ccc_ddp_centroidal_create(int32_t horizon_steps, int32_t max_batch, int32_t max_sched);
void ccc_ddp_centroidal_destroy(ccc_ddp_centroidal_ws_t * ws);

/* Solve `batch->batch` independent DdpCentroidal problems: rollout, then up to
 * cfg->max_iter iterations of {derivatives, backward pass with BoxQP, line search, lambda
 * schedule} per problem, one warp per problem; a batch of at most one problem per SM (the reference's own caller
 * solves ONE problem per control tick) gets a thread block of eight warps per problem instead, with the line-search
 * candidates rolled out concurrently.  Results do not depend on the batch size or the kernel chosen.
 * CCC_MEM_HOST calls whose arrays fit 1 MB move through one pinned staging block (one copy in, one copy out).
 * Replaces: nmpc_ddp::DDPSolver<9,Dynamic>::solve as called at reference
 * src/DdpCentroidal.cpp:229,233 together with the DdpProblem callbacks at :32-177. */
int32_t ccc_ddp_centroidal_solve(ccc_ddp_centroidal_ws_t * ws,
                                 const ccc_ddp_centroidal_batch_t * batch,
                                 const ccc_ddp_config_t * cfg,
                                 ccc_ddp_result_t * result,
                                 int32_t mem,
                                 void * stream);

/* Number of kernels the last solve on this workspace launched (bench.py's gpu_launches). */
int32_t ccc_ddp_centroidal_last_launches(const ccc_ddp_centroidal_ws_t * ws);

/* ---- closed loop (receding horizon) around CCC::DdpCentroidal, resident on the device -------------------
 * The control loop of the reference's own test (tests/src/TestDdpCentroidal.cpp:94-150) for a batch of plants:
 * every control cycle plans with the plant's current state (warm start = the previous plan, zeroed where a
 * stage's input dimension changed, :102-114; max_iter = max_iter_later after the first cycle, :116), applies
 * the first stage's wrench to the plant (CentroidalSim, tests/src/SimModels.h:233-332: exact zero-order hold
 * of a point mass + angular momentum) and moves on by sim_dt.  No host round trip between cycles.
 * The contact / reference schedule is given on the plant's time grid: stage k of cycle t reads grid entry
 * t + k * stride, stride = horizon_dt / sim_dt (an integer), i.e. the callbacks sampled at
 * time = current_time0 + entry * sim_dt. */
typedef struct
{
  int32_t horizon_steps; /* N */
  int32_t batch;         /* B */
  int32_t n_sched;       /* S */
  int32_t m_max;         /* row stride of ridge/vertex, <= CCC_DDP_M_MAX */
  double dt;             /* horizon_dt [s] */
  double mass;           /* [kg] */
  double sim_dt;         /* control / plant period [s] */
  int32_t ticks;         /* control cycles */
  int32_t stride;        /* horizon_dt / sim_dt */
  int32_t grid_len;      /* time-grid entries per schedule, >= ticks - 1 + horizon_steps * stride + 1 */
  int32_t max_iter_later; /* DDP iterations per cycle after the first one (the first uses cfg->max_iter) */
  const int32_t * sched_id; /* [B] */
  const int32_t * m;        /* [S][grid_len] */
  const double * ridge;     /* [S][grid_len][m_max][3] */
  const double * vertex;    /* [S][grid_len][m_max][3] */
  const double * ref_pos;   /* [S][grid_len][3] */
  double w_run[10];
  double w_term[9];
  double u_lo, u_hi;
  const double * plant0;    /* [B][9] initial plant state: position, velocity, angular momentum */
  int32_t disturb_tick;     /* -1: none; else the velocity impulse is added after the plant step of this cycle */
  int32_t reserved0;
  double disturb_vel[3];    /* [m/s] */
} ccc_ddp_centroidal_loop_t;

typedef struct
{
  double * plant;  /* [B][ticks+1][9] plant state at the start of every cycle and after the last one */
  double * u0;     /* [B][ticks][m_max] force scales applied in every cycle, or NULL */
  int32_t * iters; /* [B][ticks] DDP iterations of every cycle, or NULL */
} ccc_ddp_centroidal_loop_result_t;

/* Runs on the workspace of ccc_ddp_centroidal_create(horizon_steps, >= batch, >= n_sched).  Host pointers
 * (CCC_MEM_HOST, synchronous) or device pointers (CCC_MEM_DEVICE, enqueued on `stream`). */
int32_t ccc_ddp_centroidal_closed_loop(ccc_ddp_centroidal_ws_t * ws,
                                       const ccc_ddp_centroidal_loop_t * loop,
                                       const ccc_ddp_config_t * cfg,
                                       ccc_ddp_centroidal_loop_result_t * result,
                                       int32_t mem,
                                       void * stream);

/* ---- CCC::DdpSingleRigidBody -----------------------------------------------------------------
 * Flat form of what DdpSingleRigidBody::planOnce (reference src/DdpSingleRigidBody.cpp:283-307) reads
 * through its callbacks at t_k = current_time + k*dt: MotionParam.{contact_list, inertia_mat}
 * (include/CCC/DdpSingleRigidBody.h:25-34) and RefData.{pos, ori} (:37-46).  State (12):
 * pos, ZYX Euler angles (Z, Y, X order), linear velocity, angular velocity
 * (InitialParam::toState, src/DdpSingleRigidBody.cpp:255-260). */
typedef struct
{
  int32_t horizon_steps; /* N */
  int32_t batch;         /* B */
  int32_t n_sched;       /* S */
  int32_t m_max;         /* row stride of ridge/vertex/u tables, <= CCC_DDP_M_MAX */
  double dt;             /* horizon_dt [s] */
  double mass;           /* [kg] */
  const int32_t * sched_id; /* [B] */
  const int32_t * m;        /* [S][N] */
  const double * ridge;     /* [S][N][m_max][3] */
  const double * vertex;    /* [S][N][m_max][3] */
  const double * inertia;   /* [S][N][9] inertia_mat, row-major, symmetric positive definite */
  const double * ref;       /* [S][N+1][6] pos, ori */
  double w_run[13];  /* running_pos, running_ori, running_linear_vel, running_angular_vel (3 each), running_force */
  double w_term[12]; /* terminal_pos, terminal_ori, terminal_linear_vel, terminal_angular_vel */
  double u_lo, u_hi; /* force_scale_limits_ */
  const double * x0;     /* [B][12] */
  const double * u_init; /* [B][N][m_max] or NULL */
} ccc_ddp_srb_batch_t;

typedef struct ccc_ddp_srb_ws ccc_ddp_srb_ws_t;

ccc_ddp_srb_ws_t * ccc_ddp_srb_create(int32_t horizon_steps, int32_t max_batch, int32_t max_sched);
void ccc_ddp_srb_destroy(ccc_ddp_srb_ws_t * ws);

/* Replaces: nmpc_ddp::DDPSolver<12,Dynamic>::solve as called at reference
 * src/DdpSingleRigidBody.cpp:299,303 together with the DdpProblem callbacks at :41-245.
 * result->x is [B][N+1][12]. */
int32_t ccc_ddp_srb_solve(ccc_ddp_srb_ws_t * ws,
                          const ccc_ddp_srb_batch_t * batch,
                          const ccc_ddp_config_t * cfg,
                          ccc_ddp_result_t * result,
                          int32_t mem,
                          void * stream);
int32_t ccc_ddp_srb_last_launches(const ccc_ddp_srb_ws_t * ws);

/* ---- CCC::DdpZmp --------------------------------------------------------------------------------
 * Flat form of what DdpZmp::planOnce (reference src/DdpZmp.cpp:152-174) reads through ref_data_func at
 * t_k = current_time + k*dt: RefData.{zmp, com_z} (include/CCC/DdpZmp.h:20-28).  State (6):
 * (pos_x, vel_x, pos_y, vel_y, pos_z, vel_z) (InitialParam::toState, src/DdpZmp.cpp:145-150); input (3):
 * (zmp_x, zmp_y, force_z); no input limits (cfg->with_input_constraint must be 0). */
typedef struct
{
  int32_t horizon_steps; /* N */
  int32_t batch;         /* B */
  int32_t n_sched;       /* S */
  int32_t reserved0;
  double dt;             /* horizon_dt [s] */
  double mass;           /* [kg] */
  const int32_t * sched_id; /* [B] */
  const double * ref_zmp;   /* [S][N+1][3] */
  const double * com_z;     /* [S][N+1] */
  double w[6]; /* WeightParam: running_com_pos_z, running_zmp, running_force_z, terminal_com_pos_xy, terminal_com_pos_z, terminal_com_vel */
  const double * x0;     /* [B][6] */
  const double * u_init; /* [B][N][3] or NULL = zeros */
} ccc_ddp_zmp_batch_t;

typedef struct ccc_ddp_zmp_ws ccc_ddp_zmp_ws_t;
ccc_ddp_zmp_ws_t * ccc_ddp_zmp_create(int32_t horizon_steps, int32_t max_batch, int32_t max_sched);
void ccc_ddp_zmp_destroy(ccc_ddp_zmp_ws_t * ws);
/* Replaces: nmpc_ddp::DDPSolver<6,3>::solve as called at reference src/DdpZmp.cpp:160-171 together with the
 * DdpProblem callbacks at :8-143.  result->x is [B][N+1][6], result->u is [B][N][3]; clamped is all zero. */
int32_t ccc_ddp_zmp_solve(ccc_ddp_zmp_ws_t * ws,
                          const ccc_ddp_zmp_batch_t * batch,
                          const ccc_ddp_config_t * cfg,
                          ccc_ddp_result_t * result,
                          int32_t mem,
                          void * stream);
int32_t ccc_ddp_zmp_last_launches(const ccc_ddp_zmp_ws_t * ws);

/* ---- QpSolverCollection::QpSolver::solve(QpCoeff &) -------------------------------------------
 * Batched strictly convex dense QP with the matrices shared by the batch:
 *     min 0.5 x'Qx + c'x   s.t.  A x = b,  C x <= d
 * Replaces: qp_solver_->solve(qp_coeff_) at reference src/LinearMpcZmp.cpp:69 (Q = I, c = 0, no
 * equality, C = [-B_seq; B_seq]) and src/IntrinsicallyStableMpc.cpp:93 (Q = w_vel I + w_zmp P'P, one
 * equality, C = [-P; P]).  QpCoeff fields: obj_mat_ = Q, obj_vec_ = c, eq_mat_/eq_vec_ = A/b,
 * ineq_mat_/ineq_vec_ = C/d; x_min_/x_max_ are +-1e10 on those two paths and not represented.
 * src/LinearMpcXY.cpp:181 (Q = B_seq' W B_seq + w I, one equality per contact stage, finite bounds
 * x_min_/x_max_): the caller appends the bounds as inequality rows  -e_j x <= -x_min,  e_j x <= x_max.
 * Limits of the kernel: n <= 256, n_eq <= n, n_eq + n_ineq <= 512 (n <= 128) or 1024 (n > 128). */
typedef struct
{
  int32_t n, n_eq, n_ineq, batch;
  const double * Q; /* [n][n]       shared; NULL = reuse Q, A, C (and their factorisation) of the previous call on
                       this workspace: what a controller does that keeps one QpCoeff structure over its ticks */
  const double * A; /* [n_eq][n]    shared (NULL if n_eq = 0) */
  const double * C; /* [n_ineq][n]  shared */
  const double * c; /* [B][n]       or NULL = 0 */
  const double * b; /* [B][n_eq]    (NULL if n_eq = 0) */
  const double * d; /* [B][n_ineq] */
} ccc_qp_batch_t;

typedef struct
{
  double * x;         /* [B][n] minimiser */
  int32_t * iters;    /* [B] constraints picked by the dual active-set loop */
  int32_t * status;   /* [B] 0 solved, 1 infeasible, 2 iteration limit, 3 Q not positive definite (4 is no longer produced: a full active set takes the dual step of Goldfarb-Idnani) */
  int32_t * n_active; /* [B] size of the final active set (equalities included) */
  int32_t * active;   /* [B][n] ids in activation order (0..n_eq-1 equalities, n_eq + i inequality i), -1 padded */
} ccc_qp_result_t;

typedef struct ccc_qp_ws ccc_qp_ws_t;
ccc_qp_ws_t * ccc_qp_create(int32_t n, int32_t n_eq, int32_t n_ineq, int32_t max_batch);
void ccc_qp_destroy(ccc_qp_ws_t * ws);
int32_t ccc_qp_solve(ccc_qp_ws_t * ws, const ccc_qp_batch_t * batch, ccc_qp_result_t * result, int32_t mem, void * stream);
int32_t ccc_qp_last_launches(const ccc_qp_ws_t * ws);

/* Matrix groups: problem b uses Q / A of group group_id[b] (batch->Q is [G][n][n], batch->A [G][n_eq][n]; C stays
 * shared), one factorisation per group per call.  For sweeps in which the QP matrices themselves vary: LinearMpcXY
 * over contact schedules (ccc_linear_mpc_xy_solve does this internally), the wrench distribution of
 * PreviewControlCentroidal (ForceColl::WrenchDistribution::run as called at reference
 * src/PreviewControlCentroidal.cpp:127-128: the grasp matrix depends on each problem's CoM position).
 * ccc_qp_solve(ws, batch, ...) == ccc_qp_solve_grouped(ws, batch, 1, NULL, ...). */
ccc_qp_ws_t * ccc_qp_create_grouped(int32_t n, int32_t n_eq, int32_t n_ineq, int32_t max_batch, int32_t max_groups);
int32_t ccc_qp_solve_grouped(ccc_qp_ws_t * ws, const ccc_qp_batch_t * batch, int32_t n_groups, const int32_t * group_id,
                             ccc_qp_result_t * result, int32_t mem, void * stream);

/* ---- CCC::LinearMpcXY::planOnce on the device, over a sweep of contact / reference schedules -------------
 * Everything LinearMpcXY::planOnce does after sampling its callbacks (reference src/LinearMpcXY.cpp:102-181),
 * for B initial states that use S sampled schedules (problem b uses sched_id[b]):
 *  per stage     Model::Model (:59-83) and StateSpaceModel::calcDiscMatrix (include/CCC/StateSpaceModel.h:164-216);
 *                the model's A is nilpotent (A^3 = 0, no offset vector), so the zero-order hold is the finite
 *                polynomial Ad = I + A dt + A^2 dt^2/2, Bd = (I dt + A dt^2/2 + A^2 dt^3/6) B;
 *  per schedule  VariantSequentialExtension<6>::setup (include/CCC/VariantSequentialExtension.h:110-186): A_seq, B_seq;
 *                obj_mat = B_seq' W B_seq + w_force I (src :141-144) as an FP64 tensor-core GEMM; eq_mat (:149-176);
 *  per problem   obj_vec = -B_seq' W (ref_output_seq - A_seq x) (:145-146), eq_vec, x_min / x_max (:177-178);
 *  QP            qp_solver_->solve(qp_coeff_) (:181) on the ccc_qp engine, one factorisation per schedule.
 * The schedules of one call must agree in the QP's shape: sum_k m[s][k] = n and the number of contact stages
 * (m[s][k] > 0) = n_eq for every s; group a sweep by shape on the host.  Stage tables as for DdpCentroidal. */
typedef struct
{
  int32_t horizon_steps; /* N */
  int32_t batch;         /* B */
  int32_t n_sched;       /* S */
  int32_t m_max;         /* row stride of ridge / vertex, any value >= max m */
  double dt;             /* horizon_dt [s] */
  double mass;           /* [kg] */
  const int32_t * sched_id;     /* [B] */
  const int32_t * m;            /* [S][N] ridges of each stage's contacts (0: no contact) */
  const double * ridge;         /* [S][N][m_max][3] */
  const double * vertex;        /* [S][N][m_max][3] */
  const double * com_z;         /* [S][N] MotionParam.com_z */
  const double * total_force_z; /* [S][N] MotionParam.total_force_z */
  const double * ref_output;    /* [S][N][6] RefData::toOutput(mass) (src :33-38) */
  double w_output[6];   /* WeightParam::outputWeight of one stage (src :45-50) */
  double w_force;       /* WeightParam.force */
  double force_lo, force_hi; /* force_range_ (src :91) */
  const double * x0;    /* [B][6] InitialParam::toState(mass) (src :26-31) */
} ccc_linear_mpc_xy_batch_t;

typedef struct
{
  double * u;         /* [B][n] force scales of every stage; the first m[s][0] entries are planOnce's return value */
  int32_t * iters;    /* [B] as ccc_qp_result_t */
  int32_t * status;   /* [B] */
  int32_t * n_active; /* [B] */
  int32_t * active;   /* [B][n] */
  /* intermediate results (optional; what the parity tests compare stage by stage) */
  double * A_seq;     /* [S][6N][6] */
  double * B_seq;     /* [S][6N][n] */
  double * obj_mat;   /* [S][n][n]  */
  double * obj_vec;   /* [B][n]     */
} ccc_linear_mpc_xy_result_t;

typedef struct ccc_linear_mpc_xy_ws ccc_linear_mpc_xy_ws_t;
/* n = total input dimension, n_eq = contact stages of every schedule of the calls to come (n <= 256). */
ccc_linear_mpc_xy_ws_t * ccc_linear_mpc_xy_create(int32_t horizon_steps, int32_t n, int32_t n_eq, int32_t max_batch, int32_t max_sched);
void ccc_linear_mpc_xy_destroy(ccc_linear_mpc_xy_ws_t * ws);
int32_t ccc_linear_mpc_xy_solve(ccc_linear_mpc_xy_ws_t * ws,
                                const ccc_linear_mpc_xy_batch_t * batch,
                                ccc_linear_mpc_xy_result_t * result,
                                int32_t mem,
                                void * stream);
int32_t ccc_linear_mpc_xy_last_launches(const ccc_linear_mpc_xy_ws_t * ws);

/* ---- schedule compiler: footstep plans -> stage tables -----------------------------------------------
 * Batched, stateless counterpart of the reference tests' FootstepManager (tests/src/FootstepManager.h): what
 * FootstepManager::update(current_time) (:147-209) builds — the knot lists ref_zmp_list_ / ref_footstance_list_ —
 * and what refZmp(t) (:228-237) / zmpLimits(t) (:242-254) read from them, sampled on the horizon grid
 * t_k = current_time + k horizon_dt (+ eps_reps x 1e-6: the reference adds its epsilon_t once in refZmp / zmpLimits
 * and once more in make*RefData, :356-380), for P plans at once.  "Stateless": the foot stance at current_time is
 * the initial stance with every footstep whose swing_end_time <= current_time applied and the pending list is the
 * footsteps whose transit_end_time >= current_time — what a manager updated once per control cycle holds. */
typedef struct
{
  int32_t n_plans;       /* P */
  int32_t max_steps;     /* F: row stride of the footstep arrays */
  int32_t horizon_steps; /* N */
  int32_t eps_reps;      /* 1e-6 is added this many times to every sample time (1 or 2, see above) */
  double horizon_dt;
  double manager_horizon;  /* FootstepManager::horizon_duration_ (:461, default 10 s) */
  double foot_size[2];     /* FootstepManager::foot_size_ (:464) */
  const double * current_time; /* [P] */
  const double * stance0;      /* [P][2][2] initial foot positions: Left, Right (:136-137) */
  const int32_t * n_steps;     /* [P] */
  const int32_t * foot;        /* [P][F] 0 = Left, 1 = Right */
  const double * pos;          /* [P][F][2] */
  const double * times;        /* [P][F][4] transit_start, swing_start, swing_end, transit_end (Footstep, :36-77) */
} ccc_footstep_plans_t;

typedef struct
{
  double * ref_zmp; /* [P][N][2] */
  double * lim_min; /* [P][N][2] */
  double * lim_max; /* [P][N][2] */
} ccc_zmp_tables_t;

/* Stateless: no workspace.  Host pointers (synchronous) or device pointers (enqueued on `stream`). */
int32_t ccc_footstep_compile(const ccc_footstep_plans_t * plans, ccc_zmp_tables_t * tables, int32_t mem, void * stream);

/* ---- CCC::LinearMpcZmp / CCC::IntrinsicallyStableMpc planOnce on stage tables, on the device -----------
 * procOnce of both methods for B problems x 2 axes whose reference data are rows of compiled tables (problem b reads
 * plan plan_id[b]): the QP vectors are assembled on the device, the 2B one-dimensional QPs go through the ccc_qp
 * engine, the planned ZMP is post-processed there.
 *  method 0, LinearMpcZmp1d::procOnce (src/LinearMpcZmp.cpp:46-81): d = [A_seq x0 - lo; -A_seq x0 + hi], planned ZMP
 *    = clamp(C (A_d x0 + B_d u_0), lo_0, hi_0) with the control_dt model written out as the reference does (:71-79);
 *  method 1, IntrinsicallyStableMpc1d::procOnce (src/IntrinsicallyStableMpc.cpp:63-104): b = cp - z0,
 *    c = w_zmp P'(z0 1 - z_ref), d = [-lo + z0; hi - z0], planned ZMP = clamp(z0 + control_dt u_0, lo_0, hi_0).
 * The QP matrices (batch invariant, from the host-side setup of the controller: InvariantSequentialExtension,
 * P, the stability constraint) are passed once per call. */
typedef struct
{
  int32_t method;        /* 0: LinearMpcZmp, 1: IntrinsicallyStableMpc */
  int32_t horizon_steps; /* N = number of QP variables */
  int32_t batch;         /* B problems (2B QPs: x axes first, then y axes) */
  int32_t n_plans;       /* P */
  double control_dt;     /* resolved (> 0) */
  double com_height_over_g; /* method 0: -C(0,2) of ComZmpModelJerkInput = h / g */
  double weight_zmp;     /* method 1 */
  const double * Q;      /* [N][N] obj_mat_ */
  const double * A;      /* [1][N] eq_mat_ (method 1) or NULL */
  const double * C;      /* [2N][N] ineq_mat_ */
  const double * A_seq;  /* [N][3] (method 0) */
  const double * P;      /* [N][N] (method 1) */
  const int32_t * plan_id; /* [B] */
  const double * state;  /* method 0: [B][2][3] (pos, vel, acc) per axis; method 1: [B][2][2] (capture point, planned zmp) per axis */
  ccc_zmp_tables_t tables; /* device (CCC_MEM_DEVICE) or host (CCC_MEM_HOST) tables of the P plans */
} ccc_zmp_mpc_batch_t;

typedef struct
{
  double * planned_zmp; /* [B][2] */
  int32_t * iters;      /* [2B] active-set iterations of every QP, or NULL */
  int32_t * status;     /* [2B] or NULL */
} ccc_zmp_mpc_result_t;

typedef struct ccc_zmp_mpc_ws ccc_zmp_mpc_ws_t;
ccc_zmp_mpc_ws_t * ccc_zmp_mpc_create(int32_t method, int32_t horizon_steps, int32_t max_batch, int32_t max_plans);
void ccc_zmp_mpc_destroy(ccc_zmp_mpc_ws_t * ws);
int32_t ccc_zmp_mpc_plan(ccc_zmp_mpc_ws_t * ws, const ccc_zmp_mpc_batch_t * batch, ccc_zmp_mpc_result_t * result, int32_t mem, void * stream);
int32_t ccc_zmp_mpc_last_launches(const ccc_zmp_mpc_ws_t * ws);

/* ---- closed-form ZMP controllers ---------------------------------------------------------------------
 * CCC::DcmTracking::planOnce (reference src/DcmTracking.cpp:7-48, Englsberger et al. 2013) and
 * CCC::FootGuidedControl::planOnce (src/FootGuidedControl.cpp:11-92, Sugihara 2017 / Kojio 2019) for a batch of
 * initial parameters; problem b reads the reference data of plan plan_id[b].  Both are a handful of exp() and
 * arithmetic per problem: elementwise kernels, no workspace.  exp() is the CUDA library's (<= 1 ulp from the
 * correctly rounded value): results agree with a libm evaluation to ~1e-15 relative, not bit for bit.
 * The reference throws on inconsistent times (a switching time in the past, a negative transition duration, a
 * transition that does not end in the future); host-buffer calls return CCC_ERR_INVALID for those, device-pointer
 * calls write NaN for the offending problems. */
typedef struct
{
  int32_t batch;     /* B */
  int32_t n_plans;   /* P */
  int32_t max_knots; /* K: row stride of the switching lists */
  int32_t reserved0;
  double omega;         /* sqrt(g / com_height) (include/CCC/DcmTracking.h:51) */
  double feedback_gain; /* DcmTracking::feedback_gain_ (default 2.0) */
  const int32_t * plan_id;     /* [B] */
  const double * dcm;          /* [B][2] InitialParam: current DCM */
  const double * current_time; /* [P] */
  const double * current_zmp;  /* [P][2] RefData.current_zmp */
  const int32_t * n_knots;     /* [P] entries of RefData.time_zmp_list (>= 0) */
  const double * knot_time;    /* [P][K] ascending (the map's keys) */
  const double * knot_zmp;     /* [P][K][2] */
} ccc_dcm_tracking_batch_t;
int32_t ccc_dcm_tracking_plan(const ccc_dcm_tracking_batch_t * batch, double * control_zmp /* [B][2] */, int32_t mem, void * stream);

typedef struct
{
  int32_t batch;   /* B */
  int32_t n_plans; /* P */
  double omega;    /* sqrt(g / com_height) (include/CCC/FootGuidedControl.h:51) */
  const int32_t * plan_id;           /* [B] */
  const double * capture_point;      /* [B][2] InitialParam */
  const double * current_time;       /* [P] */
  const double * transit_start_zmp;  /* [P][2] RefData (include/CCC/FootGuidedControl.h:78-91) */
  const double * transit_end_zmp;    /* [P][2] */
  const double * transit_start_time; /* [P] */
  const double * transit_duration;   /* [P] */
} ccc_foot_guided_batch_t;
int32_t ccc_foot_guided_plan(const ccc_foot_guided_batch_t * batch, double * planned_zmp /* [B][2] */, int32_t mem, void * stream);

/* ---- CCC::SingularPreviewControlZmp ---------------------------------------------------------------------
 * Urata's singular LQ preview regulation (reference src/SingularPreviewControlZmp.cpp:23-57, x 2 axes :59-88) for a batch:
 * problem b = (planned ZMP, CoM position, CoM velocity) per axis, on the reference-ZMP sequence of plan plan_id[b]
 * (ref_zmp_func sampled at current_time + i horizon_dt, i < N, as planOnce does at :14-19).  The feed-forward term of a
 * plan (the backward recursion of equations (25)-(27) over the sequence) is evaluated once per plan and axis, the feedback
 * term and the ZMP update per problem.  Arithmetic: the reference's expressions term by term, with std::pow(b, 2) as b * b
 * (no transcendental on the path: results equal a host evaluation of the same expressions bit for bit). */
typedef struct
{
  int32_t batch;         /* B */
  int32_t n_plans;       /* P */
  int32_t horizon_steps; /* N = ceil(horizon_duration / horizon_dt) (include/CCC/SingularPreviewControlZmp.h:48) */
  int32_t reserved0;
  double omega;      /* sqrt(g / com_height) */
  double horizon_dt; /* discretisation step of the horizon */
  double control_dt; /* step used to integrate the planned ZMP (planOnce's control_dt) */
  const int32_t * plan_id; /* [B] */
  const double * state;    /* [B][2][3] per axis: planned_zmp, pos, vel (InitialParam, :30-31) */
  const double * ref_zmp;  /* [P][N][2] */
} ccc_singular_preview_batch_t;
int32_t ccc_singular_preview_plan(const ccc_singular_preview_batch_t * batch, double * planned_zmp /* [B][2] */, int32_t mem, void * stream);

/* ---- CCC::StepMpc ----------------------------------------------------------------------------------------
 * Xin's step-to-step MPC (reference src/StepMpc.cpp:27-193, x 2 axes :195-247) for a batch: problem b = CoM position and
 * velocity per axis on the reference data of plan plan_id[b] (support phases with their ZMP and end time,
 * include/CCC/StepMpc.h:36-55).  One unknown per phase (at most CCC_STEP_MPC_MAX_ELEMENTS); a thread per problem and axis
 * propagates the closed-form step model over the phases (StepModel, src/StepMpc.cpp:8-25: cosh / sinh of omega x duration),
 * accumulates the weighted least-squares system term by term as the reference does (every term is a sum of rank-one
 * updates with rows of the condensed output matrix) and solves it by Gaussian elimination with partial pivoting (the
 * reference: Eigen colPivHouseholderQr).  The systems are ill conditioned by construction (cosh of up to three seconds of
 * horizon squared: condition numbers of 1e8 - 1e11), so two solvers agree to ~1e-6, not to rounding; the reference's own
 * test asks for 1e-2 (tests/src/TestStepMpc.cpp:138-141). */
#define CCC_STEP_MPC_MAX_ELEMENTS 16
typedef struct
{
  int32_t batch;        /* B */
  int32_t n_plans;      /* P */
  int32_t max_elements; /* K <= CCC_STEP_MPC_MAX_ELEMENTS: row stride of the element lists */
  int32_t reserved0;
  double com_height;
  /* StepMpc1d::WeightParam (include/CCC/StepMpc.h:84-117) */
  double w_free_zmp, w_fixed_zmp, w_double_support, w_pos, w_vel, w_capture_point_abs, w_capture_point_rel;
  const int32_t * plan_id;      /* [B] */
  const double * x_pos;         /* [B][2] InitialParam.pos */
  const double * x_vel;         /* [B][2] InitialParam.vel */
  const double * current_time;  /* [P] */
  const int32_t * n_elements;   /* [P] 1 .. K */
  const int32_t * single;       /* [P][K] Element.is_single_support */
  const double * zmp;           /* [P][K][2] Element.zmp */
  const double * end_time;      /* [P][K] Element.end_time */
} ccc_step_mpc_batch_t;
typedef struct
{
  double * current_zmp;   /* [B][2] PlannedData.current_zmp */
  double * next_foot_zmp; /* [B][2] PlannedData.next_foot_zmp (0 where has_next is 0) */
  int32_t * has_next;     /* [B] 1 if a single-support phase of the next foot lies in the horizon (std::optional has a value) */
} ccc_step_mpc_result_t;
int32_t ccc_step_mpc_plan(const ccc_step_mpc_batch_t * batch, const ccc_step_mpc_result_t * result, int32_t mem, void * stream);

/* ---- CCC::PreviewControl<3,1,1>::calcOptimalInput ------------------------------------------------
 * Batched online part of preview control:  u[b] = -K x[b] + F ref_seq[b]   (gains K (1x3), F (1xN) shared).
 * Replaces: PreviewControl::calcOptimalInput (reference include/CCC/PreviewControl.h:86-89) as called from
 * PreviewControlZmp1d::procOnce (src/PreviewControlZmp.cpp:37).  The gains come from the host-side setup
 * (DARE, include/CCC/PreviewControl.h:93-172).  Stateless: no workspace. */
int32_t ccc_preview_input(int32_t batch,
                          int32_t horizon_steps,
                          const double * K,       /* [3]      */
                          const double * F,       /* [N]      */
                          const double * x,       /* [B][3]   */
                          const double * ref_seq, /* [B][N]   */
                          double * u,             /* [B]      */
                          int32_t mem,
                          void * stream);

/* ---- misc ------------------------------------------------------------------------------------ */
int32_t ccc_abi_version(void);
int32_t ccc_device_count(void);
const char * ccc_last_error(void);
/* Measurement aid (no reference counterpart): FP64 fma throughput of the current device in TFLOP/s, best of `reps`
 * launches of a pure-DFMA kernel timed with CUDA events on `stream`; the denominator of bench.py's roofline.fp64. */
double ccc_fp64_peak_tflops(int32_t reps, void * stream);

#ifdef __cplusplus
}
#endif
#endif /* CCC_B200_H */
