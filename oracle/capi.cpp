// oracle/capi.cpp — C entry points of the CPU oracle (ctypes-loadable).
//
// TEST INFRASTRUCTURE ONLY: called from tests/, from bench.py's cpu_baseline and
// --impl reference legs, and from __graft_entry__.smoke() as the checker.  The product
// library (libccc_b200.so) never links or loads this file.
//
// The batch/config/result structs are the plain-data interface structs of
// include/ccc_b200.h, so the same buffers can be handed to the oracle and to the engine.
#include "../include/ccc_b200.h"
#include "centroidal.hpp"
#include "srb.hpp"
#include "qp.hpp"
#include "xy.hpp"
#include "zmp.hpp"

#include <atomic>
#include <thread>

using namespace oracle;

namespace
{
std::atomic<long long> g_stats[6];

DdpConfig toConfig(const ccc_ddp_config_t * c)
{
  DdpConfig cfg;
  cfg.with_input_constraint = c->with_input_constraint != 0;
  cfg.max_iter = c->max_iter;
  cfg.reg_type = c->reg_type;
  cfg.initial_lambda = c->initial_lambda;
  cfg.initial_dlambda = c->initial_dlambda;
  cfg.lambda_factor = c->lambda_factor;
  cfg.lambda_min = c->lambda_min;
  cfg.lambda_max = c->lambda_max;
  cfg.k_rel_norm_thre = c->k_rel_norm_thre;
  cfg.lambda_thre = c->lambda_thre;
  cfg.cost_update_ratio_thre = c->cost_update_ratio_thre;
  cfg.cost_update_thre = c->cost_update_thre;
  cfg.alpha_list.assign(c->alpha, c->alpha + c->n_alpha);
  cfg.boxqp.max_iter = c->boxqp_max_iter;
  cfg.boxqp.grad_thre = c->boxqp_grad_thre;
  cfg.boxqp.rel_improve_thre = c->boxqp_rel_improve_thre;
  cfg.boxqp.step_factor = c->boxqp_step_factor;
  cfg.boxqp.min_step = c->boxqp_min_step;
  cfg.boxqp.armijo = c->boxqp_armijo;
  return cfg;
}

template<class F>
void parallelFor(int n, int n_threads, F && f)
{
  if(n_threads <= 1)
  {
    for(int i = 0; i < n; i++) f(i);
    return;
  }
  std::atomic<int> next(0);
  std::vector<std::thread> pool;
  for(int t = 0; t < n_threads; t++)
    pool.emplace_back([&]() {
      for(;;)
      {
        int i = next.fetch_add(1);
        if(i >= n) break;
        f(i);
      }
    });
  for(auto & th : pool) th.join();
}

/** Write solver outputs of problem b into the result arrays. */
void storeResult(const DdpSolver & s, int b, int N, int nx, int m_max, ccc_ddp_result_t * r)
{
  if(r->x)
    for(int k = 0; k <= N; k++)
      for(int i = 0; i < nx; i++) r->x[(static_cast<size_t>(b) * (N + 1) + k) * nx + i] = s.x[k][i];
  if(r->u)
    for(int k = 0; k < N; k++)
      for(int j = 0; j < m_max; j++)
        r->u[(static_cast<size_t>(b) * N + k) * m_max + j] = j < static_cast<int>(s.u[k].size()) ? s.u[k][j] : 0.0;
  if(r->cost) r->cost[b] = DdpSolver::sum_seq(s.cost);
  if(r->iters) r->iters[b] = s.trace.back().iter;
  if(r->status) r->status[b] = s.retval;
  for(int i = 0; i < r->trace_len; i++)
  {
    size_t o = static_cast<size_t>(b) * r->trace_len + i;
    bool have = static_cast<size_t>(i + 1) < s.trace.size();
    if(r->alpha_idx) r->alpha_idx[o] = have ? static_cast<int8_t>(s.trace[i + 1].alpha_idx) : static_cast<int8_t>(-4);
    if(r->lambda_trace) r->lambda_trace[o] = have ? s.trace[i + 1].lambda : 0.0;
  }
  if(r->clamped)
    for(int k = 0; k < N; k++) r->clamped[static_cast<size_t>(b) * N + k] = s.clamped_mask[k];
}

void bindCentroidal(CentroidalProblem & p, const ccc_ddp_centroidal_batch_t * bt, int sched)
{
  const int N = bt->horizon_steps, mm = bt->m_max;
  p.N = N;
  p.dt = bt->dt;
  p.mass = bt->mass;
  p.m_max = mm;
  p.m_tab = bt->m + static_cast<size_t>(sched) * N;
  p.ridge = bt->ridge + static_cast<size_t>(sched) * N * mm * 3;
  p.vertex = bt->vertex + static_cast<size_t>(sched) * N * mm * 3;
  p.ref_pos = bt->ref_pos + static_cast<size_t>(sched) * (N + 1) * 3;
  for(int i = 0; i < 10; i++) p.w_run[i] = bt->w_run[i];
  for(int i = 0; i < 9; i++) p.w_term[i] = bt->w_term[i];
  p.u_lo = bt->u_lo;
  p.u_hi = bt->u_hi;
}
} // namespace

extern "C" {

/** nmpc_ddp defaults as recalled in SURVEY.md App. A (restated independently of the engine). */
void ccc_oracle_ddp_config_default(ccc_ddp_config_t * c)
{
  DdpConfig d;
  c->with_input_constraint = 0;
  c->max_iter = d.max_iter;
  c->reg_type = d.reg_type;
  c->n_alpha = static_cast<int32_t>(d.alpha_list.size());
  c->initial_lambda = d.initial_lambda;
  c->initial_dlambda = d.initial_dlambda;
  c->lambda_factor = d.lambda_factor;
  c->lambda_min = d.lambda_min;
  c->lambda_max = d.lambda_max;
  c->k_rel_norm_thre = d.k_rel_norm_thre;
  c->lambda_thre = d.lambda_thre;
  c->cost_update_ratio_thre = d.cost_update_ratio_thre;
  c->cost_update_thre = d.cost_update_thre;
  for(int i = 0; i < CCC_DDP_MAX_ALPHA; i++) c->alpha[i] = i < c->n_alpha ? d.alpha_list[i] : 0.0;
  c->boxqp_max_iter = d.boxqp.max_iter;
  c->reserved0 = 0;
  c->boxqp_grad_thre = d.boxqp.grad_thre;
  c->boxqp_rel_improve_thre = d.boxqp.rel_improve_thre;
  c->boxqp_step_factor = d.boxqp.step_factor;
  c->boxqp_min_step = d.boxqp.min_step;
  c->boxqp_armijo = d.boxqp.armijo;
}

/** Same contract as ccc_ddp_centroidal_solve with host pointers; n_threads host threads. */
int32_t ccc_oracle_ddp_centroidal_solve(const ccc_ddp_centroidal_batch_t * bt,
                                        const ccc_ddp_config_t * c,
                                        ccc_ddp_result_t * r,
                                        int32_t n_threads)
{
  if(!bt || !c || !r || bt->m_max > 32 || bt->m_max < 0) return CCC_ERR_INVALID;
  const int N = bt->horizon_steps, mm = bt->m_max;
  DdpConfig cfg = toConfig(c);
  parallelFor(bt->batch, n_threads, [&](int b) {
    CentroidalProblem p;
    bindCentroidal(p, bt, bt->sched_id[b]);
    DdpSolver s(p);
    s.cfg = cfg;
    std::vector<std::vector<double>> u0(N);
    for(int k = 0; k < N; k++)
    {
      int m = p.inputDim(k);
      u0[k].assign(m, 0.0);
      if(bt->u_init)
        for(int j = 0; j < m; j++) u0[k][j] = bt->u_init[(static_cast<size_t>(b) * N + k) * mm + j];
    }
    s.solve(bt->x0 + static_cast<size_t>(b) * 9, u0);
    storeResult(s, b, N, 9, mm, r);
    for(int i = 0; i < 6; i++) g_stats[i] += s.stats[i];
  });
  return CCC_OK;
}

/** Same contract as ccc_ddp_centroidal_closed_loop with host pointers: the control loop of the reference's test
 *  (tests/src/TestDdpCentroidal.cpp:94-150) with the plant of tests/src/SimModels.h:233-332, one problem per task.
 *  Plant arithmetic in the engine's operation order (sequential fma over the ridges, exact ZOH polynomial). */
int32_t ccc_oracle_ddp_centroidal_closed_loop(const ccc_ddp_centroidal_loop_t * lp,
                                              const ccc_ddp_config_t * c,
                                              ccc_ddp_centroidal_loop_result_t * r,
                                              int32_t n_threads)
{
  if(!lp || !c || !r || !r->plant || lp->m_max > 32 || lp->m_max <= 0 || lp->stride <= 0) return CCC_ERR_INVALID;
  const int N = lp->horizon_steps, mm = lp->m_max, T = lp->ticks, G = lp->grid_len, stride = lp->stride;
  if(static_cast<long long>(G) < static_cast<long long>(T) - 1 + static_cast<long long>(N) * stride + 1) return CCC_ERR_INVALID;
  parallelFor(lp->batch, n_threads, [&](int b) {
    const size_t sched = static_cast<size_t>(lp->sched_id[b]);
    const int32_t * mg = lp->m + sched * G;
    const double * rg = lp->ridge + sched * G * mm * 3;
    const double * vg = lp->vertex + sched * G * mm * 3;
    const double * fg = lp->ref_pos + sched * G * 3;
    double pos[3], vel[3], L[3];
    for(int a = 0; a < 3; a++)
    {
      pos[a] = lp->plant0[static_cast<size_t>(b) * 9 + a];
      vel[a] = lp->plant0[static_cast<size_t>(b) * 9 + 3 + a];
      L[a] = lp->plant0[static_cast<size_t>(b) * 9 + 6 + a];
    }
    double * log = r->plant + static_cast<size_t>(b) * (T + 1) * 9;
    for(int a = 0; a < 3; a++)
    {
      log[a] = pos[a];
      log[3 + a] = vel[a];
      log[6 + a] = L[a];
    }
    std::vector<int32_t> m_t(N);
    std::vector<double> ridge_t(static_cast<size_t>(N) * mm * 3), vertex_t(static_cast<size_t>(N) * mm * 3), ref_t(static_cast<size_t>(N + 1) * 3);
    std::vector<std::vector<double>> u_prev;
    DdpConfig cfg = toConfig(c);
    for(int tick = 0; tick < T; tick++)
    {
      // the horizon of this cycle: stage k <- grid entry tick + k * stride
      for(int k = 0; k <= N; k++)
      {
        const size_t e = static_cast<size_t>(tick) + static_cast<size_t>(k) * stride;
        for(int a = 0; a < 3; a++) ref_t[static_cast<size_t>(k) * 3 + a] = fg[e * 3 + a];
        if(k == N) break;
        m_t[k] = mg[e];
        std::copy(rg + e * mm * 3, rg + (e + 1) * mm * 3, ridge_t.begin() + static_cast<size_t>(k) * mm * 3);
        std::copy(vg + e * mm * 3, vg + (e + 1) * mm * 3, vertex_t.begin() + static_cast<size_t>(k) * mm * 3);
      }
      CentroidalProblem p;
      p.N = N;
      p.dt = lp->dt;
      p.mass = lp->mass;
      p.m_max = mm;
      p.m_tab = m_t.data();
      p.ridge = ridge_t.data();
      p.vertex = vertex_t.data();
      p.ref_pos = ref_t.data();
      for(int i = 0; i < 10; i++) p.w_run[i] = lp->w_run[i];
      for(int i = 0; i < 9; i++) p.w_term[i] = lp->w_term[i];
      p.u_lo = lp->u_lo;
      p.u_hi = lp->u_hi;
      DdpSolver s(p);
      if(tick == 1) cfg.max_iter = lp->max_iter_later;
      s.cfg = cfg;
      // warm start: the previous plan stage by stage, zeroed where the input dimension changed
      std::vector<std::vector<double>> u0(N);
      for(int k = 0; k < N; k++)
      {
        const int m = m_t[k];
        u0[k].assign(m, 0.0);
        if(tick > 0 && mg[static_cast<size_t>(tick) - 1 + static_cast<size_t>(k) * stride] == m) u0[k] = u_prev[k];
      }
      const double x0[9] = {pos[0], pos[1], pos[2], lp->mass * vel[0], lp->mass * vel[1], lp->mass * vel[2], L[0], L[1], L[2]};
      s.solve(x0, u0);
      u_prev = s.u;
      // plant: total wrench of the first stage about the CoM, then one exact ZOH step
      const int mk = m_t[0];
      double f[3] = {0.0, 0.0, 0.0}, n[3] = {0.0, 0.0, 0.0};
      for(int j = 0; j < mk; j++)
      {
        const double uj = s.u[0][j];
        double rho[3], d[3], cr[3];
        for(int a = 0; a < 3; a++)
        {
          rho[a] = ridge_t[static_cast<size_t>(j) * 3 + a];
          d[a] = vertex_t[static_cast<size_t>(j) * 3 + a] - pos[a];
        }
        cross3(d, rho, cr);
        for(int a = 0; a < 3; a++)
        {
          f[a] = fmad(uj, rho[a], f[a]);
          n[a] = fmad(uj, cr[a], n[a]);
        }
      }
      const double sim_dt = lp->sim_dt, half_dt2 = 0.5 * (sim_dt * sim_dt);
      for(int a = 0; a < 3; a++)
      {
        const double acc = f[a] / lp->mass + (a == 2 ? -9.80665 : 0.0);
        pos[a] = fmad(half_dt2, acc, fmad(sim_dt, vel[a], pos[a]));
        vel[a] = fmad(sim_dt, acc, vel[a]);
        L[a] = fmad(sim_dt, n[a], L[a]);
        if(tick == lp->disturb_tick) vel[a] = vel[a] + lp->disturb_vel[a];
      }
      double * lg = log + static_cast<size_t>(tick + 1) * 9;
      for(int a = 0; a < 3; a++)
      {
        lg[a] = pos[a];
        lg[3 + a] = vel[a];
        lg[6 + a] = L[a];
      }
      if(r->u0)
        for(int j = 0; j < mm; j++) r->u0[(static_cast<size_t>(b) * T + tick) * mm + j] = j < mk ? s.u[0][j] : 0.0;
      if(r->iters) r->iters[static_cast<size_t>(b) * T + tick] = s.trace.back().iter;
    }
  });
  return CCC_OK;
}

/** Evaluate the DdpCentroidal problem functions of stage k of schedule 0 at (x, u): for the
 *  derivative known-answer tests (reference tests/src/TestDdpCentroidal.cpp:176-284).
 *  Fx 9x9, Fu 9xm (row-major), Lx 9, Lu m; any output may be NULL. */
int32_t ccc_oracle_centroidal_eval(const ccc_ddp_centroidal_batch_t * bt,
                                   int32_t k,
                                   const double * x,
                                   const double * u,
                                   double * xn,
                                   double * running_cost,
                                   double * terminal_cost,
                                   double * Fx,
                                   double * Fu,
                                   double * Lx,
                                   double * Lu,
                                   double * Vx)
{
  CentroidalProblem p;
  bindCentroidal(p, bt, 0);
  const int m = p.inputDim(k);
  if(xn) p.stateEq(k, x, u, xn);
  if(running_cost) *running_cost = p.runningCost(k, x, u);
  if(terminal_cost) *terminal_cost = p.terminalCost(x);
  if(Fx || Fu)
  {
    std::vector<double> fx(81), fu(9 * m);
    p.stateEqDeriv(k, x, u, fx.data(), fu.data());
    if(Fx) std::copy(fx.begin(), fx.end(), Fx);
    if(Fu) std::copy(fu.begin(), fu.end(), Fu);
  }
  if(Lx || Lu)
  {
    std::vector<double> lx(9), lu(m), lxx(81), luu(m * m), lxu(9 * m);
    p.runningCostDeriv(k, x, u, lx.data(), lu.data(), lxx.data(), luu.data(), lxu.data());
    if(Lx) std::copy(lx.begin(), lx.end(), Lx);
    if(Lu) std::copy(lu.begin(), lu.end(), Lu);
  }
  if(Vx)
  {
    std::vector<double> vxx(81);
    p.terminalCostDeriv(x, Vx, vxx.data());
  }
  return CCC_OK;
}

/** Read and reset the aggregated solver statistics (see DdpSolver::stats). */
void ccc_oracle_stats(int64_t * out)
{
  for(int i = 0; i < 6; i++) out[i] = g_stats[i].exchange(0);
}

static void bindSrb(SrbProblem & p, const ccc_ddp_srb_batch_t * bt, int sched)
{
  const int N = bt->horizon_steps, mm = bt->m_max;
  p.N = N;
  p.dt = bt->dt;
  p.mass = bt->mass;
  p.m_max = mm;
  p.m_tab = bt->m + static_cast<size_t>(sched) * N;
  p.ridge = bt->ridge + static_cast<size_t>(sched) * N * mm * 3;
  p.vertex = bt->vertex + static_cast<size_t>(sched) * N * mm * 3;
  p.inertia = bt->inertia + static_cast<size_t>(sched) * N * 9;
  p.ref = bt->ref + static_cast<size_t>(sched) * (N + 1) * 6;
  for(int i = 0; i < 13; i++) p.w_run[i] = bt->w_run[i];
  for(int i = 0; i < 12; i++) p.w_term[i] = bt->w_term[i];
  p.u_lo = bt->u_lo;
  p.u_hi = bt->u_hi;
}

/** Same contract as ccc_ddp_srb_solve with host pointers; n_threads host threads. */
int32_t ccc_oracle_ddp_srb_solve(const ccc_ddp_srb_batch_t * bt, const ccc_ddp_config_t * c, ccc_ddp_result_t * r, int32_t n_threads)
{
  if(!bt || !c || !r || bt->m_max > 32 || bt->m_max < 0) return CCC_ERR_INVALID;
  const int N = bt->horizon_steps, mm = bt->m_max;
  DdpConfig cfg = toConfig(c);
  parallelFor(bt->batch, n_threads, [&](int b) {
    SrbProblem p;
    bindSrb(p, bt, bt->sched_id[b]);
    DdpSolver s(p);
    s.cfg = cfg;
    std::vector<std::vector<double>> u0(N);
    for(int k = 0; k < N; k++)
    {
      int m = p.inputDim(k);
      u0[k].assign(m, 0.0);
      if(bt->u_init)
        for(int j = 0; j < m; j++) u0[k][j] = bt->u_init[(static_cast<size_t>(b) * N + k) * mm + j];
    }
    s.solve(bt->x0 + static_cast<size_t>(b) * 12, u0);
    storeResult(s, b, N, 12, mm, r);
    for(int i = 0; i < 6; i++) g_stats[i] += s.stats[i];
  });
  return CCC_OK;
}

/** Problem functions of stage k of schedule 0 at (x, u): derivative known-answer tests
 *  (reference tests/src/TestDdpSingleRigidBody.cpp:197-308).  Fx 12x12, Fu 12xm row-major. */
int32_t ccc_oracle_srb_eval(const ccc_ddp_srb_batch_t * bt, int32_t k, const double * x, const double * u, double * xn,
                            double * running_cost, double * terminal_cost, double * Fx, double * Fu, double * Lx,
                            double * Lu, double * Vx)
{
  SrbProblem p;
  bindSrb(p, bt, 0);
  const int m = p.inputDim(k);
  if(xn) p.stateEq(k, x, u, xn);
  if(running_cost) *running_cost = p.runningCost(k, x, u);
  if(terminal_cost) *terminal_cost = p.terminalCost(x);
  if(Fx || Fu)
  {
    std::vector<double> fx(144), fu(12 * m);
    p.stateEqDeriv(k, x, u, fx.data(), fu.data());
    if(Fx) std::copy(fx.begin(), fx.end(), Fx);
    if(Fu) std::copy(fu.begin(), fu.end(), Fu);
  }
  if(Lx || Lu)
  {
    std::vector<double> lx(12), lu(m), lxx(144), luu(m * m), lxu(12 * m);
    p.runningCostDeriv(k, x, u, lx.data(), lu.data(), lxx.data(), luu.data(), lxu.data());
    if(Lx) std::copy(lx.begin(), lx.end(), Lx);
    if(Lu) std::copy(lu.begin(), lu.end(), Lu);
  }
  if(Vx)
  {
    std::vector<double> vxx(144);
    p.terminalCostDeriv(x, Vx, vxx.data());
  }
  return CCC_OK;
}

static void bindZmp(ZmpProblem & p, const ccc_ddp_zmp_batch_t * bt, int sched)
{
  const int N = bt->horizon_steps;
  p.N = N;
  p.dt = bt->dt;
  p.mass = bt->mass;
  p.ref_zmp = bt->ref_zmp + static_cast<size_t>(sched) * (N + 1) * 3;
  p.com_z = bt->com_z + static_cast<size_t>(sched) * (N + 1);
  p.w_run_com_z = bt->w[0];
  p.w_run_zmp = bt->w[1];
  p.w_run_fz = bt->w[2];
  p.w_term_xy = bt->w[3];
  p.w_term_z = bt->w[4];
  p.w_term_vel = bt->w[5];
}

/** Same contract as ccc_ddp_zmp_solve with host pointers; n_threads host threads. */
int32_t ccc_oracle_ddp_zmp_solve(const ccc_ddp_zmp_batch_t * bt, const ccc_ddp_config_t * c, ccc_ddp_result_t * r, int32_t n_threads)
{
  if(!bt || !c || !r) return CCC_ERR_INVALID;
  const int N = bt->horizon_steps;
  DdpConfig cfg = toConfig(c);
  parallelFor(bt->batch, n_threads, [&](int b) {
    ZmpProblem p;
    bindZmp(p, bt, bt->sched_id[b]);
    DdpSolver s(p);
    s.cfg = cfg;
    std::vector<std::vector<double>> u0(N, std::vector<double>(3, 0.0));
    if(bt->u_init)
      for(int k = 0; k < N; k++)
        for(int j = 0; j < 3; j++) u0[k][j] = bt->u_init[(static_cast<size_t>(b) * N + k) * 3 + j];
    s.solve(bt->x0 + static_cast<size_t>(b) * 6, u0);
    storeResult(s, b, N, 6, 3, r);
  });
  return CCC_OK;
}

/** Problem functions of stage k of schedule 0 at (x, u): derivative known-answer tests
 *  (reference tests/src/TestDdpZmp.cpp:153-248).  Fx 6x6, Fu 6x3 row-major. */
int32_t ccc_oracle_zmp_eval(const ccc_ddp_zmp_batch_t * bt, int32_t k, const double * x, const double * u, double * xn,
                            double * running_cost, double * terminal_cost, double * Fx, double * Fu, double * Lx, double * Lu,
                            double * Vx)
{
  ZmpProblem p;
  bindZmp(p, bt, 0);
  if(xn) p.stateEq(k, x, u, xn);
  if(running_cost) *running_cost = p.runningCost(k, x, u);
  if(terminal_cost) *terminal_cost = p.terminalCost(x);
  if(Fx || Fu)
  {
    std::vector<double> fx(36), fu(18);
    p.stateEqDeriv(k, x, u, fx.data(), fu.data());
    if(Fx) std::copy(fx.begin(), fx.end(), Fx);
    if(Fu) std::copy(fu.begin(), fu.end(), Fu);
  }
  if(Lx || Lu)
  {
    std::vector<double> lx(6), lu(3), lxx(36), luu(9), lxu(18);
    p.runningCostDeriv(k, x, u, lx.data(), lu.data(), lxx.data(), luu.data(), lxu.data());
    if(Lx) std::copy(lx.begin(), lx.end(), Lx);
    if(Lu) std::copy(lu.begin(), lu.end(), Lu);
  }
  if(Vx)
  {
    std::vector<double> vxx(36);
    p.terminalCostDeriv(x, Vx, vxx.data());
  }
  return CCC_OK;
}

/** Same contract as ccc_qp_solve with host pointers; n_threads host threads. */
int32_t ccc_oracle_qp_solve(const ccc_qp_batch_t * bt, ccc_qp_result_t * r, int32_t n_threads)
{
  if(!bt || !r || bt->n <= 0 || bt->n > 256) return CCC_ERR_INVALID;
  const int n = bt->n, me = bt->n_eq, mi = bt->n_ineq;
  DenseQpShared S;
  S.n = n;
  S.me = me;
  S.mi = mi;
  S.Q.assign(bt->Q, bt->Q + static_cast<size_t>(n) * n);
  if(me) S.A.assign(bt->A, bt->A + static_cast<size_t>(me) * n);
  S.C.assign(bt->C, bt->C + static_cast<size_t>(mi) * n);
  S.setup();
  parallelFor(bt->batch, n_threads, [&](int b) {
    DenseQpSolver solver(S);
    DenseQpResult res = solver.solve(bt->c ? bt->c + static_cast<size_t>(b) * n : nullptr,
                                     me ? bt->b + static_cast<size_t>(b) * me : nullptr, bt->d + static_cast<size_t>(b) * mi);
    if(r->x)
      for(int i = 0; i < n; i++) r->x[static_cast<size_t>(b) * n + i] = res.x[i];
    if(r->iters) r->iters[b] = res.iters;
    if(r->status) r->status[b] = res.status;
    if(r->n_active) r->n_active[b] = static_cast<int32_t>(res.active.size());
    if(r->active)
      for(int i = 0; i < n; i++)
        r->active[static_cast<size_t>(b) * n + i] = i < static_cast<int>(res.active.size()) ? res.active[i] : -1;
  });
  return CCC_OK;
}

/** Same contract as ccc_linear_mpc_xy_solve with host pointers (oracle/xy.hpp); n_threads host threads. */
int32_t ccc_oracle_linear_mpc_xy_solve(const ccc_linear_mpc_xy_batch_t * bt, ccc_linear_mpc_xy_result_t * r, int32_t n_threads)
{
  if(!bt || !r || bt->horizon_steps <= 0 || bt->n_sched <= 0 || bt->batch < 0) return CCC_ERR_INVALID;
  const int S = bt->n_sched, B = bt->batch, N = bt->horizon_steps, rows = 6 * N;
  std::vector<XySchedule> sch(S);
  std::vector<DenseQpShared> shared(S);
  parallelFor(S, n_threads, [&](int s) {
    sch[s].build(*bt, s);
    const int n = sch[s].n;
    DenseQpShared & Q = shared[s];
    Q.n = n;
    Q.me = sch[s].n_eq;
    Q.mi = 2 * n;
    Q.Q = sch[s].H;
    Q.A = sch[s].Aeq;
    Q.C.assign(static_cast<size_t>(2 * n) * n, 0.0);
    for(int j = 0; j < n; j++)
    {
      Q.C[static_cast<size_t>(j) * n + j] = -1.0;
      Q.C[static_cast<size_t>(n + j) * n + j] = 1.0;
    }
    Q.setup();
  });
  const int n = sch[0].n, me = sch[0].n_eq;
  for(int s = 1; s < S; s++)
    if(sch[s].n != n || sch[s].n_eq != me) return CCC_ERR_INVALID;
  if(n <= 0 || n > 256) return CCC_ERR_INVALID;
  for(int s = 0; s < S; s++)
  {
    if(r->A_seq) std::copy(sch[s].A_seq.begin(), sch[s].A_seq.end(), r->A_seq + static_cast<size_t>(s) * rows * 6);
    if(r->B_seq) std::copy(sch[s].B_seq.begin(), sch[s].B_seq.end(), r->B_seq + static_cast<size_t>(s) * rows * n);
    if(r->obj_mat) std::copy(sch[s].H.begin(), sch[s].H.end(), r->obj_mat + static_cast<size_t>(s) * n * n);
  }
  std::atomic<int> bad(0);
  parallelFor(B, n_threads, [&](int b) {
    const int s = bt->sched_id[b];
    if(s < 0 || s >= S)
    {
      bad = 1;
      return;
    }
    std::vector<double> g(n), d(2 * n);
    sch[s].objVec(*bt, s, bt->x0 + static_cast<size_t>(b) * 6, g.data());
    for(int j = 0; j < n; j++)
    {
      d[j] = -bt->force_lo;
      d[n + j] = bt->force_hi;
    }
    if(r->obj_vec) std::copy(g.begin(), g.end(), r->obj_vec + static_cast<size_t>(b) * n);
    DenseQpSolver solver(shared[s]);
    DenseQpResult res = solver.solve(g.data(), me ? sch[s].beq.data() : nullptr, d.data());
    if(r->u)
      for(int i = 0; i < n; i++) r->u[static_cast<size_t>(b) * n + i] = res.x[i];
    if(r->iters) r->iters[b] = res.iters;
    if(r->status) r->status[b] = res.status;
    if(r->n_active) r->n_active[b] = static_cast<int32_t>(res.active.size());
    if(r->active)
      for(int i = 0; i < n; i++)
        r->active[static_cast<size_t>(b) * n + i] = i < static_cast<int>(res.active.size()) ? res.active[i] : -1;
  });
  return bad ? CCC_ERR_INVALID : CCC_OK;
}

/** u[b] = -K x[b] + F ref_seq[b] (reference include/CCC/PreviewControl.h:86-89) in canonical order:
 *  F.ref as 32 lane-strided sequential fma chains combined by tree_sum32, K.x as one chain. */
int32_t ccc_oracle_preview_input(int32_t batch, int32_t N, const double * K, const double * F, const double * x,
                                 const double * ref_seq, double * u)
{
  for(int b = 0; b < batch; b++)
  {
    double part[32];
    for(int l = 0; l < 32; l++)
    {
      double acc = 0.0;
      for(int i = l; i < N; i += 32) acc = fmad(F[i], ref_seq[static_cast<size_t>(b) * N + i], acc);
      part[l] = acc;
    }
    double kx = 0.0;
    for(int i = 0; i < 3; i++) kx = fmad(K[i], x[static_cast<size_t>(b) * 3 + i], kx);
    u[b] = (-kx) + tree_sum32(part, 32);
  }
  return CCC_OK;
}

/** sincos_canon(x) for accuracy tests. */
void ccc_oracle_sincos(double x, double * s, double * c)
{
  sincos_canon(x, s, c);
}

/** Select the unpinned algorithmic alternatives (num.hpp `Choice` bits; 0 = the oracle's definition).  Process-wide;
 *  used by tools/srb_robustness.py and tests/test_oracle_textbook.py only. */
void ccc_oracle_set_choices(uint32_t bits)
{
  choiceBits() = bits;
}

/** 1 for libccc_oracle_textbook.so (-DORACLE_TEXTBOOK), 0 for the canonical build. */
int32_t ccc_oracle_is_textbook(void)
{
  return kTextbook ? 1 : 0;
}

int32_t ccc_oracle_hardware_threads(void)
{
  return static_cast<int32_t>(std::thread::hardware_concurrency());
}

} // extern "C"
