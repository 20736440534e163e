/* tests/cpp/oracle_engine_shim.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * Implements the entry points of include/ccc_b200.h on top of the CPU oracle (oracle/libccc_oracle.so) so
 * that the host-side logic of the C++ drop-in classes (callback sampling, coefficient assembly, batching,
 * post-processing) can be exercised by the CPU test tier with the very same test programs that run against
 * libccc_b200.so on the GPU.  It is linked INSTEAD of libccc_b200.so by tests/test_cpp_dropin.py only; the
 * product library has no CPU path and nothing in the package refers to this file.
 */
#include "../../centroidalcontrolcollection_b200/csrc/step_mpc_core.cuh"
#include <algorithm>
#include <cmath>
#include <vector>

#include "../../include/ccc_b200.h"

extern "C" {
void ccc_oracle_ddp_config_default(ccc_ddp_config_t *);
int32_t ccc_oracle_ddp_centroidal_solve(const ccc_ddp_centroidal_batch_t *, const ccc_ddp_config_t *, ccc_ddp_result_t *, int32_t);
int32_t ccc_oracle_ddp_srb_solve(const ccc_ddp_srb_batch_t *, const ccc_ddp_config_t *, ccc_ddp_result_t *, int32_t);
int32_t ccc_oracle_ddp_zmp_solve(const ccc_ddp_zmp_batch_t *, const ccc_ddp_config_t *, ccc_ddp_result_t *, int32_t);
int32_t ccc_oracle_qp_solve(const ccc_qp_batch_t *, ccc_qp_result_t *, int32_t);
int32_t ccc_oracle_linear_mpc_xy_solve(const ccc_linear_mpc_xy_batch_t *, ccc_linear_mpc_xy_result_t *, int32_t);
int32_t ccc_oracle_ddp_centroidal_closed_loop(const ccc_ddp_centroidal_loop_t *, const ccc_ddp_config_t *, ccc_ddp_centroidal_loop_result_t *, int32_t);
int32_t ccc_oracle_preview_input(int32_t, int32_t, const double *, const double *, const double *, const double *, double *);
int32_t ccc_oracle_hardware_threads(void);

struct ccc_ddp_centroidal_ws { int dummy; };
struct ccc_ddp_srb_ws { int dummy; };
struct ccc_ddp_zmp_ws { int dummy; };
struct ccc_qp_ws
{
  std::vector<double> Q, A, C; // matrices of the last call that carried them (Q == NULL reuses them)
};

void ccc_ddp_config_default(ccc_ddp_config_t * c) { ccc_oracle_ddp_config_default(c); }
ccc_ddp_centroidal_ws_t * ccc_ddp_centroidal_create(int32_t, int32_t, int32_t) { return new ccc_ddp_centroidal_ws; }
void ccc_ddp_centroidal_destroy(ccc_ddp_centroidal_ws_t * w) { delete w; }
int32_t ccc_ddp_centroidal_solve(ccc_ddp_centroidal_ws_t *, const ccc_ddp_centroidal_batch_t * b, const ccc_ddp_config_t * c, ccc_ddp_result_t * r, int32_t, void *)
{
  return ccc_oracle_ddp_centroidal_solve(b, c, r, ccc_oracle_hardware_threads());
}
int32_t ccc_ddp_centroidal_closed_loop(ccc_ddp_centroidal_ws_t *, const ccc_ddp_centroidal_loop_t * l, const ccc_ddp_config_t * c, ccc_ddp_centroidal_loop_result_t * r, int32_t, void *)
{
  return ccc_oracle_ddp_centroidal_closed_loop(l, c, r, ccc_oracle_hardware_threads());
}
ccc_ddp_srb_ws_t * ccc_ddp_srb_create(int32_t, int32_t, int32_t) { return new ccc_ddp_srb_ws; }
void ccc_ddp_srb_destroy(ccc_ddp_srb_ws_t * w) { delete w; }
int32_t ccc_ddp_srb_solve(ccc_ddp_srb_ws_t *, const ccc_ddp_srb_batch_t * b, const ccc_ddp_config_t * c, ccc_ddp_result_t * r, int32_t, void *)
{
  return ccc_oracle_ddp_srb_solve(b, c, r, ccc_oracle_hardware_threads());
}
ccc_ddp_zmp_ws_t * ccc_ddp_zmp_create(int32_t, int32_t, int32_t) { return new ccc_ddp_zmp_ws; }
void ccc_ddp_zmp_destroy(ccc_ddp_zmp_ws_t * w) { delete w; }
int32_t ccc_ddp_zmp_solve(ccc_ddp_zmp_ws_t *, const ccc_ddp_zmp_batch_t * b, const ccc_ddp_config_t * c, ccc_ddp_result_t * r, int32_t, void *)
{
  return ccc_oracle_ddp_zmp_solve(b, c, r, ccc_oracle_hardware_threads());
}
struct ccc_linear_mpc_xy_ws { int dummy; };
ccc_linear_mpc_xy_ws_t * ccc_linear_mpc_xy_create(int32_t, int32_t, int32_t, int32_t, int32_t) { return new ccc_linear_mpc_xy_ws; }
void ccc_linear_mpc_xy_destroy(ccc_linear_mpc_xy_ws_t * w) { delete w; }
int32_t ccc_linear_mpc_xy_solve(ccc_linear_mpc_xy_ws_t *, const ccc_linear_mpc_xy_batch_t * b, ccc_linear_mpc_xy_result_t * r, int32_t, void *)
{
  return ccc_oracle_linear_mpc_xy_solve(b, r, ccc_oracle_hardware_threads());
}
ccc_qp_ws_t * ccc_qp_create(int32_t, int32_t, int32_t, int32_t) { return new ccc_qp_ws; }
void ccc_qp_destroy(ccc_qp_ws_t * w) { delete w; }
int32_t ccc_qp_solve(ccc_qp_ws_t * w, const ccc_qp_batch_t * b, ccc_qp_result_t * r, int32_t, void *)
{
  ccc_qp_batch_t bt = *b;
  const size_t n = b->n;
  if(b->Q)
  {
    w->Q.assign(b->Q, b->Q + n * n);
    w->A.assign(b->A ? b->A : b->Q, b->A ? b->A + static_cast<size_t>(b->n_eq) * n : b->Q);
    w->C.assign(b->C, b->C + static_cast<size_t>(b->n_ineq) * n);
  }
  else if(w->Q.empty())
    return CCC_ERR_INVALID;
  bt.Q = w->Q.data();
  bt.A = b->n_eq ? w->A.data() : nullptr;
  bt.C = w->C.data();
  return ccc_oracle_qp_solve(&bt, r, ccc_oracle_hardware_threads());
}
ccc_qp_ws_t * ccc_qp_create_grouped(int32_t, int32_t, int32_t, int32_t, int32_t) { return new ccc_qp_ws; }
/** Matrix groups on the one-group oracle: group by group, results scattered back. */
int32_t ccc_qp_solve_grouped(ccc_qp_ws_t * w, const ccc_qp_batch_t * b, int32_t n_groups, const int32_t * gid, ccc_qp_result_t * r, int32_t mem, void * st)
{
  if(n_groups == 1 || !gid) return ccc_qp_solve(w, b, r, mem, st);
  const size_t n = b->n, me = b->n_eq, mi = b->n_ineq;
  for(int g = 0; g < n_groups; g++)
  {
    std::vector<int> idx;
    for(int k = 0; k < b->batch; k++)
      if(gid[k] == g) idx.push_back(k);
    if(idx.empty()) continue;
    const size_t nb = idx.size();
    std::vector<double> c(nb * n, 0.0), bb(nb * me), d(nb * mi), x(nb * n);
    std::vector<int32_t> it(nb), stt(nb), na(nb), act(nb * n);
    for(size_t k = 0; k < nb; k++)
    {
      if(b->c) std::copy(b->c + idx[k] * n, b->c + (idx[k] + 1) * n, c.begin() + k * n);
      if(me) std::copy(b->b + idx[k] * me, b->b + (idx[k] + 1) * me, bb.begin() + k * me);
      std::copy(b->d + idx[k] * mi, b->d + (idx[k] + 1) * mi, d.begin() + k * mi);
    }
    ccc_qp_batch_t bt = *b;
    bt.batch = static_cast<int32_t>(nb);
    bt.Q = b->Q + g * n * n;
    bt.A = me ? b->A + g * me * n : nullptr;
    bt.c = b->c ? c.data() : nullptr;
    bt.b = me ? bb.data() : nullptr;
    bt.d = d.data();
    ccc_qp_result_t rr{};
    rr.x = x.data();
    rr.iters = it.data();
    rr.status = stt.data();
    rr.n_active = na.data();
    rr.active = act.data();
    const int rc = ccc_oracle_qp_solve(&bt, &rr, 1);
    if(rc != CCC_OK) return rc;
    for(size_t k = 0; k < nb; k++)
    {
      if(r->x) std::copy(x.begin() + k * n, x.begin() + (k + 1) * n, r->x + idx[k] * n);
      if(r->iters) r->iters[idx[k]] = it[k];
      if(r->status) r->status[idx[k]] = stt[k];
      if(r->n_active) r->n_active[idx[k]] = na[k];
      if(r->active) std::copy(act.begin() + k * n, act.begin() + (k + 1) * n, r->active + idx[k] * n);
    }
  }
  return CCC_OK;
}
int32_t ccc_preview_input(int32_t batch, int32_t n, const double * K, const double * F, const double * x, const double * ref, double * u, int32_t, void *)
{
  return ccc_oracle_preview_input(batch, n, K, F, x, ref, u);
}
/* The closed-form controllers have no solver behind them; the shim evaluates the reference's expressions
 * (src/DcmTracking.cpp:7-48, src/FootGuidedControl.cpp:11-69) with libm so that the C++ drop-ins' host logic runs on the CPU. */
int32_t ccc_dcm_tracking_plan(const ccc_dcm_tracking_batch_t * bt, double * out, int32_t, void *)
{
  const int K = bt->max_knots;
  for(int b = 0; b < bt->batch; b++)
  {
    const int p = bt->plan_id[b];
    if(p < 0 || p >= bt->n_plans) return CCC_ERR_INVALID;
    const int nk = bt->n_knots[p];
    const double t = bt->current_time[p];
    const double * kt = bt->knot_time + static_cast<size_t>(p) * K;
    const double * kz = bt->knot_zmp + static_cast<size_t>(p) * K * 2;
    for(int i = 0; i < nk; i++)
      if(kt[i] < t) return CCC_ERR_INVALID;
    for(int a = 0; a < 2; a++)
    {
      const double cz = bt->current_zmp[2 * p + a];
      double target = cz;
      if(nk > 0)
      {
        double dcm_switch = kz[(nk - 1) * 2 + a];
        for(int i = nk - 2; i >= 0; i--) dcm_switch = kz[i * 2 + a] + std::exp(-1 * bt->omega * (kt[i + 1] - kt[i])) * (dcm_switch - kz[i * 2 + a]);
        target = cz + std::exp(bt->omega * (t - kt[0])) * (dcm_switch - cz);
      }
      out[2 * b + a] = cz + (1.0 + bt->feedback_gain / bt->omega) * (bt->dcm[2 * b + a] - target);
    }
  }
  return CCC_OK;
}
int32_t ccc_foot_guided_plan(const ccc_foot_guided_batch_t * bt, double * out, int32_t, void *)
{
  for(int b = 0; b < bt->batch; b++)
  {
    const int p = bt->plan_id[b];
    if(p < 0 || p >= bt->n_plans) return CCC_ERR_INVALID;
    const double w = bt->omega, t = bt->current_time[p], ts = bt->transit_start_time[p], td = bt->transit_duration[p], te = ts + td;
    if(!(td >= 0) || !(te >= t + 1e-6)) return CCC_ERR_INVALID;
    for(int a = 0; a < 2; a++)
    {
      const double cp = bt->capture_point[2 * b + a], zs = bt->transit_start_zmp[2 * p + a], ze = bt->transit_end_zmp[2 * p + a];
      double z;
      if(td == 0)
        z = zs + 2 * ((cp - zs) - (ze - zs) * std::exp(-1 * w * (ts - t))) / (1.0 - std::exp(-2 * w * (ts - t)));
      else
      {
        const double vel = (ze - zs) / td;
        if(t <= ts)
          z = zs + (2 * (cp - zs) + 2 * vel / w * (std::exp(-1 * w * (te - t)) - std::exp(-1 * w * (ts - t)))) / (1.0 - std::exp(-2 * w * (te - t)));
        else
        {
          const double cur = zs + vel * (t - ts);
          z = cur + (2 * (cp - cur) + 2 * vel / w * (std::exp(-1 * w * (te - t)) - 1.0)) / (1.0 - std::exp(-2 * w * (te - t)));
        }
      }
      out[2 * b + a] = z;
    }
  }
  return CCC_OK;
}
/* src/SingularPreviewControlZmp.cpp:23-57 per problem and axis, the reference's expressions with std::pow as written there. */
int32_t ccc_singular_preview_plan(const ccc_singular_preview_batch_t * bt, double * out, int32_t, void *)
{
  const int N = bt->horizon_steps;
  const double w = bt->omega, dt = bt->horizon_dt;
  for(int b = 0; b < bt->batch; b++)
  {
    const int p = bt->plan_id[b];
    if(p < 0 || p >= bt->n_plans) return CCC_ERR_INVALID;
    for(int a = 0; a < 2; a++)
    {
      const double * x = bt->state + (static_cast<size_t>(b) * 2 + a) * 3;
      const double * seq = bt->ref_zmp + static_cast<size_t>(p) * N * 2 + a;
      const double K[3] = {(1 + w * dt) / dt + w / (1 + w * dt), -1 * (2 + w * dt) / dt, -1 * (2 + w * dt) / (w * dt)};
      const double u_fb = -1 * (K[0] * x[0] + K[1] * x[1] + K[2] * x[2]);
      double S0 = seq[static_cast<size_t>(N - 1) * 2] * (1 + w * dt) / (w * dt);
      for(int i = N - 2; i >= 1; i--) S0 = seq[static_cast<size_t>(i) * 2] + S0 / (1 + w * dt);
      const double u_ff = seq[0] / dt - w * (2 + w * dt) * S0 / std::pow(1 + w * dt, 2);
      out[2 * b + a] = x[0] + bt->control_dt * (u_fb + u_ff);
    }
  }
  return CCC_OK;
}
/* StepMpc: the kernel's scalar core (csrc/step_mpc_core.cuh) compiled for the host, a problem and axis at a time. */
int32_t ccc_step_mpc_plan(const ccc_step_mpc_batch_t * bt, const ccc_step_mpc_result_t * res, int32_t, void *)
{
  const int K = bt->max_elements;
  const ccc_step::Weights w = {bt->w_free_zmp, bt->w_fixed_zmp, bt->w_double_support, bt->w_pos, bt->w_vel, bt->w_capture_point_abs,
                               bt->w_capture_point_rel};
  for(int b = 0; b < bt->batch; b++)
  {
    const int p = bt->plan_id[b];
    if(p < 0 || p >= bt->n_plans || bt->n_elements[p] < 1 || bt->n_elements[p] > K) return CCC_ERR_INVALID;
    for(int a = 0; a < 2; a++)
    {
      double cur = 0, nxt = 0;
      bool has = false;
      if(!ccc_step::plan_1d(bt->n_elements[p], bt->single + static_cast<size_t>(p) * K, bt->zmp + static_cast<size_t>(p) * K * 2 + a, 2,
                            bt->end_time + static_cast<size_t>(p) * K, bt->current_time[p], bt->com_height, w, bt->x_pos[2 * b + a],
                            bt->x_vel[2 * b + a], cur, nxt, has))
        return CCC_ERR_INVALID;
      res->current_zmp[2 * b + a] = cur;
      if(res->next_foot_zmp) res->next_foot_zmp[2 * b + a] = nxt;
      if(res->has_next) res->has_next[b] = has ? 1 : 0;
    }
  }
  return CCC_OK;
}
const char * ccc_last_error(void) { return "oracle shim"; }
int32_t ccc_device_count(void) { return 0; }
int32_t ccc_abi_version(void) { return CCC_B200_ABI_VERSION; }
}
