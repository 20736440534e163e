// zmp_mpc.cu — schedule compiler (footstep plans -> reference-ZMP / ZMP-limit stage tables) and the device-side
// planOnce of the two QP-based ZMP methods on those tables (include/ccc_b200.h: ccc_footstep_compile, ccc_zmp_mpc_*).
//
// Replaces (reference): FootstepManager::update / refZmp / zmpLimits (tests/src/FootstepManager.h:147-254) sampled on
// the horizon grid as LinearMpcZmp::planOnce (src/LinearMpcZmp.cpp:83-112) and IntrinsicallyStableMpc::planOnce
// (src/IntrinsicallyStableMpc.cpp:106-139) do; LinearMpcZmp1d::procOnce (src/LinearMpcZmp.cpp:46-81) and
// IntrinsicallyStableMpc1d::procOnce (src/IntrinsicallyStableMpc.cpp:63-104) around qp_solver_->solve().
//
// Arithmetic: plain IEEE * + - / in the order the reference's expressions are written (the file is compiled with
// -fmad=false), sums over an index sequentially ascending from +0.0 — what tests/test_zmp_mpc_*.py restate in numpy.
#include "../../include/ccc_b200.h"
#include "qp_host.cuh"

namespace
{
constexpr int kCompileThreads = 128;
constexpr int kMaxKnots = 4 * 64 + 4; // footsteps per plan <= 64

struct Stance
{
  double p[2][2]; // Left, Right
  int mask;       // bit f set: foot f on the ground
};

__device__ inline void stance_mid(const Stance & s, double * out)
{
  // Footstance::midPos (:88-103)
  if(s.mask == 3)
  {
    out[0] = 0.5 * (s.p[0][0] + s.p[1][0]);
    out[1] = 0.5 * (s.p[0][1] + s.p[1][1]);
  }
  else
  {
    const int f = s.mask == 1 ? 0 : 1;
    out[0] = s.p[f][0];
    out[1] = s.p[f][1];
  }
}

/** One CTA per plan: thread 0 replays FootstepManager::update (:147-209) into knot lists in shared memory (std::map::emplace
 *  keeps the first value of a key; the keys arrive in non-decreasing order), the threads sample refZmp / zmpLimits. */
__global__ void __launch_bounds__(kCompileThreads) footstep_compile_kernel(ccc_footstep_plans_t in, ccc_zmp_tables_t out, int * err)
{
  __shared__ double zt[kMaxKnots], zv[kMaxKnots][2], st[kMaxKnots], sv[kMaxKnots][4];
  __shared__ int smask[kMaxKnots], nz_s, ns_s;
  const int p = blockIdx.x, F = in.max_steps, N = in.horizon_steps;
  const double t0 = in.current_time[p];
  if(threadIdx.x == 0)
  {
    int nz = 0, ns = 0;
    auto put_z = [&](double t, const double * v) {
      if(nz > 0 && zt[nz - 1] == t) return;
      zt[nz] = t;
      zv[nz][0] = v[0];
      zv[nz][1] = v[1];
      nz++;
    };
    auto put_s = [&](double t, const Stance & s) {
      if(ns > 0 && st[ns - 1] == t) return;
      st[ns] = t;
      sv[ns][0] = s.p[0][0];
      sv[ns][1] = s.p[0][1];
      sv[ns][2] = s.p[1][0];
      sv[ns][3] = s.p[1][1];
      smask[ns] = s.mask;
      ns++;
    };
    Stance cur;
    for(int f = 0; f < 2; f++)
      for(int a = 0; a < 2; a++) cur.p[f][a] = in.stance0[(p * 2 + f) * 2 + a];
    cur.mask = 3;
    const int nsteps = in.n_steps[p] < F ? in.n_steps[p] : F;
    const int * foot = in.foot + (size_t)p * F;
    const double * pos = in.pos + (size_t)p * F * 2;
    const double * tm = in.times + (size_t)p * F * 4;
    if(nsteps > 64 && err) *err = 1;
    int first = 0;
    for(int i = 0; i < nsteps && i < 64; i++)
    {
      if(tm[i * 4 + 2] <= t0) // swing_end_time <= current_time: the foot has landed (:150-153)
      {
        cur.p[foot[i]][0] = pos[i * 2];
        cur.p[foot[i]][1] = pos[i * 2 + 1];
      }
      if(tm[i * 4 + 3] < t0) first = i + 1; // transit_end_time < current_time: removed (:156-160)
    }
    const double t_end = t0 + in.manager_horizon;
    double mid[2];
    if(first >= nsteps)
    {
      stance_mid(cur, mid);
      put_z(t0, mid);
      put_z(t_end, mid);
      put_s(t0, cur);
      put_s(t_end, cur);
    }
    else
    {
      if(t0 < tm[first * 4 + 0])
      {
        stance_mid(cur, mid);
        put_z(t0, mid);
        put_s(t0, cur);
      }
      Stance tmp = cur;
      for(int i = first; i < nsteps && i < 64 && tm[i * 4 + 0] <= t_end; i++)
      {
        const int f = foot[i], o = 1 - f;
        stance_mid(tmp, mid);
        put_z(tm[i * 4 + 0], mid);
        put_s(tm[i * 4 + 0], tmp);
        tmp.mask &= ~(1 << f);
        put_z(tm[i * 4 + 1], tmp.p[o]);
        put_s(tm[i * 4 + 1], tmp);
        tmp.mask |= 1 << f;
        tmp.p[f][0] = pos[i * 2];
        tmp.p[f][1] = pos[i * 2 + 1];
        put_z(tm[i * 4 + 2], tmp.p[o]);
        put_s(tm[i * 4 + 2], tmp);
        stance_mid(tmp, mid);
        put_z(tm[i * 4 + 3], mid);
      }
      if(zt[nz - 1] < t_end)
      {
        stance_mid(tmp, mid);
        put_z(t_end, mid);
        put_s(t_end, tmp);
      }
    }
    nz_s = nz;
    ns_s = ns;
  }
  __syncthreads();
  const int nz = nz_s, ns = ns_s;
  for(int k = threadIdx.x; k < N; k += kCompileThreads)
  {
    double t = t0 + k * in.horizon_dt;
    for(int e = 0; e < in.eps_reps; e++) t += 1e-6;
    // refZmp (:228-237): upper_bound = first key > t
    int j = 0;
    while(j < nz && !(zt[j] > t)) j++;
    double r[2] = {0.0, 0.0};
    if(j == 0 || j >= nz)
    {
      if(err) *err = 2; // sample time outside the knot list (the reference would dereference end())
    }
    else
    {
      const double ratio = (t - zt[j - 1]) / (zt[j] - zt[j - 1]);
      for(int a = 0; a < 2; a++) r[a] = (1 - ratio) * zv[j - 1][a] + ratio * zv[j][a];
    }
    // zmpLimits (:242-254): the stance of the last key <= t, widened by half the foot size
    int i = 0;
    while(i < ns && !(st[i] > t)) i++;
    double lo[2] = {0.0, 0.0}, hi[2] = {0.0, 0.0};
    if(i == 0)
    {
      if(err) *err = 2;
    }
    else
    {
      const double * v = sv[i - 1];
      const int m = smask[i - 1];
      for(int a = 0; a < 2; a++)
      {
        const double l = v[a], rr = v[2 + a];
        const double mn = m == 3 ? (l < rr ? l : rr) : (m == 1 ? l : rr);
        const double mx = m == 3 ? (l < rr ? rr : l) : (m == 1 ? l : rr);
        lo[a] = mn - 0.5 * in.foot_size[a];
        hi[a] = mx + 0.5 * in.foot_size[a];
      }
    }
    const size_t o = ((size_t)p * N + k) * 2;
    for(int a = 0; a < 2; a++)
    {
      if(out.ref_zmp) out.ref_zmp[o + a] = r[a];
      if(out.lim_min) out.lim_min[o + a] = lo[a];
      if(out.lim_max) out.lim_max[o + a] = hi[a];
    }
  }
}

/** QP vectors of QP q = axis * B + b; one CTA per QP, thread i = stage i. */
__global__ void __launch_bounds__(128) zmp_assemble_kernel(int method, int N, int B, int P, double weight_zmp, const int * __restrict__ plan_id,
                                                           const double * __restrict__ state, const double * __restrict__ A_seq,
                                                           const double * __restrict__ Pm, const double * __restrict__ ref_zmp,
                                                           const double * __restrict__ lim_min, const double * __restrict__ lim_max,
                                                           double * __restrict__ c, double * __restrict__ bq, double * __restrict__ d,
                                                           int * __restrict__ err)
{
  extern __shared__ double v[]; // method 1: z0 - z_ref over the horizon
  const int q = blockIdx.x, axis = q / B, b = q % B;
  const int pl = plan_id[b];
  if(pl < 0 || pl >= P)
  {
    if(threadIdx.x == 0) *err = 3;
    return;
  }
  const size_t trow = (size_t)pl * N;
  if(method == 0)
  {
    const double * x0 = state + ((size_t)b * 2 + axis) * 3;
    for(int i = threadIdx.x; i < N; i += blockDim.x)
    {
      // ineq_vec = [A_seq x0 - lo; -A_seq x0 + hi] (src/LinearMpcZmp.cpp:54-60)
      const double ax = (A_seq[i * 3] * x0[0] + A_seq[i * 3 + 1] * x0[1]) + A_seq[i * 3 + 2] * x0[2];
      d[(size_t)q * 2 * N + i] = ax - lim_min[(trow + i) * 2 + axis];
      d[(size_t)q * 2 * N + N + i] = (-ax) + lim_max[(trow + i) * 2 + axis];
    }
    return;
  }
  const double cp = state[((size_t)b * 2 + axis) * 2], z0 = state[((size_t)b * 2 + axis) * 2 + 1];
  for(int i = threadIdx.x; i < N; i += blockDim.x)
  {
    v[i] = z0 - ref_zmp[(trow + i) * 2 + axis];
    // ineq_vec = [-lo + z0; hi - z0] (src/IntrinsicallyStableMpc.cpp:79-80, 91-92)
    d[(size_t)q * 2 * N + i] = (-lim_min[(trow + i) * 2 + axis]) + z0;
    d[(size_t)q * 2 * N + N + i] = lim_max[(trow + i) * 2 + axis] - z0;
  }
  if(threadIdx.x == 0) bq[q] = cp - z0; // eq_vec (:72)
  __syncthreads();
  for(int j = threadIdx.x; j < N; j += blockDim.x)
  {
    // obj_vec = w_zmp P' (z0 1 - z_ref) (:88-89)
    double acc = 0.0;
    for(int i = 0; i < N; i++) acc = acc + Pm[(size_t)i * N + j] * v[i];
    c[(size_t)q * N + j] = weight_zmp * acc;
  }
}

__global__ void zmp_post_kernel(int method, int N, int B, int P, double control_dt, double c02, const int * __restrict__ plan_id,
                                const double * __restrict__ state, const double * __restrict__ lim_min, const double * __restrict__ lim_max,
                                const double * __restrict__ x, double * __restrict__ planned)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if(q >= 2 * B) return;
  const int axis = q / B, b = q % B;
  if(plan_id[b] < 0 || plan_id[b] >= P) return; // reported by the assembly kernel
  const size_t trow = (size_t)plan_id[b] * N;
  const double lo = lim_min[trow * 2 + axis], hi = lim_max[trow * 2 + axis];
  const double u0 = x[(size_t)q * N];
  double z;
  if(method == 0)
  {
    // src/LinearMpcZmp.cpp:71-79
    const double * x0 = state + ((size_t)b * 2 + axis) * 3;
    const double com_acc = x0[2] + control_dt * u0;
    const double com_pos = (x0[0] + control_dt * x0[1]) + (0.5 * (control_dt * control_dt)) * x0[2];
    z = com_pos + c02 * com_acc;
  }
  else
    z = state[((size_t)b * 2 + axis) * 2 + 1] + control_dt * u0; // src/IntrinsicallyStableMpc.cpp:100-101
  z = z < lo ? lo : z; // std::clamp
  z = hi < z ? hi : z;
  planned[(size_t)b * 2 + axis] = z;
}

template<class T>
bool dev_alloc(T *& p, size_t n)
{
  return ccc_host::check(cudaMalloc(reinterpret_cast<void **>(&p), (n ? n : 1) * sizeof(T)), "cudaMalloc");
}
} // namespace

struct ccc_zmp_mpc_ws
{
  int method = 0, N = 0, max_batch = 0, max_plans = 0, launches = 0;
  ccc_qp_ws * qp = nullptr;
  cudaStream_t own_stream = nullptr;
  double *d_Q = nullptr, *d_A = nullptr, *d_C = nullptr, *d_Aseq = nullptr, *d_P = nullptr, *d_state = nullptr;
  double *d_ref = nullptr, *d_lo = nullptr, *d_hi = nullptr, *d_planned = nullptr;
  int *d_plan = nullptr, *d_iters = nullptr, *d_status = nullptr, *err = nullptr;
  double *c = nullptr, *bq = nullptr, *d = nullptr, *x = nullptr;
};

extern "C" {

int32_t ccc_footstep_compile(const ccc_footstep_plans_t * plans, ccc_zmp_tables_t * tables, int32_t mem, void * stream_v)
{
  using ccc_host::check;
  if(!plans || !tables) return ccc_host::fail(CCC_ERR_INVALID, "null argument");
  const int P = plans->n_plans, F = plans->max_steps, N = plans->horizon_steps;
  if(P <= 0 || F < 0 || F > 64 || N <= 0 || plans->eps_reps < 0) return ccc_host::fail(CCC_ERR_INVALID, "ccc_footstep_compile: sizes (max_steps <= 64)");
  if(!plans->current_time || !plans->stance0 || !plans->n_steps || (F && (!plans->foot || !plans->pos || !plans->times)))
    return ccc_host::fail(CCC_ERR_INVALID, "null input");
  if(mem != CCC_MEM_HOST)
  {
    footstep_compile_kernel<<<P, kCompileThreads, 0, reinterpret_cast<cudaStream_t>(stream_v)>>>(*plans, *tables, nullptr);
    return check(cudaGetLastError(), "launch footstep_compile_kernel") ? CCC_OK : CCC_ERR_CUDA;
  }
  // host buffers: stage through one device allocation
  int ndev = 0;
  if(!check(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount") || ndev == 0)
    return ccc_host::fail(CCC_ERR_CUDA, "ccc_footstep_compile: no CUDA device (this library has no CPU fallback)");
  const size_t nd_in = (size_t)P * (1 + 4 + F * 2 + F * 4), ni_in = (size_t)P * (1 + F), n_out = (size_t)P * N * 2;
  double * dd = nullptr;
  int * di = nullptr;
  if(!dev_alloc(dd, nd_in + 3 * n_out) || !dev_alloc(di, ni_in + 1))
  {
    cudaFree(dd);
    return CCC_ERR_CUDA;
  }
  ccc_footstep_plans_t dp = *plans;
  double * w = dd;
  auto up_d = [&](const double * src, size_t n) {
    double * dst = w;
    if(n) cudaMemcpy(dst, src, n * sizeof(double), cudaMemcpyHostToDevice);
    w += n;
    return dst;
  };
  dp.current_time = up_d(plans->current_time, P);
  dp.stance0 = up_d(plans->stance0, (size_t)P * 4);
  dp.pos = up_d(plans->pos, (size_t)P * F * 2);
  dp.times = up_d(plans->times, (size_t)P * F * 4);
  cudaMemcpy(di, plans->n_steps, P * sizeof(int), cudaMemcpyHostToDevice);
  if(F) cudaMemcpy(di + P, plans->foot, (size_t)P * F * sizeof(int), cudaMemcpyHostToDevice);
  dp.n_steps = di;
  dp.foot = di + P;
  int * derr = di + ni_in;
  cudaMemset(derr, 0, sizeof(int));
  ccc_zmp_tables_t dt;
  dt.ref_zmp = w;
  dt.lim_min = w + n_out;
  dt.lim_max = w + 2 * n_out;
  footstep_compile_kernel<<<P, kCompileThreads>>>(dp, dt, derr);
  bool ok = check(cudaGetLastError(), "launch footstep_compile_kernel");
  int err = 0;
  ok = ok && check(cudaMemcpy(&err, derr, sizeof(int), cudaMemcpyDeviceToHost), "D2H");
  if(ok && tables->ref_zmp) ok = check(cudaMemcpy(tables->ref_zmp, dt.ref_zmp, n_out * sizeof(double), cudaMemcpyDeviceToHost), "D2H");
  if(ok && tables->lim_min) ok = check(cudaMemcpy(tables->lim_min, dt.lim_min, n_out * sizeof(double), cudaMemcpyDeviceToHost), "D2H");
  if(ok && tables->lim_max) ok = check(cudaMemcpy(tables->lim_max, dt.lim_max, n_out * sizeof(double), cudaMemcpyDeviceToHost), "D2H");
  cudaFree(dd);
  cudaFree(di);
  if(!ok) return CCC_ERR_CUDA;
  if(err == 1) return ccc_host::fail(CCC_ERR_INVALID, "more than 64 footsteps in a plan");
  if(err == 2) return ccc_host::fail(CCC_ERR_INVALID, "a sample time lies outside the plan's knot list (horizon longer than manager_horizon?)");
  return CCC_OK;
}

ccc_zmp_mpc_ws_t * ccc_zmp_mpc_create(int32_t method, int32_t horizon_steps, int32_t max_batch, int32_t max_plans)
{
  using ccc_host::check;
  if(method < 0 || method > 1 || horizon_steps <= 0 || horizon_steps > 256 || max_batch <= 0 || max_plans <= 0)
  {
    ccc_host::set_error("ccc_zmp_mpc_create: method 0/1, horizon_steps <= 256");
    return nullptr;
  }
  auto * ws = new ccc_zmp_mpc_ws();
  ws->method = method;
  ws->N = horizon_steps;
  ws->max_batch = max_batch;
  ws->max_plans = max_plans;
  const int me = method == 1 ? 1 : 0;
  ws->qp = ccc_host::qp_ws_create(horizon_steps, me, 2 * horizon_steps, 2 * max_batch, 1, false);
  if(!ws->qp)
  {
    delete ws;
    return nullptr;
  }
  const size_t N = horizon_steps, B = max_batch, P = max_plans;
  bool ok = check(cudaStreamCreateWithFlags(&ws->own_stream, cudaStreamNonBlocking), "cudaStreamCreate");
  ok = ok && dev_alloc(ws->d_Q, N * N) && dev_alloc(ws->d_A, N) && dev_alloc(ws->d_C, 2 * N * N) && dev_alloc(ws->d_Aseq, N * 3)
       && dev_alloc(ws->d_P, N * N) && dev_alloc(ws->d_state, B * 6) && dev_alloc(ws->d_ref, P * N * 2) && dev_alloc(ws->d_lo, P * N * 2)
       && dev_alloc(ws->d_hi, P * N * 2) && dev_alloc(ws->d_planned, B * 2) && dev_alloc(ws->d_plan, B) && dev_alloc(ws->d_iters, 2 * B)
       && dev_alloc(ws->d_status, 2 * B) && dev_alloc(ws->err, 1);
  ok = ok && dev_alloc(ws->c, 2 * B * N) && dev_alloc(ws->bq, 2 * B) && dev_alloc(ws->d, 2 * B * 2 * N) && dev_alloc(ws->x, 2 * B * N);
  if(!ok)
  {
    ccc_zmp_mpc_destroy(ws);
    return nullptr;
  }
  return ws;
}

void ccc_zmp_mpc_destroy(ccc_zmp_mpc_ws_t * ws)
{
  if(!ws) return;
  void * ptrs[] = {ws->d_Q, ws->d_A, ws->d_C, ws->d_Aseq, ws->d_P, ws->d_state, ws->d_ref, ws->d_lo, ws->d_hi, ws->d_planned,
                   ws->d_plan, ws->d_iters, ws->d_status, ws->err, ws->c, ws->bq, ws->d, ws->x};
  for(void * p : ptrs)
    if(p) cudaFree(p);
  if(ws->own_stream) cudaStreamDestroy(ws->own_stream);
  ccc_qp_destroy(ws->qp);
  delete ws;
}

int32_t ccc_zmp_mpc_plan(ccc_zmp_mpc_ws_t * ws, const ccc_zmp_mpc_batch_t * bt, ccc_zmp_mpc_result_t * res, int32_t mem, void * stream_v)
{
  using ccc_host::check;
  if(!ws || !bt || !res) return ccc_host::fail(CCC_ERR_INVALID, "null argument");
  const int N = bt->horizon_steps, B = bt->batch, P = bt->n_plans, method = bt->method;
  if(method != ws->method || N != ws->N) return ccc_host::fail(CCC_ERR_INVALID, "method / horizon_steps differ from the workspace's");
  if(B <= 0 || B > ws->max_batch || P <= 0 || P > ws->max_plans) return ccc_host::fail(CCC_ERR_ALLOC, "batch / n_plans exceeds workspace");
  if(!(bt->control_dt > 0)) return ccc_host::fail(CCC_ERR_INVALID, "control_dt must be resolved (> 0)");
  const bool reuse = bt->Q == nullptr; // keep the QP matrices (and their factorisation) of the previous call
  if(reuse && !ws->qp->have_setup) return ccc_host::fail(CCC_ERR_INVALID, "Q is NULL but this workspace holds no matrices yet");
  if(!bt->plan_id || !bt->state || !res->planned_zmp || !bt->tables.lim_min || !bt->tables.lim_max
     || (!reuse && (!bt->C || (method == 1 && (!bt->A || !bt->P)) || (method == 0 && !bt->A_seq))) || (method == 1 && !bt->tables.ref_zmp))
    return ccc_host::fail(CCC_ERR_INVALID, "null input");
  const bool host = mem == CCC_MEM_HOST;
  cudaStream_t st = host ? ws->own_stream : reinterpret_cast<cudaStream_t>(stream_v);
  ws->launches = 0;
  ws->qp->launches = 0;
  const int sdim = method == 0 ? 6 : 4;
  const double *Q = bt->Q, *A = bt->A, *C = bt->C, *Aseq = bt->A_seq, *Pm = bt->P, *state = bt->state;
  const double *ref = bt->tables.ref_zmp, *lo = bt->tables.lim_min, *hi = bt->tables.lim_max;
  const int * plan = bt->plan_id;
#define CCC_H2D(dst, src, nbytes) \
  if(!check(cudaMemcpyAsync(dst, src, (nbytes), cudaMemcpyHostToDevice, st), "H2D")) return CCC_ERR_CUDA
  if(host)
  {
    const size_t NN = (size_t)N * N, T = (size_t)P * N * 2;
    if(!reuse)
    {
      CCC_H2D(ws->d_Q, Q, sizeof(double) * NN);
      CCC_H2D(ws->d_C, C, sizeof(double) * 2 * NN);
      if(method == 1)
      {
        CCC_H2D(ws->d_A, A, sizeof(double) * N);
        CCC_H2D(ws->d_P, Pm, sizeof(double) * NN);
      }
      else
        CCC_H2D(ws->d_Aseq, Aseq, sizeof(double) * N * 3);
    }
    CCC_H2D(ws->d_state, state, sizeof(double) * B * sdim);
    CCC_H2D(ws->d_plan, plan, sizeof(int) * B);
    if(ref) CCC_H2D(ws->d_ref, ref, sizeof(double) * T);
    CCC_H2D(ws->d_lo, lo, sizeof(double) * T);
    CCC_H2D(ws->d_hi, hi, sizeof(double) * T);
    Q = ws->d_Q;
    A = ws->d_A;
    C = ws->d_C;
    Aseq = ws->d_Aseq;
    Pm = ws->d_P;
    state = ws->d_state;
    plan = ws->d_plan;
    ref = ws->d_ref;
    lo = ws->d_lo;
    hi = ws->d_hi;
  }
  else if(reuse)
  {
    Aseq = ws->d_Aseq;
    Pm = ws->d_P;
  }
  else
  {
    // device-pointer call: keep copies of the assembly matrices for later calls with Q == NULL
    if(method == 0 && !check(cudaMemcpyAsync(ws->d_Aseq, Aseq, sizeof(double) * N * 3, cudaMemcpyDeviceToDevice, st), "D2D")) return CCC_ERR_CUDA;
    if(method == 1 && !check(cudaMemcpyAsync(ws->d_P, Pm, sizeof(double) * N * N, cudaMemcpyDeviceToDevice, st), "D2D")) return CCC_ERR_CUDA;
  }
#undef CCC_H2D
  if(!check(cudaMemsetAsync(ws->err, 0, sizeof(int), st), "memset")) return CCC_ERR_CUDA;
  if(!reuse)
  {
    const int rc = ccc_host::qp_setup_launch(ws->qp, 1, Q, method == 1 ? A : nullptr, C, st);
    if(rc != CCC_OK) return rc;
  }
  zmp_assemble_kernel<<<2 * B, 128, sizeof(double) * N, st>>>(method, N, B, P, bt->weight_zmp, plan, state, Aseq, Pm, ref, lo, hi, ws->c, ws->bq, ws->d,
                                                             ws->err);
  ws->launches++;
  ccc::QpParams Pq = ccc_host::qp_params(ws->qp, 2 * B, nullptr);
  Pq.c = method == 1 ? ws->c : nullptr;
  Pq.b = ws->bq;
  Pq.d = ws->d;
  Pq.out_x = ws->x;
  Pq.out_iters = host ? (res->iters ? ws->d_iters : nullptr) : res->iters;
  Pq.out_status = host ? (res->status ? ws->d_status : nullptr) : res->status;
  int rc = ccc_host::qp_launch(ws->qp, Pq, st);
  if(rc != CCC_OK) return rc;
  double * planned = host ? ws->d_planned : res->planned_zmp;
  zmp_post_kernel<<<(2 * B + 255) / 256, 256, 0, st>>>(method, N, B, P, bt->control_dt, -bt->com_height_over_g, plan, state, lo, hi, ws->x, planned);
  ws->launches += 1 + ws->qp->launches;
  if(!check(cudaGetLastError(), "launch zmp kernels")) return CCC_ERR_CUDA;
  if(!host) return CCC_OK;
  if(!check(cudaMemcpyAsync(res->planned_zmp, planned, sizeof(double) * B * 2, cudaMemcpyDeviceToHost, st), "D2H")) return CCC_ERR_CUDA;
  if(res->iters && !check(cudaMemcpyAsync(res->iters, ws->d_iters, sizeof(int) * 2 * B, cudaMemcpyDeviceToHost, st), "D2H")) return CCC_ERR_CUDA;
  if(res->status && !check(cudaMemcpyAsync(res->status, ws->d_status, sizeof(int) * 2 * B, cudaMemcpyDeviceToHost, st), "D2H")) return CCC_ERR_CUDA;
  int err = 0;
  if(!check(cudaMemcpyAsync(&err, ws->err, sizeof(int), cudaMemcpyDeviceToHost, st), "D2H")) return CCC_ERR_CUDA;
  if(!check(cudaStreamSynchronize(st), "cudaStreamSynchronize")) return CCC_ERR_CUDA;
  if(err == 3) return ccc_host::fail(CCC_ERR_INVALID, "plan_id out of range");
  return CCC_OK;
}

int32_t ccc_zmp_mpc_last_launches(const ccc_zmp_mpc_ws_t * ws)
{
  return ws ? ws->launches : 0;
}

} // extern "C"
