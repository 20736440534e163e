// ddp_zmp.cu — C-ABI entry points ccc_ddp_zmp_* (include/ccc_b200.h).
// Default: zmp_thread_kernel, one thread per problem (ddp_thread_zmp.cuh: 6 states, 3 unconstrained inputs as
// straight-line scalar code, 32 problems per warp, per-problem arrays interleaved so that a warp's accesses coalesce).
// Variant 0 (ccc_ddp_zmp_set_variant, A/B measurements): the generic warp-per-problem engine (ddp_host.cuh) with the
// CoM-ZMP model policy (model_zmp.cuh), where 3 of 32 lanes carry an input.  Both reproduce oracle/zmp.hpp bit for bit.
#include "ddp_host.cuh"
#include "ddp_thread_zmp.cuh"
#include "model_zmp.cuh"

#include <vector>

namespace
{
/** Derived per-stage tables on the device: m = 3, zero ridge/vertex, state references, constants row. */
__global__ void zmp_build_tables_kernel(const double * __restrict__ ref_zmp, const double * __restrict__ com_z, int S, int N,
                                        int * __restrict__ m, double * __restrict__ ridge, double * __restrict__ vertex,
                                        double * __restrict__ ref)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if(idx >= S * (N + 1)) return;
  const int s = idx / (N + 1), k = idx - s * (N + 1);
  double * r = ref + (size_t)idx * 6;
  const bool term = k == N;
  r[0] = term ? ref_zmp[(size_t)idx * 3] : 0.0;
  r[1] = 0.0;
  r[2] = term ? ref_zmp[(size_t)idx * 3 + 1] : 0.0;
  r[3] = 0.0;
  r[4] = com_z[idx];
  r[5] = 0.0;
  if(!term)
  {
    const size_t st = (size_t)s * N + k;
    m[st] = 3;
    for(int i = 0; i < 9; i++)
    {
      ridge[st * 9 + i] = 0.0;
      vertex[st * 9 + i] = 0.0;
    }
  }
}

__global__ void zmp_pack_consts_kernel(const double * __restrict__ ref_zmp, int S, int N, double mass, double * __restrict__ tab)
{
  const int st = blockIdx.x * blockDim.x + threadIdx.x;
  if(st >= S * N) return;
  const int s = st / N, k = st - s * N;
  const double * z = ref_zmp + ((size_t)s * (N + 1) + k) * 3;
  double * row = tab + ((size_t)st * ccc::ZmpModel::TAB_ROWS + 6) * 32;
  row[0] = z[0];
  row[1] = z[1];
  row[2] = mass * 9.80665;
  row[3] = z[2];
  for(int i = 4; i < 32; i++) row[i] = 0.0;
}
/** One thread per problem; `slab` holds the interleaved per-problem arrays (stride = padded batch). */
__global__ void __launch_bounds__(64) zmp_thread_kernel(ccc_ddp_zmp_batch_t bt, ccc_ddp_config_t cfg, ccc_ddp_result_t res, double * slab, size_t stride)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= bt.batch) return;
  ccc_thread::zmp_thread_run(bt, cfg, res, ccc_thread::zmp_work_at(slab, bt.horizon_steps, stride, (size_t)b), b);
}

int & zmp_variant()
{
  static int v = 1; // 1: thread per problem (default), 0: warp per problem
  return v;
}
} // namespace

struct ccc_ddp_zmp_ws
{
  bool thread_path = true;
  int N = 0, max_batch = 0, max_sched = 0, launches = 0;
  // thread path
  double * slab = nullptr;
  size_t stride = 0;
  cudaStream_t own_stream = nullptr;
  double *d_x0 = nullptr, *d_uinit = nullptr, *d_x = nullptr, *d_u = nullptr, *d_cost = nullptr, *d_lam = nullptr;
  int *d_sid = nullptr, *d_iters = nullptr, *d_status = nullptr;
  int8_t * d_alpha = nullptr;
  uint32_t * d_clamped = nullptr;
  int trace_cap = 0;
  // warp path
  ccc_host::DdpEngine<ccc::ZmpModel> eng;
  double *d_ref_zmp = nullptr, *d_com_z = nullptr, *d_ridge = nullptr, *d_vertex = nullptr, *d_ref = nullptr;
  int * d_m = nullptr;
};

extern "C" {

ccc_ddp_zmp_ws_t * ccc_ddp_zmp_create(int32_t horizon_steps, int32_t max_batch, int32_t max_sched)
{
  if(horizon_steps <= 0 || max_batch <= 0 || max_sched <= 0)
  {
    ccc_host::set_error("ccc_ddp_zmp_create: non-positive size");
    return nullptr;
  }
  auto * ws = new ccc_ddp_zmp_ws();
  ws->thread_path = zmp_variant() != 0;
  ws->N = horizon_steps;
  ws->max_batch = max_batch;
  ws->max_sched = max_sched;
  const size_t S = max_sched, N = horizon_steps, B = max_batch;
  bool ok = true;
  if(ws->thread_path)
  {
    int ndev = 0;
    if(!ccc_host::check(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount") || ndev == 0)
    {
      ccc_host::set_error("ccc_ddp_zmp_create: no CUDA device (this library has no CPU fallback)");
      delete ws;
      return nullptr;
    }
    ws->stride = (B + 31) & ~(size_t)31;
    ws->trace_cap = 64;
    ok = ccc_host::check(cudaStreamCreateWithFlags(&ws->own_stream, cudaStreamNonBlocking), "cudaStreamCreate");
    ok = ok && ccc_host::dev_alloc(ws->slab, ccc_thread::zmp_work_doubles(horizon_steps) * ws->stride);
    ok = ok && ccc_host::dev_alloc(ws->d_x0, B * 6) && ccc_host::dev_alloc(ws->d_uinit, B * N * 3) && ccc_host::dev_alloc(ws->d_x, B * (N + 1) * 6)
         && ccc_host::dev_alloc(ws->d_u, B * N * 3) && ccc_host::dev_alloc(ws->d_cost, B) && ccc_host::dev_alloc(ws->d_sid, B)
         && ccc_host::dev_alloc(ws->d_iters, B) && ccc_host::dev_alloc(ws->d_status, B) && ccc_host::dev_alloc(ws->d_clamped, B * N)
         && ccc_host::dev_alloc(ws->d_alpha, B * ws->trace_cap) && ccc_host::dev_alloc(ws->d_lam, B * ws->trace_cap);
  }
  else
  {
    ok = ws->eng.create(horizon_steps, max_batch, max_sched);
    ok = ok && ccc_host::dev_alloc(ws->d_ridge, S * N * 9) && ccc_host::dev_alloc(ws->d_vertex, S * N * 9);
    ok = ok && ccc_host::dev_alloc(ws->d_ref, S * (N + 1) * 6) && ccc_host::dev_alloc(ws->d_m, S * N);
  }
  ok = ok && ccc_host::dev_alloc(ws->d_ref_zmp, S * (N + 1) * 3) && ccc_host::dev_alloc(ws->d_com_z, S * (N + 1));
  if(!ok)
  {
    ccc_ddp_zmp_destroy(ws);
    return nullptr;
  }
  return ws;
}

void ccc_ddp_zmp_destroy(ccc_ddp_zmp_ws_t * ws)
{
  if(!ws) return;
  if(!ws->thread_path) ws->eng.destroy();
  void * ptrs[] = {ws->d_ref_zmp, ws->d_com_z, ws->d_ridge, ws->d_vertex, ws->d_ref, ws->d_m, ws->slab, ws->d_x0, ws->d_uinit, ws->d_x,
                   ws->d_u, ws->d_cost, ws->d_lam, ws->d_sid, ws->d_iters, ws->d_status, ws->d_alpha, ws->d_clamped};
  for(void * p : ptrs)
    if(p) cudaFree(p);
  if(ws->own_stream) cudaStreamDestroy(ws->own_stream);
  delete ws;
}

/** The thread-per-problem path of ccc_ddp_zmp_solve. */
static int32_t zmp_thread_solve(ccc_ddp_zmp_ws_t * ws, const ccc_ddp_zmp_batch_t * bt, const ccc_ddp_config_t * cfg, ccc_ddp_result_t * res,
                                int32_t mem, void * stream)
{
  using ccc_host::check;
  const int N = bt->horizon_steps, S = bt->n_sched, B = bt->batch;
  if(B <= 0 || B > ws->max_batch) return ccc_host::fail(CCC_ERR_ALLOC, "batch exceeds workspace");
  if(!bt->sched_id || !bt->x0) return ccc_host::fail(CCC_ERR_INVALID, "null input");
  if(cfg->n_alpha < 1 || cfg->n_alpha > CCC_DDP_MAX_ALPHA || cfg->max_iter < 0) return ccc_host::fail(CCC_ERR_INVALID, "bad solver configuration");
  const bool host = mem == CCC_MEM_HOST;
  cudaStream_t st = host ? ws->own_stream : reinterpret_cast<cudaStream_t>(stream);
  ccc_ddp_zmp_batch_t dbt = *bt;
  ccc_ddp_result_t dres = *res;
  const int tl = res->trace_len;
  if(host)
  {
    if(tl > ws->trace_cap && (res->alpha_idx || res->lambda_trace)) return ccc_host::fail(CCC_ERR_ALLOC, "trace_len exceeds the workspace's 64 slots");
    for(int b = 0; b < B; b++)
      if(bt->sched_id[b] < 0 || bt->sched_id[b] >= S) return ccc_host::fail(CCC_ERR_INVALID, "sched_id out of range");
#define CCC_H2D(dst, src, nbytes) \
  if(!check(cudaMemcpyAsync(dst, src, (nbytes), cudaMemcpyHostToDevice, st), "H2D")) return CCC_ERR_CUDA
    CCC_H2D(ws->d_ref_zmp, bt->ref_zmp, sizeof(double) * S * (N + 1) * 3);
    CCC_H2D(ws->d_com_z, bt->com_z, sizeof(double) * S * (N + 1));
    CCC_H2D(ws->d_sid, bt->sched_id, sizeof(int) * B);
    CCC_H2D(ws->d_x0, bt->x0, sizeof(double) * B * 6);
    if(bt->u_init) CCC_H2D(ws->d_uinit, bt->u_init, sizeof(double) * B * N * 3);
#undef CCC_H2D
    dbt.ref_zmp = ws->d_ref_zmp;
    dbt.com_z = ws->d_com_z;
    dbt.sched_id = ws->d_sid;
    dbt.x0 = ws->d_x0;
    dbt.u_init = bt->u_init ? ws->d_uinit : nullptr;
    dres.x = res->x ? ws->d_x : nullptr;
    dres.u = res->u ? ws->d_u : nullptr;
    dres.cost = res->cost ? ws->d_cost : nullptr;
    dres.iters = res->iters ? ws->d_iters : nullptr;
    dres.status = res->status ? ws->d_status : nullptr;
    dres.alpha_idx = res->alpha_idx ? ws->d_alpha : nullptr;
    dres.lambda_trace = res->lambda_trace ? ws->d_lam : nullptr;
    dres.clamped = res->clamped ? ws->d_clamped : nullptr;
  }
  zmp_thread_kernel<<<(B + 63) / 64, 64, 0, st>>>(dbt, *cfg, dres, ws->slab, ws->stride);
  ws->launches = 1;
  if(!check(cudaGetLastError(), "launch zmp_thread_kernel")) return CCC_ERR_CUDA;
  if(!host) return CCC_OK;
#define CCC_D2H(dst, src, nbytes) \
  if((dst) && !check(cudaMemcpyAsync(dst, src, (nbytes), cudaMemcpyDeviceToHost, st), "D2H")) return CCC_ERR_CUDA
  CCC_D2H(res->x, ws->d_x, sizeof(double) * B * (N + 1) * 6);
  CCC_D2H(res->u, ws->d_u, sizeof(double) * B * N * 3);
  CCC_D2H(res->cost, ws->d_cost, sizeof(double) * B);
  CCC_D2H(res->iters, ws->d_iters, sizeof(int) * B);
  CCC_D2H(res->status, ws->d_status, sizeof(int) * B);
  CCC_D2H(res->alpha_idx, ws->d_alpha, sizeof(int8_t) * B * tl);
  CCC_D2H(res->lambda_trace, ws->d_lam, sizeof(double) * B * tl);
  CCC_D2H(res->clamped, ws->d_clamped, sizeof(uint32_t) * B * N);
#undef CCC_D2H
  if(!check(cudaStreamSynchronize(st), "cudaStreamSynchronize")) return CCC_ERR_CUDA;
  return CCC_OK;
}

int32_t ccc_ddp_zmp_solve(ccc_ddp_zmp_ws_t * ws,
                          const ccc_ddp_zmp_batch_t * bt,
                          const ccc_ddp_config_t * cfg,
                          ccc_ddp_result_t * res,
                          int32_t mem,
                          void * stream)
{
  if(!ws || !bt || !cfg || !res) return ccc_host::fail(CCC_ERR_INVALID, "null argument");
  const int N = bt->horizon_steps, S = bt->n_sched;
  if(N != ws->N) return ccc_host::fail(CCC_ERR_INVALID, "horizon_steps differs from the workspace's");
  if(S <= 0 || S > ws->max_sched) return ccc_host::fail(CCC_ERR_ALLOC, "n_sched exceeds workspace");
  if(!bt->ref_zmp || !bt->com_z) return ccc_host::fail(CCC_ERR_INVALID, "null reference table");
  if(cfg->with_input_constraint) return ccc_host::fail(CCC_ERR_INVALID, "DdpZmp has no input limits: with_input_constraint must be 0");
  if(ws->thread_path) return zmp_thread_solve(ws, bt, cfg, res, mem, stream);
  ccc_host::DdpInputs<ccc::ZmpModel> in;
  in.B = bt->batch;
  in.S = S;
  in.m_max = 3;
  in.sched_id = bt->sched_id;
  in.x0 = bt->x0;
  in.u_init = bt->u_init;
  const double w_run[7] = {0, 0, 0, 0, bt->w[0], 0, 1.0};
  const double w_term[6] = {bt->w[3], bt->w[5], bt->w[3], bt->w[5], bt->w[4], bt->w[5]};
  for(int i = 0; i < 7; i++) in.w_run[i] = w_run[i];
  for(int i = 0; i < 6; i++) in.w_term[i] = w_term[i];
  in.mp.dt = bt->dt;
  in.mp.mass = bt->mass;
  in.mp.w_u[0] = bt->w[1];
  in.mp.w_u[1] = bt->w[1];
  in.mp.w_u[2] = bt->w[2];
  const bool host = mem == CCC_MEM_HOST;
  cudaStream_t own = ws->eng.own_stream;
  const double * ref_zmp_dev = bt->ref_zmp;
  // host-mode staging of the derived tables happens on the host so that the engine's own H2D path applies
  std::vector<int> h_m;
  std::vector<double> h_zero, h_ref;
  if(host)
  {
    h_m.assign((size_t)S * N, 3);
    h_zero.assign((size_t)S * N * 9, 0.0);
    h_ref.assign((size_t)S * (N + 1) * 6, 0.0);
    for(int s = 0; s < S; s++)
      for(int k = 0; k <= N; k++)
      {
        const size_t idx = (size_t)s * (N + 1) + k;
        double * r = h_ref.data() + idx * 6;
        if(k == N)
        {
          r[0] = bt->ref_zmp[idx * 3];
          r[2] = bt->ref_zmp[idx * 3 + 1];
        }
        r[4] = bt->com_z[idx];
      }
    in.m = h_m.data();
    in.ridge = h_zero.data();
    in.vertex = h_zero.data();
    in.ref = h_ref.data();
    if(!ccc_host::check(cudaMemcpyAsync(ws->d_ref_zmp, bt->ref_zmp, sizeof(double) * S * (N + 1) * 3, cudaMemcpyHostToDevice, own),
                        "H2D"))
      return CCC_ERR_CUDA;
    ref_zmp_dev = ws->d_ref_zmp;
  }
  else
  {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int total = S * (N + 1);
    zmp_build_tables_kernel<<<(total + 127) / 128, 128, 0, st>>>(bt->ref_zmp, bt->com_z, S, N, ws->d_m, ws->d_ridge, ws->d_vertex,
                                                                  ws->d_ref);
    in.m = ws->d_m;
    in.ridge = ws->d_ridge;
    in.vertex = ws->d_vertex;
    in.ref = ws->d_ref;
  }
  const double mass = bt->mass;
  const int rc = ws->eng.solve(in, cfg, res, mem, stream, [=](cudaStream_t st, double * tab) {
    zmp_pack_consts_kernel<<<(S * N + 127) / 128, 128, 0, st>>>(ref_zmp_dev, S, N, mass, tab);
    return 1;
  });
  if(!host) ws->eng.launches++;
  return rc;
}

int32_t ccc_ddp_zmp_last_launches(const ccc_ddp_zmp_ws_t * ws)
{
  return ws ? (ws->thread_path ? ws->launches : ws->eng.launches) : 0;
}

/* Tuning hook (not part of the stable ABI): 1 = one thread per problem (default), 0 = the warp-per-problem engine;
 * read by ccc_ddp_zmp_create.  Returns the previous value. */
int32_t ccc_ddp_zmp_set_variant(int32_t v)
{
  const int old = zmp_variant();
  zmp_variant() = v != 0 ? 1 : 0;
  return old;
}

} // extern "C"
