// ddp_centroidal.cu — kernels + C-ABI host code of ccc_ddp_centroidal_* (include/ccc_b200.h).
//
// One persistent kernel: every warp repeatedly claims the next unsolved problem from an
// atomic counter (DDP iteration counts vary by 10x between problems, so static assignment
// would leave most of a CTA idle) and runs ccc::CentroidalWarp::solve() on it.
// Compile with -fmad=false: only the explicit fma() calls of the cores may fuse.
#include "../../include/ccc_b200.h"
#include "common_host.cuh"
#include "ddp_centroidal_core.cuh"

namespace
{
constexpr int kResumeFlag = 1 << 30;

/** Work queue shared by all warps of the persistent kernel (device memory).
 *  slot[i] >= 0: problem id (| kResumeFlag if it is a suspended solve); -1: not published yet.
 *  Problems 0..B-1 are published up front; a warp that suspends a solve appends it at `tail`, so
 *  suspended solves come round again after everything that was queued before them. */
struct SolveQueue
{
  int * slot;
  int * head; // next ticket to hand out
  int * tail; // next free slot
  int * done; // finished problems
  int capacity;
};

/** Launch shape variants (occupancy vs register budget; picked at run time, see solve()). */
template<int WARPS, int CTAS, bool CONSTRAINED>
__global__ void __launch_bounds__(WARPS * 32, CTAS)
    ddp_centroidal_solve_kernel(const __grid_constant__ ccc::CentroidalParams P, const SolveQueue q)
{
  extern __shared__ __align__(16) double smem[];
  double * s = smem + (threadIdx.x >> 5) * ccc::sm::TOTAL;
  const int lane = threadIdx.x & 31;
  for(;;)
  {
    int e = -1;
    if(lane == 0)
    {
      const int ticket = atomicAdd(q.head, 1);
      if(ticket < q.capacity)
      {
        volatile int * vs = q.slot + ticket;
        volatile int * vd = q.done;
        for(;;)
        {
          e = *vs;
          if(e >= 0) break;
          if(*vd >= P.B) break; // everything is finished: nothing will be published any more
          __nanosleep(256);
        }
      }
    }
    e = __shfl_sync(0xffffffffu, e, 0);
    if(e < 0) break;
    // acquire: the previous visit of this problem may have run on another SM
    __threadfence();
    const int b = e & (kResumeFlag - 1);
    ccc::CentroidalWarp<CONSTRAINED> w(P, s, b);
    const bool finished = w.solve((e & kResumeFlag) != 0);
    __syncwarp();
    if(lane == 0)
    {
      if(finished)
      {
        atomicAdd(q.done, 1);
      }
      else
      {
        __threadfence(); // release: trajectories, gains and resume state before the slot
        const int t = atomicAdd(q.tail, 1);
        if(t < q.capacity) *(volatile int *)(q.slot + t) = b | kResumeFlag;
      }
    }
  }
}

__global__ void init_queue_kernel(SolveQueue q, int B)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < q.capacity) q.slot[i] = i < B ? i : -1;
  if(i == 0)
  {
    *q.head = 0;
    *q.tail = B;
    *q.done = 0;
  }
}

struct Variant
{
  int warps, ctas;
  void (*kernel[2])(const ccc::CentroidalParams, const SolveQueue); // [unconstrained, constrained]
};
#define CCC_VARIANT(W, C) {W, C, {ddp_centroidal_solve_kernel<W, C, false>, ddp_centroidal_solve_kernel<W, C, true>}}
const Variant kVariants[] = {
    CCC_VARIANT(8, 2), // 16 warps/SM, 128 registers
    CCC_VARIANT(4, 3), // 12 warps/SM, 168 registers
    CCC_VARIANT(8, 1), //  8 warps/SM, 255 registers
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
int g_variant = 2; // measured best on B200 (profiles/r01_summary.md): 8 warps/SM, no spills, least I-cache pressure
int g_chunk = 32; // DDP iterations per visit before a solve is suspended and re-queued

/** ridge/vertex [S][N][m_max][3]  ->  tab [S][N][6][32] (component-major, lane-contiguous, zero padded). */
__global__ void pack_tables_kernel(const double * __restrict__ ridge,
                                   const double * __restrict__ vertex,
                                   double * __restrict__ tab,
                                   int stages,
                                   int m_max)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if(idx >= stages * 192) return;
  const int st = idx / 192, r = idx - st * 192;
  const int comp = r >> 5, j = r & 31;
  double v = 0.0;
  if(j < m_max)
  {
    const size_t src = ((size_t)st * m_max + j) * 3 + (comp % 3);
    v = comp < 3 ? ridge[src] : vertex[src];
  }
  tab[idx] = v;
}

/** copy rows of `cols` doubles between two row strides (zero-fills dst columns >= cols). */
__global__ void restride_kernel(const double * __restrict__ src, int sstride, double * __restrict__ dst, int dstride, size_t rows, int cols)
{
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(idx >= rows * (size_t)dstride) return;
  const size_t r = idx / dstride;
  const int c = (int)(idx - r * dstride);
  dst[idx] = c < cols && c < sstride ? src[r * sstride + c] : 0.0;
}
} // namespace

struct ccc_ddp_centroidal_ws
{
  int N = 0, max_batch = 0, max_sched = 0;
  int device = 0;
  int launches = 0;
  // solver workspace
  double *tab = nullptr, *xbuf = nullptr, *ubuf = nullptr, *gains = nullptr, *u32 = nullptr, *uo32 = nullptr;
  int * qslot = nullptr; // work queue: slots + head/tail/done
  int * qctl = nullptr;
  int qcap = 0;
  ccc::DdpResume * resume = nullptr;
  // device staging for CCC_MEM_HOST calls (inputs and outputs)
  int *d_sched_id = nullptr, *d_m = nullptr;
  double *d_ridge = nullptr, *d_vertex = nullptr, *d_ref = nullptr, *d_x0 = nullptr, *d_uinit = nullptr;
  double *d_x = nullptr, *d_u = nullptr, *d_cost = nullptr, *d_lambda = nullptr;
  int *d_iters = nullptr, *d_status = nullptr;
  signed char * d_alpha = nullptr;
  unsigned * d_clamped = nullptr;
  int trace_cap = 0;
  cudaStream_t own_stream = nullptr;
};

namespace
{
ccc::DdpCfg toCfg(const ccc_ddp_config_t * c)
{
  ccc::DdpCfg d;
  d.with_input_constraint = c->with_input_constraint;
  d.max_iter = c->max_iter;
  d.n_alpha = c->n_alpha;
  d.initial_lambda = c->initial_lambda;
  d.initial_dlambda = c->initial_dlambda;
  d.lambda_factor = c->lambda_factor;
  d.lambda_min = c->lambda_min;
  d.lambda_max = c->lambda_max;
  d.k_rel_norm_thre = c->k_rel_norm_thre;
  d.lambda_thre = c->lambda_thre;
  d.cost_update_ratio_thre = c->cost_update_ratio_thre;
  d.cost_update_thre = c->cost_update_thre;
  for(int i = 0; i < 16; i++) d.alpha[i] = c->alpha[i];
  d.boxqp.max_iter = c->boxqp_max_iter;
  d.boxqp.grad_thre = c->boxqp_grad_thre;
  d.boxqp.rel_improve_thre = c->boxqp_rel_improve_thre;
  d.boxqp.step_factor = c->boxqp_step_factor;
  d.boxqp.min_step = c->boxqp_min_step;
  d.boxqp.armijo = c->boxqp_armijo;
  return d;
}

template<class T>
bool devAlloc(T *& p, size_t n)
{
  return ccc_host::check(cudaMalloc(reinterpret_cast<void **>(&p), n * sizeof(T)), "cudaMalloc");
}
} // namespace

extern "C" {

ccc_ddp_centroidal_ws_t * ccc_ddp_centroidal_create(int32_t horizon_steps, int32_t max_batch, int32_t max_sched)
{
  if(horizon_steps <= 0 || max_batch <= 0 || max_sched <= 0)
  {
    ccc_host::set_error("ccc_ddp_centroidal_create: non-positive size");
    return nullptr;
  }
  int ndev = 0;
  if(!ccc_host::check(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount") || ndev == 0)
  {
    ccc_host::set_error("ccc_ddp_centroidal_create: no CUDA device (this library has no CPU fallback)");
    return nullptr;
  }
  auto * ws = new ccc_ddp_centroidal_ws();
  ws->N = horizon_steps;
  ws->max_batch = max_batch;
  ws->max_sched = max_sched;
  cudaGetDevice(&ws->device);
  const size_t N = horizon_steps, B = max_batch, S = max_sched;
  bool ok = true;
  ok = ok && devAlloc(ws->tab, S * N * 192);
  ok = ok && devAlloc(ws->xbuf, 2 * B * (N + 1) * 9);
  ok = ok && devAlloc(ws->ubuf, 2 * B * N * 32);
  ok = ok && devAlloc(ws->gains, B * N * 320);
  ok = ok && devAlloc(ws->u32, B * N * 32);
  ok = ok && devAlloc(ws->uo32, B * N * 32);
  ws->qcap = (int)(B * 18);
  ok = ok && devAlloc(ws->qslot, (size_t)ws->qcap);
  ok = ok && devAlloc(ws->qctl, 4);
  ok = ok && devAlloc(ws->resume, B);
  ok = ok && devAlloc(ws->d_sched_id, B);
  ok = ok && devAlloc(ws->d_m, S * N);
  ok = ok && devAlloc(ws->d_ridge, S * N * 32 * 3);
  ok = ok && devAlloc(ws->d_vertex, S * N * 32 * 3);
  ok = ok && devAlloc(ws->d_ref, S * (N + 1) * 3);
  ok = ok && devAlloc(ws->d_x0, B * 9);
  ok = ok && devAlloc(ws->d_uinit, B * N * 32);
  ok = ok && devAlloc(ws->d_x, B * (N + 1) * 9);
  ok = ok && devAlloc(ws->d_u, B * N * 32);
  ok = ok && devAlloc(ws->d_cost, B);
  ok = ok && devAlloc(ws->d_iters, B);
  ok = ok && devAlloc(ws->d_status, B);
  ok = ok && devAlloc(ws->d_clamped, B * N);
  ok = ok && ccc_host::check(cudaStreamCreateWithFlags(&ws->own_stream, cudaStreamNonBlocking), "cudaStreamCreate");
  for(int v = 0; v < kNumVariants; v++)
    for(int c = 0; c < 2; c++)
      ok = ok
           && ccc_host::check(cudaFuncSetAttribute(kVariants[v].kernel[c], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)(kVariants[v].warps * ccc::sm::TOTAL * sizeof(double))),
                              "cudaFuncSetAttribute(smem)");
  if(!ok)
  {
    ccc_ddp_centroidal_destroy(ws);
    return nullptr;
  }
  return ws;
}

void ccc_ddp_centroidal_destroy(ccc_ddp_centroidal_ws_t * ws)
{
  if(!ws) return;
  void * ptrs[] = {ws->tab,     ws->xbuf,   ws->ubuf,    ws->gains,  ws->u32,    ws->uo32, ws->qslot, ws->qctl, ws->resume, ws->d_sched_id,
                   ws->d_m,     ws->d_ridge, ws->d_vertex, ws->d_ref, ws->d_x0,   ws->d_uinit, ws->d_x,
                   ws->d_u,     ws->d_cost, ws->d_lambda, ws->d_iters, ws->d_status, ws->d_alpha, ws->d_clamped};
  for(void * p : ptrs)
    if(p) cudaFree(p);
  if(ws->own_stream) cudaStreamDestroy(ws->own_stream);
  delete ws;
}

int32_t ccc_ddp_centroidal_solve(ccc_ddp_centroidal_ws_t * ws,
                                 const ccc_ddp_centroidal_batch_t * bt,
                                 const ccc_ddp_config_t * cfg,
                                 ccc_ddp_result_t * res,
                                 int32_t mem,
                                 void * stream_v)
{
  using ccc_host::check;
  if(!ws || !bt || !cfg || !res) return ccc_host::fail(CCC_ERR_INVALID, "null argument");
  const int N = bt->horizon_steps, B = bt->batch, S = bt->n_sched, mm = bt->m_max;
  if(N != ws->N) return ccc_host::fail(CCC_ERR_INVALID, "horizon_steps differs from the workspace's");
  if(B > ws->max_batch || S > ws->max_sched) return ccc_host::fail(CCC_ERR_ALLOC, "batch or n_sched exceeds workspace");
  if(B <= 0 || S <= 0 || mm <= 0 || mm > CCC_DDP_M_MAX) return ccc_host::fail(CCC_ERR_INVALID, "bad batch/n_sched/m_max");
  if(cfg->reg_type != 1) return ccc_host::fail(CCC_ERR_INVALID, "only reg_type 1 is implemented");
  if(cfg->n_alpha < 1 || cfg->n_alpha > CCC_DDP_MAX_ALPHA) return ccc_host::fail(CCC_ERR_INVALID, "bad n_alpha");
  if(!bt->sched_id || !bt->m || !bt->ridge || !bt->vertex || !bt->ref_pos || !bt->x0)
    return ccc_host::fail(CCC_ERR_INVALID, "null input table");
  if(res->trace_len < 0) return ccc_host::fail(CCC_ERR_INVALID, "negative trace_len");
  cudaStream_t st = mem == CCC_MEM_HOST ? ws->own_stream : reinterpret_cast<cudaStream_t>(stream_v);
  ws->launches = 0;

  const int * sched_id = bt->sched_id;
  const int * m = bt->m;
  const double *ridge = bt->ridge, *vertex = bt->vertex, *ref = bt->ref_pos, *x0 = bt->x0, *u_init = bt->u_init;
  double *o_x = res->x, *o_u = res->u, *o_cost = res->cost, *o_lambda = res->lambda_trace;
  int *o_iters = res->iters, *o_status = res->status;
  signed char * o_alpha = reinterpret_cast<signed char *>(res->alpha_idx);
  unsigned * o_clamped = res->clamped;
  const size_t tl = (size_t)res->trace_len;

  if(mem == CCC_MEM_HOST)
  {
    // validate on the host where the tables are readable
    for(int i = 0; i < S * N; i++)
      if(m[i] < 0 || m[i] > mm) return ccc_host::fail(CCC_ERR_INVALID, "stage input dimension outside [0, m_max]");
    for(int i = 0; i < B; i++)
      if(sched_id[i] < 0 || sched_id[i] >= S) return ccc_host::fail(CCC_ERR_INVALID, "sched_id out of range");
    if(tl > 0 && (size_t)ws->trace_cap < (size_t)B * tl)
    {
      if(ws->d_alpha) cudaFree(ws->d_alpha);
      if(ws->d_lambda) cudaFree(ws->d_lambda);
      ws->d_alpha = nullptr;
      ws->d_lambda = nullptr;
      if(!devAlloc(ws->d_alpha, (size_t)ws->max_batch * tl) || !devAlloc(ws->d_lambda, (size_t)ws->max_batch * tl))
        return CCC_ERR_CUDA;
      ws->trace_cap = (int)((size_t)ws->max_batch * tl);
    }
#define CCC_H2D(dst, src, n) \
  if(!check(cudaMemcpyAsync(dst, src, (n), cudaMemcpyHostToDevice, st), "H2D")) return CCC_ERR_CUDA
    CCC_H2D(ws->d_sched_id, sched_id, sizeof(int) * B);
    CCC_H2D(ws->d_m, m, sizeof(int) * S * N);
    CCC_H2D(ws->d_ridge, ridge, sizeof(double) * S * N * mm * 3);
    CCC_H2D(ws->d_vertex, vertex, sizeof(double) * S * N * mm * 3);
    CCC_H2D(ws->d_ref, ref, sizeof(double) * S * (N + 1) * 3);
    CCC_H2D(ws->d_x0, x0, sizeof(double) * B * 9);
    if(u_init) CCC_H2D(ws->d_uinit, u_init, sizeof(double) * B * N * mm);
#undef CCC_H2D
    sched_id = ws->d_sched_id;
    m = ws->d_m;
    ridge = ws->d_ridge;
    vertex = ws->d_vertex;
    ref = ws->d_ref;
    x0 = ws->d_x0;
    if(u_init) u_init = ws->d_uinit;
    o_x = res->x ? ws->d_x : nullptr;
    o_u = res->u ? ws->d_u : nullptr;
    o_cost = res->cost ? ws->d_cost : nullptr;
    o_iters = res->iters ? ws->d_iters : nullptr;
    o_status = res->status ? ws->d_status : nullptr;
    o_alpha = (res->alpha_idx && tl) ? ws->d_alpha : nullptr;
    o_lambda = (res->lambda_trace && tl) ? ws->d_lambda : nullptr;
    o_clamped = res->clamped ? ws->d_clamped : nullptr;
  }

  // stage tables -> lane-contiguous layout
  {
    const int total = S * N * 192;
    pack_tables_kernel<<<(total + 255) / 256, 256, 0, st>>>(ridge, vertex, ws->tab, S * N, mm);
    ws->launches++;
  }
  // inputs use a row stride of 32 inside the solver
  const double * u_init32 = u_init;
  if(u_init && mm != 32)
  {
    const size_t rows = (size_t)B * N;
    restride_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(u_init, mm, ws->u32, 32, rows, mm);
    ws->launches++;
    u_init32 = ws->u32;
  }
  double * solver_out_u = (o_u && mm != 32) ? ws->uo32 : o_u;

  ccc::CentroidalParams P;
  P.N = N;
  P.B = B;
  P.S = S;
  P.dt = bt->dt;
  P.mass = bt->mass;
  P.sched_id = sched_id;
  P.m = m;
  P.tab = ws->tab;
  P.ref_pos = ref;
  for(int i = 0; i < 10; i++) P.w_run[i] = bt->w_run[i];
  for(int i = 0; i < 9; i++) P.w_term[i] = bt->w_term[i];
  P.u_lo = bt->u_lo;
  P.u_hi = bt->u_hi;
  P.x0 = x0;
  P.u_init = u_init32;
  P.cfg = toCfg(cfg);
  P.xbuf = ws->xbuf;
  P.ubuf = ws->ubuf;
  P.gains = ws->gains;
  P.out_x = o_x;
  P.out_u = solver_out_u;
  P.out_cost = o_cost;
  P.out_iters = o_iters;
  P.out_status = o_status;
  P.trace_len = (o_alpha || o_lambda) ? (int)tl : 0;
  P.out_alpha_idx = o_alpha;
  P.out_lambda = o_lambda;
  P.out_clamped = o_clamped;

  // work queue: every problem once, plus room for re-queued (suspended) solves
  const Variant & var = kVariants[g_variant];
  void (*kernel)(const ccc::CentroidalParams, const SolveQueue) = var.kernel[cfg->with_input_constraint ? 1 : 0];
  const size_t smem_bytes = (size_t)var.warps * ccc::sm::TOTAL * sizeof(double);
  SolveQueue q;
  q.slot = ws->qslot;
  q.head = ws->qctl;
  q.tail = ws->qctl + 1;
  q.done = ws->qctl + 2;
  P.chunk_iters = g_chunk;
  P.resume = ws->resume;
  if(g_chunk > 0)
  {
    // a solve is re-queued at most ceil(max_iter / chunk) - 1 times; fall back to run-to-completion
    // if that does not fit the queue allocated with the workspace
    const long long visits = ((long long)cfg->max_iter + g_chunk - 1) / g_chunk + 1;
    if(visits * B > ws->qcap) P.chunk_iters = 0;
  }
  q.capacity = P.chunk_iters > 0 ? ws->qcap : B;
  init_queue_kernel<<<(q.capacity + 255) / 256, 256, 0, st>>>(q, B);
  ws->launches++;
  // persistent grid: exactly the CTAs that are co-resident (warps spin on the queue, so every
  // launched CTA must be running)
  int n_sm = 148, per_sm = 0;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ws->device);
  if(!check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, var.warps * 32, smem_bytes), "occupancy"))
    return CCC_ERR_CUDA;
  if(per_sm < 1) return ccc_host::fail(CCC_ERR_CUDA, "solve kernel does not fit on an SM");
  if(per_sm > var.ctas) per_sm = var.ctas;
  int grid = (B + var.warps - 1) / var.warps;
  if(grid > n_sm * per_sm) grid = n_sm * per_sm;
  kernel<<<grid, var.warps * 32, smem_bytes, st>>>(P, q);
  ws->launches++;
  if(!check(cudaGetLastError(), "launch ddp_centroidal_solve_kernel")) return CCC_ERR_CUDA;

  if(o_u && solver_out_u != o_u)
  {
    const size_t rows = (size_t)B * N;
    restride_kernel<<<(unsigned)((rows * mm + 255) / 256), 256, 0, st>>>(solver_out_u, 32, o_u, mm, rows, mm);
    ws->launches++;
  }

  if(mem == CCC_MEM_HOST)
  {
#define CCC_D2H(dst, src, n) \
  if((dst) && !check(cudaMemcpyAsync(dst, src, (n), cudaMemcpyDeviceToHost, st), "D2H")) return CCC_ERR_CUDA
    CCC_D2H(res->x, ws->d_x, sizeof(double) * B * (N + 1) * 9);
    CCC_D2H(res->u, ws->d_u, sizeof(double) * B * N * mm);
    CCC_D2H(res->cost, ws->d_cost, sizeof(double) * B);
    CCC_D2H(res->iters, ws->d_iters, sizeof(int) * B);
    CCC_D2H(res->status, ws->d_status, sizeof(int) * B);
    if(tl)
    {
      CCC_D2H(res->alpha_idx, ws->d_alpha, B * tl);
      CCC_D2H(res->lambda_trace, ws->d_lambda, sizeof(double) * B * tl);
    }
    CCC_D2H(res->clamped, ws->d_clamped, sizeof(unsigned) * B * N);
#undef CCC_D2H
    if(!check(cudaStreamSynchronize(st), "cudaStreamSynchronize")) return CCC_ERR_CUDA;
  }
  return CCC_OK;
}

/* Tuning hook (not part of the stable ABI): pick the launch-shape variant, returns the count. */
int32_t ccc_ddp_centroidal_set_variant(int32_t v)
{
  if(v >= 0 && v < kNumVariants) g_variant = v;
  return kNumVariants;
}

/* Tuning hook: DDP iterations per visit before a solve is suspended and re-queued (0 = never). */
void ccc_ddp_centroidal_set_chunk(int32_t chunk)
{
  g_chunk = chunk < 0 ? 0 : chunk;
}

int32_t ccc_ddp_centroidal_last_launches(const ccc_ddp_centroidal_ws_t * ws)
{
  return ws ? ws->launches : 0;
}

} // extern "C"
