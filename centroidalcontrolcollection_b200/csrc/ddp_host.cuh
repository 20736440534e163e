// ddp_host.cuh — persistent solve kernel, work queue and host-side engine shared by the DDP
// entry points (ccc_ddp_centroidal_*, ccc_ddp_srb_*), generic over the model policy M.
//
// One persistent kernel per solve call: every warp repeatedly takes the next ticket of a device
// work queue and runs DdpWarp<M>::solve() on that problem (DDP iteration counts vary by 30x
// between problems, so static assignment would leave most warps idle).  A solve that has used
// its iteration budget for this visit is suspended and re-queued behind everything else.
// Compile with -fmad=false: only the explicit fma() calls of the cores may fuse.
#pragma once
#include <cstring>

#include "../../include/ccc_b200.h"
#include "common_host.cuh"
#include "ddp_team.cuh"
#include "ddp_warp_core.cuh"

namespace ccc_host
{
constexpr int kResumeFlag = 1 << 30;
constexpr int kTeam = 8; // warps per problem of the small-batch kernel (ddp_team.cuh)

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
  int warps_active; // warps of a CTA that take tickets (small batches are spread over the SMs instead of filling a few)
};

template<class M, int WARPS, int CTAS, bool CONSTRAINED, int FEAT = ccc::kFeatDefault>
__global__ void __launch_bounds__(WARPS * 32, CTAS)
    ddp_solve_kernel(const __grid_constant__ ccc::DdpParams<M> P, const SolveQueue q)
{
  using Warp = ccc::DdpWarp<M, CONSTRAINED, FEAT>;
  extern __shared__ __align__(16) double smem[];
  double * s = smem + (threadIdx.x >> 5) * Warp::sm::TOTAL;
  const int lane = threadIdx.x & 31;
  unsigned ring_parity = 0;
  if((int)(threadIdx.x >> 5) >= q.warps_active) return; // (no CTA-wide barrier anywhere in this kernel)
  Warp::init_warp(s);
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
    Warp w(P, s, b, &ring_parity);
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

/** Small batches (B <= number of SMs): one CTA of TEAM warps per problem, concurrent line-search rollouts
 *  (ddp_team.cuh).  Bit-identical to ddp_solve_kernel. */
template<class M, int TEAM, bool CONSTRAINED>
__global__ void __launch_bounds__(TEAM * 32, 1) ddp_team_kernel(const __grid_constant__ ccc::DdpParams<M> P)
{
  extern __shared__ __align__(16) double smem[];
  ccc::team_solve<M, CONSTRAINED, ccc::kFeatAbort, TEAM>(P, smem, (int)blockIdx.x);
}

template<class M>
constexpr size_t team_smem_bytes()
{
  return (size_t)(kTeam * ccc::SmLayout<M::NX, M::NXP>::TOTAL + ccc::team_ctl_doubles<kTeam>()) * sizeof(double);
}

static __global__ void init_queue_kernel(SolveQueue q, int B)
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

/** ridge/vertex [S][N][m_max][3] -> rows 0..5 of tab [S][N][rows][32] (component-major,
 *  lane-contiguous, zero padded).  Rows >= 6 are model specific and filled separately. */
static __global__ void pack_tables_kernel(const double * __restrict__ ridge,
                                   const double * __restrict__ vertex,
                                   double * __restrict__ tab,
                                   size_t stages,
                                   int m_max,
                                   int rows)
{
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(idx >= stages * 192) return;
  const size_t st = idx / 192;
  const int r = (int)(idx - st * 192);
  const int comp = r >> 5, j = r & 31;
  double v = 0.0;
  if(j < m_max)
  {
    const size_t src = ((size_t)st * m_max + j) * 3 + (comp % 3);
    v = comp < 3 ? ridge[src] : vertex[src];
  }
  tab[(size_t)st * rows * 32 + r] = v;
}

/** copy rows of `cols` doubles between two row strides (zero-fills dst columns >= cols). */
static __global__ void restride_kernel(const double * __restrict__ src, int sstride, double * __restrict__ dst, int dstride, size_t rows, int cols)
{
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(idx >= rows * (size_t)dstride) return;
  const size_t r = idx / dstride;
  const int c = (int)(idx - r * dstride);
  dst[idx] = c < cols && c < sstride ? src[r * sstride + c] : 0.0;
}

inline ccc::DdpCfg to_cfg(const ccc_ddp_config_t * c)
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
bool dev_alloc(T *& p, size_t n)
{
  return check(cudaMalloc(reinterpret_cast<void **>(&p), (n ? n : 1) * sizeof(T)), "cudaMalloc");
}

/** Kernel variants: the product default first, then the A/B builds of the solver core's feature bits
 *  (ddp_warp_core.cuh kFeat*; measured in profiles/r02_summary.md).  All run 8 warps per SM in one CTA: the other launch
 *  shapes of round 1 (6 x 1, 4 x 2, 12 warps) were measured slower and are gone (profiles/r01_summary.md). */
template<class M>
struct Variants
{
  struct V
  {
    int warps, ctas, feat;
    void (*kernel[2])(const ccc::DdpParams<M>, const SolveQueue); // [unconstrained, constrained]
  };
  template<int FEAT>
  static constexpr V make()
  {
    return V{8, 1, FEAT, {ddp_solve_kernel<M, 8, 1, false, FEAT>, ddp_solve_kernel<M, 8, 1, true, FEAT>}};
  }
  static const V * table(int & n)
  {
    static const V t[] = {
        make<ccc::kFeatDefault>(),
        make<ccc::kFeatAll>(), // variant 1: + TMA-staged gain lists (tests/test_gpu_ddp_centroidal.py runs it too)
#ifdef CCC_AB_VARIANTS
        make<0>(),
        make<ccc::kFeatTma>(),
#endif
    };
    n = (int)(sizeof(t) / sizeof(t[0]));
    return t;
  }
};

/** Tuning state shared by the DDP engines (not part of the stable ABI). */
inline int & g_variant()
{
  static int v = 0; // the product default (Variants<M>::table)
  return v;
}
inline int & g_team()
{
  static int t = 1; // 1: batches of at most one problem per SM run on the team kernel (ddp_team.cuh)
  return t;
}
inline int & g_spread()
{
  static int sp = 1; // 1: batches smaller than the resident warps are spread over all SMs
  return sp;
}
inline int & g_packed_io()
{
  static int p = 1; // 1: host-buffer calls whose buffers fit the staging block move through it (DdpEngine::solve)
  return p;
}
inline int & g_chunk()
{
  static int c = 64; // DDP iterations per visit before a solve is suspended and re-queued (sweep: profiles/r02i_ab_chunk*.txt)
  return c;
}

/** What the model-specific entry point hands to the engine: host or device pointers (per `mem`). */
template<class M>
struct DdpInputs
{
  int B = 0, S = 0, m_max = 0;
  const int * sched_id = nullptr; // [B]
  const int * m = nullptr;        // [S][N]
  const double * ridge = nullptr; // [S][N][m_max][3]
  const double * vertex = nullptr;
  const double * ref = nullptr;   // [S][N+1][NREF]
  const double * x0 = nullptr;    // [B][NX]
  const double * u_init = nullptr; // [B][N][m_max] or null
  double w_run[M::NX + 1], w_term[M::NX];
  double u_lo = 0, u_hi = 0;
  typename M::Params mp;
};

template<class M>
struct DdpEngine
{
  static constexpr int NX = M::NX;
  int N = 0, max_batch = 0, max_sched = 0, device = 0, launches = 0, n_sm = 148;
  int last_team = 0; // 1 if the last solve ran on the team kernel
  // solver workspace
  double *tab = nullptr, *xbuf = nullptr, *ubuf = nullptr, *gains = nullptr, *u32 = nullptr, *uo32 = nullptr;
  int *qslot = nullptr, *qctl = nullptr;
  int qcap = 0;
  ccc::DdpResume * resume = nullptr;
  // device staging for CCC_MEM_HOST calls (inputs and outputs)
  int *d_sched_id = nullptr, *d_m = nullptr;
  double *d_ridge = nullptr, *d_vertex = nullptr, *d_ref = nullptr, *d_x0 = nullptr, *d_uinit = nullptr;
  double *d_x = nullptr, *d_u = nullptr, *d_cost = nullptr, *d_lambda = nullptr;
  int *d_iters = nullptr, *d_status = nullptr;
  signed char * d_alpha = nullptr;
  unsigned * d_clamped = nullptr;
  size_t trace_cap = 0;
  cudaStream_t own_stream = nullptr;
  // small host-buffer calls (a planOnce-sized batch): every input goes through ONE pinned staging block and one H2D copy,
  // every output through one D2H copy — a dozen small cudaMemcpyAsync calls from pageable memory cost more than the copies
  static constexpr size_t kStageBytes = 1u << 20;
  char *h_stage = nullptr, *d_stage = nullptr;
  int last_packed = 0; // 1 if the last host-buffer solve went through the staging block

  bool create(int horizon_steps, int B_, int S_)
  {
    int ndev = 0;
    if(!check(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount") || ndev == 0)
    {
      set_error("no CUDA device (this library has no CPU fallback)");
      return false;
    }
    N = horizon_steps;
    max_batch = B_;
    max_sched = S_;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    const size_t n = horizon_steps, B = B_, S = S_;
    // trajectory buffers: nominal + candidate per problem; the team kernel keeps kTeam candidates for up to n_sm problems
    const size_t team_B = B < (size_t)n_sm ? B : (size_t)n_sm;
    const size_t traj = 2 * B > (kTeam + 1) * team_B ? 2 * B : (kTeam + 1) * team_B;
    bool ok = true;
    ok = ok && dev_alloc(tab, S * n * 32 * M::TAB_ROWS);
    ok = ok && dev_alloc(xbuf, traj * (n + 1) * NX);
    ok = ok && dev_alloc(ubuf, traj * n * 32);
    ok = ok && dev_alloc(gains, B * n * 32 * M::NXP);
    ok = ok && dev_alloc(u32, B * n * 32);
    ok = ok && dev_alloc(uo32, B * n * 32);
    qcap = (int)(B * 18);
    ok = ok && dev_alloc(qslot, (size_t)qcap);
    ok = ok && dev_alloc(qctl, 4);
    ok = ok && dev_alloc(resume, B);
    ok = ok && dev_alloc(d_sched_id, B);
    ok = ok && dev_alloc(d_m, S * n);
    ok = ok && dev_alloc(d_ridge, S * n * 32 * 3);
    ok = ok && dev_alloc(d_vertex, S * n * 32 * 3);
    ok = ok && dev_alloc(d_ref, S * (n + 1) * M::NREF);
    ok = ok && dev_alloc(d_x0, B * NX);
    ok = ok && dev_alloc(d_uinit, B * n * 32);
    ok = ok && dev_alloc(d_x, B * (n + 1) * NX);
    ok = ok && dev_alloc(d_u, B * n * 32);
    ok = ok && dev_alloc(d_cost, B);
    ok = ok && dev_alloc(d_iters, B);
    ok = ok && dev_alloc(d_status, B);
    ok = ok && dev_alloc(d_clamped, B * n);
    ok = ok && check(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking), "cudaStreamCreate");
    ok = ok && check(cudaMallocHost(reinterpret_cast<void **>(&h_stage), kStageBytes), "cudaMallocHost") && dev_alloc(d_stage, kStageBytes);
    int nv = 0;
    const auto * vt = Variants<M>::table(nv);
    for(int v = 0; v < nv && ok; v++)
      for(int c = 0; c < 2 && ok; c++)
        ok = check(cudaFuncSetAttribute(vt[v].kernel[c], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(vt[v].warps * ccc::SmLayout<M::NX, M::NXP>::TOTAL * sizeof(double))),
                   "cudaFuncSetAttribute(smem)");
    ok = ok && check(cudaFuncSetAttribute(ddp_team_kernel<M, kTeam, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)team_smem_bytes<M>()),
                     "cudaFuncSetAttribute(team smem)");
    ok = ok && check(cudaFuncSetAttribute(ddp_team_kernel<M, kTeam, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)team_smem_bytes<M>()),
                     "cudaFuncSetAttribute(team smem)");
    return ok;
  }

  void destroy()
  {
    void * ptrs[] = {tab,     xbuf,     ubuf,  gains, u32,    uo32,   qslot,   qctl,    resume,   d_sched_id, d_m,      d_ridge,
                     d_vertex, d_ref,   d_x0,  d_uinit, d_x,  d_u,    d_cost,  d_lambda, d_iters, d_status,   d_alpha,  d_clamped};
    for(void * p : ptrs)
      if(p) cudaFree(p);
    if(own_stream) cudaStreamDestroy(own_stream);
    if(h_stage) cudaFreeHost(h_stage);
    if(d_stage) cudaFree(d_stage);
  }

  /** Everything of the kernel parameters that does not depend on where inputs / outputs live (device pointers
   *  in `in`); the caller sets u_init and the out_* pointers. */
  ccc::DdpParams<M> make_params(const DdpInputs<M> & in, const ccc_ddp_config_t * cfg) const
  {
    ccc::DdpParams<M> P;
    P.N = N;
    P.B = in.B;
    P.S = in.S;
    P.tab_len = N;
    P.ref_len = N + 1;
    P.tab_off = 0;
    P.tab_stride = 1;
    P.sched_id = in.sched_id;
    P.m = in.m;
    P.tab = tab;
    P.ref = in.ref;
    for(int i = 0; i <= NX; i++) P.w_run[i] = in.w_run[i];
    for(int i = 0; i < NX; i++) P.w_term[i] = in.w_term[i];
    P.mp = in.mp;
    P.u_lo = in.u_lo;
    P.u_hi = in.u_hi;
    P.x0 = in.x0;
    P.u_init = nullptr;
    P.cfg = to_cfg(cfg);
    P.chunk_iters = 0;
    // early abort of hopeless line-search rollouts needs costs that only grow along the horizon and the default
    // acceptance threshold
    P.abort_ok = cfg->cost_update_ratio_thre >= 0.0 ? 1 : 0;
    for(int i = 0; i <= NX; i++)
      if(!(in.w_run[i] >= 0.0)) P.abort_ok = 0;
    for(int i = 0; i < NX; i++)
      if(!(in.w_term[i] >= 0.0)) P.abort_ok = 0;
    P.xbuf = xbuf;
    P.ubuf = ubuf;
    P.gains = gains;
    P.resume = resume;
    P.out_x = nullptr;
    P.out_u = nullptr;
    P.out_cost = nullptr;
    P.out_iters = nullptr;
    P.out_status = nullptr;
    P.trace_len = 0;
    P.out_alpha_idx = nullptr;
    P.out_lambda = nullptr;
    P.out_clamped = nullptr;
    return P;
  }

  /** Work queue + persistent solve kernel for the problems described by P (all pointers on the device). */
  int launch_solve(ccc::DdpParams<M> & P, const ccc_ddp_config_t * cfg, cudaStream_t st)
  {
    const int B = P.B;
    last_team = 0;
    if(g_team() && B <= n_sm)
    {
      // one problem per SM: the team kernel (no queue, no suspension)
      P.chunk_iters = 0;
      if(cfg->with_input_constraint)
        ddp_team_kernel<M, kTeam, true><<<B, kTeam * 32, team_smem_bytes<M>(), st>>>(P);
      else
        ddp_team_kernel<M, kTeam, false><<<B, kTeam * 32, team_smem_bytes<M>(), st>>>(P);
      launches++;
      last_team = 1;
      if(!check(cudaGetLastError(), "launch ddp_team_kernel")) return CCC_ERR_CUDA;
      return CCC_OK;
    }
    // work queue: every problem once, plus room for re-queued (suspended) solves
    int nv = 0;
    const auto * vt = Variants<M>::table(nv);
    const auto & var = vt[g_variant() < nv ? g_variant() : 0];
    auto kernel = var.kernel[cfg->with_input_constraint ? 1 : 0];
    const size_t smem_bytes = (size_t)var.warps * ccc::SmLayout<M::NX, M::NXP>::TOTAL * sizeof(double);
    SolveQueue q;
    q.slot = qslot;
    q.head = qctl;
    q.tail = qctl + 1;
    q.done = qctl + 2;
    P.chunk_iters = g_chunk();
    if(P.chunk_iters > 0)
    {
      // a solve is re-queued at most ceil(max_iter / chunk) - 1 times; fall back to
      // run-to-completion if that does not fit the queue allocated with the workspace
      const long long visits = ((long long)cfg->max_iter + P.chunk_iters - 1) / P.chunk_iters + 1;
      if(visits * B > qcap) P.chunk_iters = 0;
    }
    q.capacity = P.chunk_iters > 0 ? qcap : B;
    init_queue_kernel<<<(q.capacity + 255) / 256, 256, 0, st>>>(q, B);
    launches++;
    // persistent grid: exactly the CTAs that are co-resident (warps spin on the queue, so every
    // launched CTA must be running)
    int per_sm = 0;
    if(!check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, var.warps * 32, smem_bytes), "occupancy"))
      return CCC_ERR_CUDA;
    if(per_sm < 1) return fail(CCC_ERR_CUDA, "solve kernel does not fit on an SM");
    if(per_sm > var.ctas) per_sm = var.ctas;
    int grid = (B + var.warps - 1) / var.warps;
    if(grid > n_sm * per_sm) grid = n_sm * per_sm;
    q.warps_active = var.warps;
    if(g_spread() && B < n_sm * per_sm * var.warps)
    {
      // fewer problems than resident warps: use every SM with ceil(B / grid) warps each (a warp on a less crowded SM
      // runs its serial recursion faster) instead of eight warps on B / 8 SMs
      grid = B < n_sm * per_sm ? B : n_sm * per_sm;
      q.warps_active = (B + grid - 1) / grid;
    }
    kernel<<<grid, var.warps * 32, smem_bytes, st>>>(P, q);
    launches++;
    if(!check(cudaGetLastError(), "launch ddp_solve_kernel")) return CCC_ERR_CUDA;
    return CCC_OK;
  }

  /** Validate, stage (host mode), pack, launch, collect.  `extra_pack(stream, in_dev)` is called
   *  after the common table rows are packed, with device pointers, to fill model-specific rows. */
  template<class ExtraPack>
  int solve(DdpInputs<M> in, const ccc_ddp_config_t * cfg, ccc_ddp_result_t * res, int mem, void * stream_v, ExtraPack && extra_pack)
  {
    const int B = in.B, S = in.S, mm = in.m_max;
    if(B > max_batch || S > max_sched) return fail(CCC_ERR_ALLOC, "batch or n_sched exceeds workspace");
    if(B <= 0 || S <= 0 || mm <= 0 || mm > CCC_DDP_M_MAX) return fail(CCC_ERR_INVALID, "bad batch/n_sched/m_max");
    if(cfg->reg_type != 1) return fail(CCC_ERR_INVALID, "only reg_type 1 is implemented");
    if(cfg->n_alpha < 1 || cfg->n_alpha > CCC_DDP_MAX_ALPHA) return fail(CCC_ERR_INVALID, "bad n_alpha");
    if(!in.sched_id || !in.m || !in.ridge || !in.vertex || !in.ref || !in.x0) return fail(CCC_ERR_INVALID, "null input table");
    if(res->trace_len < 0) return fail(CCC_ERR_INVALID, "negative trace_len");
    cudaStream_t st = mem == CCC_MEM_HOST ? own_stream : reinterpret_cast<cudaStream_t>(stream_v);
    launches = 0;

    double *o_x = res->x, *o_u = res->u, *o_cost = res->cost, *o_lambda = res->lambda_trace;
    int *o_iters = res->iters, *o_status = res->status;
    signed char * o_alpha = reinterpret_cast<signed char *>(res->alpha_idx);
    unsigned * o_clamped = res->clamped;
    const size_t tl = (size_t)res->trace_len;

    // layout of the staging block (offsets in bytes, 16-byte aligned): inputs first, outputs behind them
    struct Blk
    {
      const void * src;
      void * dst;
      size_t bytes, off;
    };
    Blk bin[7] = {{in.sched_id, nullptr, sizeof(int) * (size_t)B, 0},
                  {in.m, nullptr, sizeof(int) * (size_t)S * N, 0},
                  {in.ridge, nullptr, sizeof(double) * (size_t)S * N * mm * 3, 0},
                  {in.vertex, nullptr, sizeof(double) * (size_t)S * N * mm * 3, 0},
                  {in.ref, nullptr, sizeof(double) * (size_t)S * (N + 1) * M::NREF, 0},
                  {in.x0, nullptr, sizeof(double) * (size_t)B * NX, 0},
                  {in.u_init, nullptr, in.u_init ? sizeof(double) * (size_t)B * N * mm : 0, 0}};
    Blk bout[8] = {{nullptr, res->x, sizeof(double) * (size_t)B * (N + 1) * NX, 0},
                   {nullptr, res->u, sizeof(double) * (size_t)B * N * mm, 0},
                   {nullptr, res->cost, sizeof(double) * (size_t)B, 0},
                   {nullptr, res->iters, sizeof(int) * (size_t)B, 0},
                   {nullptr, res->status, sizeof(int) * (size_t)B, 0},
                   {nullptr, tl ? res->alpha_idx : nullptr, (size_t)B * tl, 0},
                   {nullptr, tl ? res->lambda_trace : nullptr, sizeof(double) * (size_t)B * tl, 0},
                   {nullptr, res->clamped, sizeof(unsigned) * (size_t)B * N, 0}};
    size_t in_total = 0, all_total = 0;
    bool packed = false;
    last_packed = 0;

    if(mem == CCC_MEM_HOST)
    {
      for(int i = 0; i < S * N; i++)
        if(in.m[i] < 0 || in.m[i] > mm) return fail(CCC_ERR_INVALID, "stage input dimension outside [0, m_max]");
      for(int i = 0; i < B; i++)
        if(in.sched_id[i] < 0 || in.sched_id[i] >= S) return fail(CCC_ERR_INVALID, "sched_id out of range");
      size_t off = 0;
      for(Blk & b : bin)
      {
        b.off = off;
        off += (b.bytes + 15) & ~(size_t)15;
      }
      in_total = off;
      for(Blk & b : bout)
      {
        if(!b.dst) b.bytes = 0;
        b.off = off;
        off += (b.bytes + 15) & ~(size_t)15;
      }
      all_total = off;
      packed = g_packed_io() && all_total <= kStageBytes;
    }
    if(packed)
    {
      for(const Blk & b : bin)
        if(b.bytes) std::memcpy(h_stage + b.off, b.src, b.bytes);
      if(!check(cudaMemcpyAsync(d_stage, h_stage, in_total, cudaMemcpyHostToDevice, st), "H2D")) return CCC_ERR_CUDA;
      in.sched_id = reinterpret_cast<const int *>(d_stage + bin[0].off);
      in.m = reinterpret_cast<const int *>(d_stage + bin[1].off);
      in.ridge = reinterpret_cast<const double *>(d_stage + bin[2].off);
      in.vertex = reinterpret_cast<const double *>(d_stage + bin[3].off);
      in.ref = reinterpret_cast<const double *>(d_stage + bin[4].off);
      in.x0 = reinterpret_cast<const double *>(d_stage + bin[5].off);
      if(in.u_init) in.u_init = reinterpret_cast<const double *>(d_stage + bin[6].off);
      o_x = res->x ? reinterpret_cast<double *>(d_stage + bout[0].off) : nullptr;
      o_u = res->u ? reinterpret_cast<double *>(d_stage + bout[1].off) : nullptr;
      o_cost = res->cost ? reinterpret_cast<double *>(d_stage + bout[2].off) : nullptr;
      o_iters = res->iters ? reinterpret_cast<int *>(d_stage + bout[3].off) : nullptr;
      o_status = res->status ? reinterpret_cast<int *>(d_stage + bout[4].off) : nullptr;
      o_alpha = (res->alpha_idx && tl) ? reinterpret_cast<signed char *>(d_stage + bout[5].off) : nullptr;
      o_lambda = (res->lambda_trace && tl) ? reinterpret_cast<double *>(d_stage + bout[6].off) : nullptr;
      o_clamped = res->clamped ? reinterpret_cast<unsigned *>(d_stage + bout[7].off) : nullptr;
      last_packed = 1;
    }
    else if(mem == CCC_MEM_HOST)
    {
      if(tl > 0 && trace_cap < (size_t)max_batch * tl)
      {
        if(d_alpha) cudaFree(d_alpha);
        if(d_lambda) cudaFree(d_lambda);
        d_alpha = nullptr;
        d_lambda = nullptr;
        if(!dev_alloc(d_alpha, (size_t)max_batch * tl) || !dev_alloc(d_lambda, (size_t)max_batch * tl)) return CCC_ERR_CUDA;
        trace_cap = (size_t)max_batch * tl;
      }
#define CCC_H2D(dst, src, n) \
  if(!check(cudaMemcpyAsync(dst, src, (n), cudaMemcpyHostToDevice, st), "H2D")) return CCC_ERR_CUDA
      CCC_H2D(d_sched_id, in.sched_id, sizeof(int) * B);
      CCC_H2D(d_m, in.m, sizeof(int) * S * N);
      CCC_H2D(d_ridge, in.ridge, sizeof(double) * S * N * mm * 3);
      CCC_H2D(d_vertex, in.vertex, sizeof(double) * S * N * mm * 3);
      CCC_H2D(d_ref, in.ref, sizeof(double) * S * (N + 1) * M::NREF);
      CCC_H2D(d_x0, in.x0, sizeof(double) * B * NX);
      if(in.u_init) CCC_H2D(d_uinit, in.u_init, sizeof(double) * B * N * mm);
#undef CCC_H2D
      in.sched_id = d_sched_id;
      in.m = d_m;
      in.ridge = d_ridge;
      in.vertex = d_vertex;
      in.ref = d_ref;
      in.x0 = d_x0;
      if(in.u_init) in.u_init = d_uinit;
      o_x = res->x ? d_x : nullptr;
      o_u = res->u ? d_u : nullptr;
      o_cost = res->cost ? d_cost : nullptr;
      o_iters = res->iters ? d_iters : nullptr;
      o_status = res->status ? d_status : nullptr;
      o_alpha = (res->alpha_idx && tl) ? d_alpha : nullptr;
      o_lambda = (res->lambda_trace && tl) ? d_lambda : nullptr;
      o_clamped = res->clamped ? d_clamped : nullptr;
    }

    // stage tables -> lane-contiguous layout
    {
      const size_t total = (size_t)S * N * 192;
      pack_tables_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in.ridge, in.vertex, tab, (size_t)S * N, mm, M::TAB_ROWS);
      launches++;
      launches += extra_pack(st, tab);
    }
    // inputs use a row stride of 32 inside the solver
    const double * u_init32 = in.u_init;
    if(in.u_init && mm != 32)
    {
      const size_t rows = (size_t)B * N;
      restride_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(in.u_init, mm, u32, 32, rows, mm);
      launches++;
      u_init32 = u32;
    }
    double * solver_out_u = (o_u && mm != 32) ? uo32 : o_u;

    ccc::DdpParams<M> P = make_params(in, cfg);
    P.u_init = u_init32;
    P.out_x = o_x;
    P.out_u = solver_out_u;
    P.out_cost = o_cost;
    P.out_iters = o_iters;
    P.out_status = o_status;
    P.trace_len = (o_alpha || o_lambda) ? (int)tl : 0;
    P.out_alpha_idx = o_alpha;
    P.out_lambda = o_lambda;
    P.out_clamped = o_clamped;
    const int rc = launch_solve(P, cfg, st);
    if(rc != CCC_OK) return rc;

    if(o_u && solver_out_u != o_u)
    {
      const size_t rows = (size_t)B * N;
      restride_kernel<<<(unsigned)((rows * mm + 255) / 256), 256, 0, st>>>(solver_out_u, 32, o_u, mm, rows, mm);
      launches++;
    }

    if(packed)
    {
      if(all_total > in_total
         && !check(cudaMemcpyAsync(h_stage + in_total, d_stage + in_total, all_total - in_total, cudaMemcpyDeviceToHost, st), "D2H"))
        return CCC_ERR_CUDA;
      if(!check(cudaStreamSynchronize(st), "cudaStreamSynchronize")) return CCC_ERR_CUDA;
      for(const Blk & b : bout)
        if(b.bytes) std::memcpy(b.dst, h_stage + b.off, b.bytes);
    }
    else if(mem == CCC_MEM_HOST)
    {
#define CCC_D2H(dst, src, n) \
  if((dst) && !check(cudaMemcpyAsync(dst, src, (n), cudaMemcpyDeviceToHost, st), "D2H")) return CCC_ERR_CUDA
      CCC_D2H(res->x, d_x, sizeof(double) * B * (N + 1) * NX);
      CCC_D2H(res->u, d_u, sizeof(double) * B * N * mm);
      CCC_D2H(res->cost, d_cost, sizeof(double) * B);
      CCC_D2H(res->iters, d_iters, sizeof(int) * B);
      CCC_D2H(res->status, d_status, sizeof(int) * B);
      if(tl)
      {
        CCC_D2H(res->alpha_idx, d_alpha, B * tl);
        CCC_D2H(res->lambda_trace, d_lambda, sizeof(double) * B * tl);
      }
      CCC_D2H(res->clamped, d_clamped, sizeof(unsigned) * B * N);
#undef CCC_D2H
      if(!check(cudaStreamSynchronize(st), "cudaStreamSynchronize")) return CCC_ERR_CUDA;
    }
    return CCC_OK;
  }
};
} // namespace ccc_host
