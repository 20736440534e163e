// qp.cu — kernels + C-ABI host code of ccc_qp_* (include/ccc_b200.h): batched strictly convex dense
// QP with shared matrices, one CTA per problem (qp_cta_core.cuh), persistent CTAs pulling problems
// from an atomic counter.
#include "../../include/ccc_b200.h"
#include "qp_host.cuh"

namespace
{
/** One CTA per matrix group g = blockIdx.x: Q, A and every output are per group, C (and its transpose) is shared
 *  and written by group 0 only. */
__global__ void __launch_bounds__(ccc::kQpThreads, 1) qp_setup_kernel(int n, int me, int mi, const double * Q, const double * A,
                                                                       const double * C, double * Lg, double * invd, double * J0,
                                                                       double * At, double * Ct, int * ok_flag, double * J0s)
{
  __shared__ double s_bcast[256];
  const size_t g = blockIdx.x, nn = (size_t)n * n;
  ccc::qp_setup_cta(n, me, mi, Q + g * nn, A ? A + g * me * n : nullptr, C, Lg + g * nn, invd + g * n, J0 + g * nn, At + g * n * me, Ct,
                    ok_flag + 4 * g, J0s ? J0s + g * n * (n | 1) : nullptr, g == 0, s_bcast);
}

template<int NT, bool kGlobal, bool kPackedR>
__global__ void __launch_bounds__(NT, kPackedR ? 2 : 1) qp_solve_kernel(const __grid_constant__ ccc::QpParams P, int * __restrict__ counter,
                                                                        double * __restrict__ gmat)
{
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_b;
  __shared__ __align__(8) unsigned long long s_mbar; // completion barrier of the per-problem TMA load of J
  double * slab = kGlobal ? gmat + (size_t)blockIdx.x * 2 * P.n * P.ld : nullptr;
  if(!kGlobal && threadIdx.x == 0) ccc::mbar_init(&s_mbar, 1);
  __syncthreads();
  unsigned mbar_phase = 0;
  // the fallback pass works through the list of problems whose active set outgrew the packed R of the first pass
  const int total = P.list ? *P.list_count : P.B;
  for(;;)
  {
    if(threadIdx.x == 0) s_b = atomicAdd(counter, 1);
    __syncthreads();
    const int t = s_b;
    __syncthreads();
    if(t >= total) break;
    const int b = P.list ? P.list[t] : t;
    ccc::QpCta<NT, kGlobal, kPackedR> cta(P, smem, b, slab);
    cta.mbar = kGlobal ? nullptr : &s_mbar;
    cta.mbar_phase = mbar_phase;
    cta.solve();
    mbar_phase = cta.mbar_phase;
  }
}

constexpr size_t kSmemLimit = 227 * 1024;
constexpr size_t kSmemDynMax = kSmemLimit - 1024; // the opt-in limit covers static shared memory too (the kernels hold a few bytes)
constexpr size_t kSmemTwoCtas = (228 * 1024 - 2 * 1024) / 2 - 128; // dynamic shared memory of one of two co-resident CTAs

/** Largest number of columns a packed R may have so that two CTAs fit on an SM (0: not even the smallest useful one). */
int qp_rcap(int n, int me)
{
  const int ld = n | 1;
  int cap = 0;
  for(int c = 8; c <= n; c++)
    if(ccc::QpSm<128, false, true>::bytes(n, ld, c) <= kSmemTwoCtas) cap = c;
  return cap > me + 8 ? cap : 0;
}

/** 0: 128 threads, J/R in shared memory; 1: 128 threads, J/R in global memory; 2: 256 threads, global. */
int qp_shape(int n)
{
  if(n > 128) return 2;
  return ccc::QpSm<128, false, false>::bytes(n, n | 1) <= kSmemDynMax ? 0 : 1;
}

template<class T>
bool dev_alloc(T *& p, size_t n)
{
  return ccc_host::check(cudaMalloc(reinterpret_cast<void **>(&p), (n ? n : 1) * sizeof(T)), "cudaMalloc");
}
} // namespace

namespace ccc_host
{
ccc_qp_ws * qp_ws_create(int n, int n_eq, int n_ineq, int max_batch, int max_groups, bool staging)
{
  const int nt_threads = n > 128 ? 256 : 128;
  if(n <= 0 || n > 256 || n_eq < 0 || n_eq > n || n_ineq <= 0 || n_eq + n_ineq > 4 * nt_threads || max_batch <= 0 || max_groups <= 0)
  {
    ccc_host::set_error("ccc_qp_create: sizes outside the kernel's limits (n <= 256, n_eq <= n, n_eq + n_ineq <= 4 x threads)");
    return nullptr;
  }
  int ndev = 0;
  if(!ccc_host::check(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount") || ndev == 0)
  {
    ccc_host::set_error("ccc_qp_create: no CUDA device (this library has no CPU fallback)");
    return nullptr;
  }
  auto * ws = new ccc_qp_ws();
  ws->n = n;
  ws->me = n_eq;
  ws->mi = n_ineq;
  ws->max_batch = max_batch;
  ws->max_groups = max_groups;
  cudaGetDevice(&ws->device);
  const size_t N = n, ME = n_eq, MI = n_ineq, B = max_batch, G = max_groups;
  bool ok = true;
  ok = ok && dev_alloc(ws->Lg, G * N * N) && dev_alloc(ws->invd, G * N) && dev_alloc(ws->J0, G * N * N);
  ok = ok && (qp_shape(n) != 0 || dev_alloc(ws->J0s, G * N * (N | 1))); // the TMA source of the shared-memory kernels
  ok = ok && dev_alloc(ws->At, G * N * ME) && dev_alloc(ws->Ct, N * MI) && dev_alloc(ws->ok_flag, 4 * G) && dev_alloc(ws->counter, 4) && dev_alloc(ws->ovf_list, B);
  if(staging)
  {
    ok = ok && dev_alloc(ws->d_Q, G * N * N) && dev_alloc(ws->d_A, G * N * ME) && dev_alloc(ws->d_C, N * MI) && dev_alloc(ws->d_grp, B);
    ok = ok && dev_alloc(ws->d_c, B * N) && dev_alloc(ws->d_b, B * ME) && dev_alloc(ws->d_d, B * MI) && dev_alloc(ws->d_x, B * N);
    ok = ok && dev_alloc(ws->d_iters, B) && dev_alloc(ws->d_status, B) && dev_alloc(ws->d_nact, B) && dev_alloc(ws->d_active, B * N);
  }
  ok = ok && ccc_host::check(cudaStreamCreateWithFlags(&ws->own_stream, cudaStreamNonBlocking), "cudaStreamCreate");
  ok = ok && ccc_host::check(cudaStreamCreateWithFlags(&ws->h2d_stream, cudaStreamNonBlocking), "cudaStreamCreate");
  ok = ok && ccc_host::check(cudaStreamCreateWithFlags(&ws->d2h_stream, cudaStreamNonBlocking), "cudaStreamCreate");
  ok = ok && ccc_host::check(cudaEventCreateWithFlags(&ws->ev_setup, cudaEventDisableTiming), "cudaEventCreate");
  const int nchunk_max = max_batch <= 16384 ? 1 : (max_batch + 32767) / 32768;
  for(int k = 0; ok && nchunk_max > 1 && k < nchunk_max && k < ccc_qp_ws::kMaxChunks; k++)
    ok = ccc_host::check(cudaEventCreateWithFlags(&ws->ev_h2d[k], cudaEventDisableTiming), "cudaEventCreate")
         && ccc_host::check(cudaEventCreateWithFlags(&ws->ev_solved[k], cudaEventDisableTiming), "cudaEventCreate");
  const int ld = n | 1;
  cudaDeviceGetAttribute(&ws->n_sm, cudaDevAttrMultiProcessorCount, ws->device);
  const int shape = qp_shape(n);
  if(shape == 0)
  {
    // the attribute belongs to the kernel, not to this workspace: workspaces of several shapes are alive at once (a
    // controller whose QP size changes from tick to tick), so it is raised to the limit, not to this shape's need
    ok = ok
         && ccc_host::check(cudaFuncSetAttribute(qp_solve_kernel<128, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemDynMax),
                            "cudaFuncSetAttribute(smem)");
    ws->rcap = qp_rcap(n, n_eq);
    if(ws->rcap > 0)
      ok = ok
           && ccc_host::check(cudaFuncSetAttribute(qp_solve_kernel<128, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemDynMax),
                              "cudaFuncSetAttribute(smem, packed R)")
           && ccc_host::check(cudaFuncSetAttribute(qp_solve_kernel<128, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100),
                              "cudaFuncSetAttribute(carveout)");
  }
  else
    ok = ok && dev_alloc(ws->gmat, (size_t)ws->n_sm * 2 * N * ld);
  if(!ok)
  {
    ccc_qp_destroy(ws);
    return nullptr;
  }
  return ws;
}


int qp_setup_launch(ccc_qp_ws * ws, int groups, const double * Q, const double * A, const double * C, cudaStream_t st)
{
  if(groups <= 0 || groups > ws->max_groups) return fail(CCC_ERR_ALLOC, "more matrix groups than the QP workspace holds");
  qp_setup_kernel<<<groups, ccc::kQpThreads, 0, st>>>(ws->n, ws->me, ws->mi, Q, A, C, ws->Lg, ws->invd, ws->J0, ws->At, ws->Ct, ws->ok_flag,
                                                     ws->J0s);
  ws->launches++;
  ws->have_setup = true;
  if(!check(cudaGetLastError(), "launch qp_setup_kernel")) return CCC_ERR_CUDA;
  return CCC_OK;
}

ccc::QpParams qp_params(const ccc_qp_ws * ws, int B, const int * grp)
{
  ccc::QpParams P;
  const size_t n = ws->n;
  P.n = ws->n;
  P.me = ws->me;
  P.mi = ws->mi;
  P.B = B;
  P.ld = ws->n | 1;
  P.J0 = ws->J0;
  P.J0s = ws->J0s;
  P.At = ws->At;
  P.Ct = ws->Ct;
  P.setup_ok = ws->ok_flag;
  P.max_iter = 1000;
  P.viol_tol = 1e-10;
  P.c = P.b = P.d = nullptr;
  P.out_x = nullptr;
  P.out_iters = P.out_status = P.out_n_active = P.out_active = nullptr;
  P.grp = grp;
  if(grp)
  {
    P.gs_J0 = n * n;
    P.gs_J0s = n * (n | 1);
    P.gs_At = n * ws->me;
    P.gs_Ct = 0;
    P.gs_ok = 4;
  }
  return P;
}
} // namespace ccc_host

extern "C" {

ccc_qp_ws_t * ccc_qp_create(int32_t n, int32_t n_eq, int32_t n_ineq, int32_t max_batch)
{
  return ccc_host::qp_ws_create(n, n_eq, n_ineq, max_batch, 1, true);
}

void ccc_qp_destroy(ccc_qp_ws_t * ws)
{
  if(!ws) return;
  void * ptrs[] = {ws->d_grp, ws->ovf_list, ws->J0s, ws->gmat, ws->Lg,  ws->invd, ws->J0,  ws->At,  ws->Ct,     ws->ok_flag, ws->counter, ws->d_Q,    ws->d_A,
                   ws->d_C, ws->d_c,  ws->d_b, ws->d_d, ws->d_x,    ws->d_iters, ws->d_status, ws->d_nact, ws->d_active};
  for(void * p : ptrs)
    if(p) cudaFree(p);
  if(ws->own_stream) cudaStreamDestroy(ws->own_stream);
  if(ws->h2d_stream) cudaStreamDestroy(ws->h2d_stream);
  if(ws->d2h_stream) cudaStreamDestroy(ws->d2h_stream);
  if(ws->ev_setup) cudaEventDestroy(ws->ev_setup);
  for(int k = 0; k < ccc_qp_ws::kMaxChunks; k++)
  {
    if(ws->ev_h2d[k]) cudaEventDestroy(ws->ev_h2d[k]);
    if(ws->ev_solved[k]) cudaEventDestroy(ws->ev_solved[k]);
  }
  delete ws;
}

} // extern "C"

/** Launch the solve kernels for problems [lo, lo + nb) of the batch described by P (device pointers, per-problem
 *  arrays already offset by the caller) on stream st. */
int ccc_host::qp_launch(ccc_qp_ws * ws, ccc::QpParams P, cudaStream_t st)
{
  const int n = P.n, B = P.B;
  if(!check(cudaMemsetAsync(ws->counter, 0, 4 * sizeof(int), st), "memset")) return CCC_ERR_CUDA;
  const int grid = B < ws->n_sm ? B : ws->n_sm;
  switch(qp_shape(n))
  {
    case 0:
      if(ws->rcap > 0 && ccc_host::g_qp_packed())
      {
        // first pass: two CTAs per SM with a packed R of rcap columns; problems that need more are listed ...
        ccc::QpParams P1 = P;
        P1.rcap = ws->rcap < n ? ws->rcap : n;
        P1.ovf_count = ws->counter + 2;
        P1.ovf_list = ws->ovf_list;
        const int grid2 = B < 2 * ws->n_sm ? B : 2 * ws->n_sm;
        qp_solve_kernel<128, false, true><<<grid2, 128, ccc::QpSm<128, false, true>::bytes(n, P.ld, ws->rcap), st>>>(P1, ws->counter, nullptr);
        ws->launches++;
        if(P1.rcap >= n) break;
        // ... and solved again from scratch with the full R (the grid finds an empty list in almost every call)
        P.list = ws->ovf_list;
        P.list_count = ws->counter + 2;
        qp_solve_kernel<128, false, false><<<ws->n_sm, 128, ccc::QpSm<128, false, false>::bytes(n, P.ld), st>>>(P, ws->counter + 1, nullptr);
      }
      else
        qp_solve_kernel<128, false, false><<<grid, 128, ccc::QpSm<128, false, false>::bytes(n, P.ld), st>>>(P, ws->counter, nullptr);
      break;
    case 1:
      qp_solve_kernel<128, true, false><<<grid, 128, ccc::QpSm<128, true, false>::bytes(n, P.ld), st>>>(P, ws->counter, ws->gmat);
      break;
    default:
      qp_solve_kernel<256, true, false><<<grid, 256, ccc::QpSm<256, true, false>::bytes(n, P.ld), st>>>(P, ws->counter, ws->gmat);
      break;
  }
  ws->launches++;
  if(!check(cudaGetLastError(), "launch qp_solve_kernel")) return CCC_ERR_CUDA;
  return CCC_OK;
}

extern "C" {

int32_t ccc_qp_solve(ccc_qp_ws_t * ws, const ccc_qp_batch_t * bt, ccc_qp_result_t * res, int32_t mem, void * stream_v)
{
  return ccc_qp_solve_grouped(ws, bt, 1, nullptr, res, mem, stream_v);
}

ccc_qp_ws_t * ccc_qp_create_grouped(int32_t n, int32_t n_eq, int32_t n_ineq, int32_t max_batch, int32_t max_groups)
{
  return ccc_host::qp_ws_create(n, n_eq, n_ineq, max_batch, max_groups, true);
}

int32_t ccc_qp_solve_grouped(ccc_qp_ws_t * ws, const ccc_qp_batch_t * bt, int32_t n_groups, const int32_t * group_id, ccc_qp_result_t * res,
                             int32_t mem, void * stream_v)
{
  using ccc_host::check;
  if(!ws || !bt || !res) return ccc_host::fail(CCC_ERR_INVALID, "null argument");
  const int n = bt->n, me = bt->n_eq, mi = bt->n_ineq, B = bt->batch;
  const size_t Gn = n_groups;
  if(n_groups <= 0 || n_groups > ws->max_groups) return ccc_host::fail(CCC_ERR_ALLOC, "more matrix groups than the workspace holds");
  if(n_groups > 1 && !group_id) return ccc_host::fail(CCC_ERR_INVALID, "group_id is NULL");
  if(mem == CCC_MEM_HOST && group_id)
    for(int b = 0; b < B; b++)
      if(group_id[b] < 0 || group_id[b] >= n_groups) return ccc_host::fail(CCC_ERR_INVALID, "group_id out of range");
  if(n != ws->n || me != ws->me || mi != ws->mi) return ccc_host::fail(CCC_ERR_INVALID, "sizes differ from the workspace's");
  if(B <= 0 || B > ws->max_batch) return ccc_host::fail(CCC_ERR_ALLOC, "batch exceeds workspace");
  const bool reuse = bt->Q == nullptr; // Q == NULL: keep the matrices (and their factorisation) of the previous call
  if(reuse && !ws->have_setup) return ccc_host::fail(CCC_ERR_INVALID, "Q is NULL but this workspace holds no matrices yet");
  if(!bt->d || (me && !bt->b) || (!reuse && (!bt->C || (me && !bt->A)))) return ccc_host::fail(CCC_ERR_INVALID, "null input");
  cudaStream_t st = mem == CCC_MEM_HOST ? ws->own_stream : reinterpret_cast<cudaStream_t>(stream_v);
  ws->launches = 0;
  const double *Q = bt->Q, *A = bt->A, *C = bt->C;
#define CCC_H2D(dst, src, nbytes, stream) \
  if(!check(cudaMemcpyAsync(dst, src, (nbytes), cudaMemcpyHostToDevice, stream), "H2D")) return CCC_ERR_CUDA
#define CCC_D2H(dst, src, nbytes, stream) \
  if((dst) && !check(cudaMemcpyAsync(dst, src, (nbytes), cudaMemcpyDeviceToHost, stream), "D2H")) return CCC_ERR_CUDA
  if(mem == CCC_MEM_HOST && !reuse)
  {
    CCC_H2D(ws->d_Q, Q, sizeof(double) * Gn * n * n, st);
    if(me) CCC_H2D(ws->d_A, A, sizeof(double) * Gn * me * n, st);
    CCC_H2D(ws->d_C, C, sizeof(double) * mi * n, st);
    Q = ws->d_Q;
    A = ws->d_A;
    C = ws->d_C;
  }
  if(!reuse)
  {
    const int rc = ccc_host::qp_setup_launch(ws, n_groups, Q, A, C, st);
    if(rc != CCC_OK) return rc;
  }
  const int * grp = group_id;
  if(mem == CCC_MEM_HOST && group_id)
  {
    CCC_H2D(ws->d_grp, group_id, sizeof(int) * B, st);
    grp = ws->d_grp;
  }
  ccc::QpParams P = ccc_host::qp_params(ws, B, grp);
  if(mem != CCC_MEM_HOST)
  {
    P.c = bt->c;
    P.b = bt->b;
    P.d = bt->d;
    P.out_x = res->x;
    P.out_iters = res->iters;
    P.out_status = res->status;
    P.out_n_active = res->n_active;
    P.out_active = res->active;
    return ccc_host::qp_launch(ws, P, st);
  }
  // Host buffers: the batch goes through in chunks on three streams — H2D of chunk k + 1 and D2H of chunk k - 1 run
  // while chunk k is being solved (the per-problem vectors are 2.4 KB in, 1.2 KB out per QP at n = 100: for short
  // solves the copies would otherwise be as long as the kernel).
  const int chunk = B <= 16384 ? B : 32768;
  const int nchunk = (B + chunk - 1) / chunk;
  if(nchunk > ccc_qp_ws::kMaxChunks) return ccc_host::fail(CCC_ERR_ALLOC, "batch too large for the chunked host path");
  if(!check(cudaEventRecord(ws->ev_setup, st), "event")) return CCC_ERR_CUDA;
  if(!check(cudaStreamWaitEvent(ws->h2d_stream, ws->ev_setup, 0), "wait")) return CCC_ERR_CUDA; // orders after earlier calls
  for(int k = 0; k < nchunk; k++)
  {
    const size_t lo = (size_t)k * chunk;
    const size_t nb = (size_t)(B - lo < (size_t)chunk ? B - lo : chunk);
    cudaStream_t hs = nchunk > 1 ? ws->h2d_stream : st, ds = nchunk > 1 ? ws->d2h_stream : st;
    if(bt->c) CCC_H2D(ws->d_c + lo * n, bt->c + lo * n, sizeof(double) * nb * n, hs);
    if(me) CCC_H2D(ws->d_b + lo * me, bt->b + lo * me, sizeof(double) * nb * me, hs);
    CCC_H2D(ws->d_d + lo * mi, bt->d + lo * mi, sizeof(double) * nb * mi, hs);
    if(nchunk > 1)
    {
      if(!check(cudaEventRecord(ws->ev_h2d[k], hs), "event")) return CCC_ERR_CUDA;
      if(!check(cudaStreamWaitEvent(st, ws->ev_h2d[k], 0), "wait")) return CCC_ERR_CUDA;
    }
    ccc::QpParams Pk = P;
    Pk.B = (int)nb;
    Pk.grp = grp ? grp + lo : nullptr;
    Pk.c = bt->c ? ws->d_c + lo * n : nullptr;
    Pk.b = ws->d_b + lo * me;
    Pk.d = ws->d_d + lo * mi;
    Pk.out_x = res->x ? ws->d_x + lo * n : nullptr;
    Pk.out_iters = res->iters ? ws->d_iters + lo : nullptr;
    Pk.out_status = res->status ? ws->d_status + lo : nullptr;
    Pk.out_n_active = res->n_active ? ws->d_nact + lo : nullptr;
    Pk.out_active = res->active ? ws->d_active + lo * n : nullptr;
    const int rc = ccc_host::qp_launch(ws, Pk, st);
    if(rc != CCC_OK) return rc;
    if(nchunk > 1)
    {
      if(!check(cudaEventRecord(ws->ev_solved[k], st), "event")) return CCC_ERR_CUDA;
      if(!check(cudaStreamWaitEvent(ds, ws->ev_solved[k], 0), "wait")) return CCC_ERR_CUDA;
    }
    CCC_D2H(res->x ? res->x + lo * n : nullptr, ws->d_x + lo * n, sizeof(double) * nb * n, ds);
    CCC_D2H(res->iters ? res->iters + lo : nullptr, ws->d_iters + lo, sizeof(int) * nb, ds);
    CCC_D2H(res->status ? res->status + lo : nullptr, ws->d_status + lo, sizeof(int) * nb, ds);
    CCC_D2H(res->n_active ? res->n_active + lo : nullptr, ws->d_nact + lo, sizeof(int) * nb, ds);
    CCC_D2H(res->active ? res->active + lo * n : nullptr, ws->d_active + lo * n, sizeof(int) * nb * n, ds);
  }
#undef CCC_H2D
#undef CCC_D2H
  if(nchunk > 1 && !check(cudaStreamSynchronize(ws->d2h_stream), "cudaStreamSynchronize")) return CCC_ERR_CUDA;
  if(!check(cudaStreamSynchronize(st), "cudaStreamSynchronize")) return CCC_ERR_CUDA;
  return CCC_OK;
}

/* Tuning hook (not part of the stable ABI): 0 = always one CTA per SM with the full R (round-1 kernel), 1 = packed R,
 * two CTAs per SM, with the full-R fallback pass (default). */
void ccc_qp_set_packed(int32_t on)
{
  ccc_host::g_qp_packed() = on != 0;
}

int32_t ccc_qp_last_launches(const ccc_qp_ws_t * ws)
{
  return ws ? ws->launches : 0;
}

} // extern "C"
