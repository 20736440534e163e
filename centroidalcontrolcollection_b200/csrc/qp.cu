// qp.cu — kernels + C-ABI host code of ccc_qp_* (include/ccc_b200.h): batched strictly convex dense
// QP with shared matrices, one CTA per problem (qp_cta_core.cuh), persistent CTAs pulling problems
// from an atomic counter.
#include "../../include/ccc_b200.h"
#include "common_host.cuh"
#include "qp_cta_core.cuh"

namespace
{
__global__ void __launch_bounds__(ccc::kQpThreads, 1) qp_setup_kernel(int n, int me, int mi, const double * Q, const double * A,
                                                                       const double * C, double * Lg, double * invd, double * J0,
                                                                       double * At, double * Ct, int * ok_flag, double * J0s)
{
  ccc::qp_setup_cta(n, me, mi, Q, A, C, Lg, invd, J0, At, Ct, ok_flag, J0s);
}

template<int NT, bool kGlobal>
__global__ void __launch_bounds__(NT, 1) qp_solve_kernel(const __grid_constant__ ccc::QpParams P, int * __restrict__ counter,
                                                         double * __restrict__ gmat)
{
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_b;
  __shared__ __align__(8) unsigned long long s_mbar; // completion barrier of the per-problem TMA load of J
  double * slab = kGlobal ? gmat + (size_t)blockIdx.x * 2 * P.n * P.ld : nullptr;
  if(!kGlobal && threadIdx.x == 0) ccc::mbar_init(&s_mbar, 1);
  __syncthreads();
  unsigned mbar_phase = 0;
  for(;;)
  {
    if(threadIdx.x == 0) s_b = atomicAdd(counter, 1);
    __syncthreads();
    const int b = s_b;
    __syncthreads();
    if(b >= P.B) break;
    ccc::QpCta<NT, kGlobal> cta(P, smem, b, slab);
    cta.mbar = kGlobal ? nullptr : &s_mbar;
    cta.mbar_phase = mbar_phase;
    cta.solve();
    mbar_phase = cta.mbar_phase;
  }
}

constexpr size_t kSmemLimit = 227 * 1024;

/** 0: 128 threads, J/R in shared memory; 1: 128 threads, J/R in global memory; 2: 256 threads, global. */
int qp_shape(int n)
{
  if(n > 128) return 2;
  return ccc::QpSm<128, false>::bytes(n, n | 1) <= kSmemLimit ? 0 : 1;
}

template<class T>
bool dev_alloc(T *& p, size_t n)
{
  return ccc_host::check(cudaMalloc(reinterpret_cast<void **>(&p), (n ? n : 1) * sizeof(T)), "cudaMalloc");
}
} // namespace

struct ccc_qp_ws
{
  int n = 0, me = 0, mi = 0, max_batch = 0, device = 0, launches = 0;
  bool have_setup = false; // the matrices of an earlier call are factorised and resident (reused when Q == NULL)
  double *Lg = nullptr, *invd = nullptr, *J0 = nullptr, *J0s = nullptr, *At = nullptr, *Ct = nullptr;
  int *ok_flag = nullptr, *counter = nullptr;
  double * gmat = nullptr; // per-CTA J/R slabs when they do not fit in shared memory
  int n_sm = 148;
  // staging for CCC_MEM_HOST
  double *d_Q = nullptr, *d_A = nullptr, *d_C = nullptr, *d_c = nullptr, *d_b = nullptr, *d_d = nullptr, *d_x = nullptr;
  int *d_iters = nullptr, *d_status = nullptr, *d_nact = nullptr, *d_active = nullptr;
  cudaStream_t own_stream = nullptr;
};

extern "C" {

ccc_qp_ws_t * ccc_qp_create(int32_t n, int32_t n_eq, int32_t n_ineq, int32_t max_batch)
{
  const int nt_threads = n > 128 ? 256 : 128;
  if(n <= 0 || n > 256 || n_eq < 0 || n_eq > n || n_ineq <= 0 || n_eq + n_ineq > 4 * nt_threads || max_batch <= 0)
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
  cudaGetDevice(&ws->device);
  const size_t N = n, ME = n_eq, MI = n_ineq, B = max_batch;
  bool ok = true;
  ok = ok && dev_alloc(ws->Lg, N * N) && dev_alloc(ws->invd, N) && dev_alloc(ws->J0, N * N) && dev_alloc(ws->J0s, N * (N | 1));
  ok = ok && dev_alloc(ws->At, N * ME) && dev_alloc(ws->Ct, N * MI) && dev_alloc(ws->ok_flag, 2) && dev_alloc(ws->counter, 1);
  ok = ok && dev_alloc(ws->d_Q, N * N) && dev_alloc(ws->d_A, N * ME) && dev_alloc(ws->d_C, N * MI);
  ok = ok && dev_alloc(ws->d_c, B * N) && dev_alloc(ws->d_b, B * ME) && dev_alloc(ws->d_d, B * MI) && dev_alloc(ws->d_x, B * N);
  ok = ok && dev_alloc(ws->d_iters, B) && dev_alloc(ws->d_status, B) && dev_alloc(ws->d_nact, B) && dev_alloc(ws->d_active, B * N);
  ok = ok && ccc_host::check(cudaStreamCreateWithFlags(&ws->own_stream, cudaStreamNonBlocking), "cudaStreamCreate");
  const int ld = n | 1;
  cudaDeviceGetAttribute(&ws->n_sm, cudaDevAttrMultiProcessorCount, ws->device);
  const int shape = qp_shape(n);
  if(shape == 0)
    ok = ok
         && ccc_host::check(cudaFuncSetAttribute(qp_solve_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)ccc::QpSm<128, false>::bytes(n, ld)),
                            "cudaFuncSetAttribute(smem)");
  else
    ok = ok && dev_alloc(ws->gmat, (size_t)ws->n_sm * 2 * N * ld);
  if(!ok)
  {
    ccc_qp_destroy(ws);
    return nullptr;
  }
  return ws;
}

void ccc_qp_destroy(ccc_qp_ws_t * ws)
{
  if(!ws) return;
  void * ptrs[] = {ws->J0s, ws->gmat, ws->Lg,  ws->invd, ws->J0,  ws->At,  ws->Ct,     ws->ok_flag, ws->counter, ws->d_Q,    ws->d_A,
                   ws->d_C, ws->d_c,  ws->d_b, ws->d_d, ws->d_x,    ws->d_iters, ws->d_status, ws->d_nact, ws->d_active};
  for(void * p : ptrs)
    if(p) cudaFree(p);
  if(ws->own_stream) cudaStreamDestroy(ws->own_stream);
  delete ws;
}

int32_t ccc_qp_solve(ccc_qp_ws_t * ws, const ccc_qp_batch_t * bt, ccc_qp_result_t * res, int32_t mem, void * stream_v)
{
  using ccc_host::check;
  if(!ws || !bt || !res) return ccc_host::fail(CCC_ERR_INVALID, "null argument");
  const int n = bt->n, me = bt->n_eq, mi = bt->n_ineq, B = bt->batch;
  if(n != ws->n || me != ws->me || mi != ws->mi) return ccc_host::fail(CCC_ERR_INVALID, "sizes differ from the workspace's");
  if(B <= 0 || B > ws->max_batch) return ccc_host::fail(CCC_ERR_ALLOC, "batch exceeds workspace");
  const bool reuse = bt->Q == nullptr; // Q == NULL: keep the matrices (and their factorisation) of the previous call
  if(reuse && !ws->have_setup) return ccc_host::fail(CCC_ERR_INVALID, "Q is NULL but this workspace holds no matrices yet");
  if(!bt->d || (me && !bt->b) || (!reuse && (!bt->C || (me && !bt->A)))) return ccc_host::fail(CCC_ERR_INVALID, "null input");
  cudaStream_t st = mem == CCC_MEM_HOST ? ws->own_stream : reinterpret_cast<cudaStream_t>(stream_v);
  ws->launches = 0;
  const double *Q = bt->Q, *A = bt->A, *C = bt->C, *c = bt->c, *b = bt->b, *d = bt->d;
  double * o_x = res->x;
  int *o_iters = res->iters, *o_status = res->status, *o_nact = res->n_active, *o_active = res->active;
  if(mem == CCC_MEM_HOST)
  {
#define CCC_H2D(dst, src, nbytes) \
  if(!check(cudaMemcpyAsync(dst, src, (nbytes), cudaMemcpyHostToDevice, st), "H2D")) return CCC_ERR_CUDA
    if(!reuse)
    {
      CCC_H2D(ws->d_Q, Q, sizeof(double) * n * n);
      if(me) CCC_H2D(ws->d_A, A, sizeof(double) * me * n);
      CCC_H2D(ws->d_C, C, sizeof(double) * mi * n);
    }
    if(c) CCC_H2D(ws->d_c, c, sizeof(double) * B * n);
    if(me) CCC_H2D(ws->d_b, b, sizeof(double) * B * me);
    CCC_H2D(ws->d_d, d, sizeof(double) * B * mi);
#undef CCC_H2D
    Q = ws->d_Q;
    A = ws->d_A;
    C = ws->d_C;
    if(c) c = ws->d_c;
    b = ws->d_b;
    d = ws->d_d;
    o_x = res->x ? ws->d_x : nullptr;
    o_iters = res->iters ? ws->d_iters : nullptr;
    o_status = res->status ? ws->d_status : nullptr;
    o_nact = res->n_active ? ws->d_nact : nullptr;
    o_active = res->active ? ws->d_active : nullptr;
  }
  if(!reuse)
  {
    qp_setup_kernel<<<1, ccc::kQpThreads, 0, st>>>(n, me, mi, Q, A, C, ws->Lg, ws->invd, ws->J0, ws->At, ws->Ct, ws->ok_flag, ws->J0s);
    ws->launches++;
    ws->have_setup = true;
  }
  ccc::QpParams P;
  P.n = n;
  P.me = me;
  P.mi = mi;
  P.B = B;
  P.ld = n | 1;
  P.J0 = ws->J0;
  P.J0s = ws->J0s;
  P.At = ws->At;
  P.Ct = ws->Ct;
  P.c = c;
  P.b = b;
  P.d = d;
  P.setup_ok = ws->ok_flag;
  P.max_iter = 1000;
  P.viol_tol = 1e-10;
  P.out_x = o_x;
  P.out_iters = o_iters;
  P.out_status = o_status;
  P.out_n_active = o_nact;
  P.out_active = o_active;
  if(!check(cudaMemsetAsync(ws->counter, 0, sizeof(int), st), "memset")) return CCC_ERR_CUDA;
  const int grid = B < ws->n_sm ? B : ws->n_sm;
  switch(qp_shape(n))
  {
    case 0:
      qp_solve_kernel<128, false><<<grid, 128, ccc::QpSm<128, false>::bytes(n, P.ld), st>>>(P, ws->counter, nullptr);
      break;
    case 1:
      qp_solve_kernel<128, true><<<grid, 128, ccc::QpSm<128, true>::bytes(n, P.ld), st>>>(P, ws->counter, ws->gmat);
      break;
    default:
      qp_solve_kernel<256, true><<<grid, 256, ccc::QpSm<256, true>::bytes(n, P.ld), st>>>(P, ws->counter, ws->gmat);
      break;
  }
  ws->launches++;
  if(!check(cudaGetLastError(), "launch qp_solve_kernel")) return CCC_ERR_CUDA;
  if(mem == CCC_MEM_HOST)
  {
#define CCC_D2H(dst, src, nbytes) \
  if((dst) && !check(cudaMemcpyAsync(dst, src, (nbytes), cudaMemcpyDeviceToHost, st), "D2H")) return CCC_ERR_CUDA
    CCC_D2H(res->x, ws->d_x, sizeof(double) * B * n);
    CCC_D2H(res->iters, ws->d_iters, sizeof(int) * B);
    CCC_D2H(res->status, ws->d_status, sizeof(int) * B);
    CCC_D2H(res->n_active, ws->d_nact, sizeof(int) * B);
    CCC_D2H(res->active, ws->d_active, sizeof(int) * B * n);
#undef CCC_D2H
    if(!check(cudaStreamSynchronize(st), "cudaStreamSynchronize")) return CCC_ERR_CUDA;
  }
  return CCC_OK;
}

int32_t ccc_qp_last_launches(const ccc_qp_ws_t * ws)
{
  return ws ? ws->launches : 0;
}

} // extern "C"
