// linear_mpc_xy.cu — kernels + C-ABI host code of ccc_linear_mpc_xy_* (include/ccc_b200.h): everything
// CCC::LinearMpcXY::planOnce does after sampling its callbacks, on the device, for B initial states over a sweep of
// S contact / reference schedules.
//
// Replaces (reference): Model::Model (src/LinearMpcXY.cpp:59-83), StateSpaceModel::calcDiscMatrix
// (include/CCC/StateSpaceModel.h:164-216; closed form of the nilpotent model),
// VariantSequentialExtension<6>::setup (include/CCC/VariantSequentialExtension.h:110-186), procOnce
// (src/LinearMpcXY.cpp:116-181).  Arithmetic and evaluation order: oracle/xy.hpp.
//
//   xy_condense_kernel   one CTA per schedule: Ad of every stage, A_seq (thread per column), B_seq (thread per
//                        input column, carried down the stages by dense 6 x 6 fma chains), eq_mat / eq_vec
//   xy_hessian_kernel    obj_mat = B_seq' W B_seq + w_force I on the FP64 tensor cores (mma.sync m8n8k4.f64): B_seq of
//                        one schedule staged in shared memory serves as both operands, a warp carries six 8 x 8
//                        tiles of one tile row so that the scaled A fragment is formed once per k-step
//   xy_gradient_kernel   one CTA per problem: resid = ref - A_seq x0, obj_vec = -B_seq' W resid, eq_vec, bounds
//   QP                   ccc_host::qp_setup_launch (one factorisation per schedule) + qp_launch (qp.cu)
#include "../../include/ccc_b200.h"
#include "qp_host.cuh"

namespace
{
constexpr int kXyThreads = 256;
constexpr int kXyMmax = 64;      // largest row stride of the ridge / vertex tables the staging buffers hold
constexpr int kTilesPerWarp = 6; // 8 x 8 output tiles a warp accumulates at once

__device__ __forceinline__ double chain6(const double * a, const double * x)
{
  double acc = 0.0;
#pragma unroll
  for(int k = 0; k < 6; k++) acc = fma(a[k], x[k], acc);
  return acc;
}

__global__ void __launch_bounds__(kXyThreads) xy_condense_kernel(int N, int n, int me, int mm, int rows_pad, double dt, double mass,
                                                                 const int * __restrict__ m, const double * __restrict__ ridge,
                                                                 const double * __restrict__ vertex, const double * __restrict__ com_z,
                                                                 const double * __restrict__ fz, double * __restrict__ A_seq,
                                                                 double * __restrict__ B_seq, double * __restrict__ Aeq,
                                                                 double * __restrict__ beq, int * __restrict__ err)
{
  extern __shared__ __align__(16) double sm[];
  double * Ad = sm;                                   // [N][36]
  int * off = reinterpret_cast<int *>(sm + N * 36);   // [N + 1] first column of each stage
  int * eidx = off + N + 1;                           // [N] equality row of each stage (-1: no contact)
  __shared__ int s_bad;
  const int s = blockIdx.x, tid = threadIdx.x, rows = 6 * N;
  const double h2 = (dt * dt) * 0.5, h3 = ((dt * dt) * dt) / 6.0;
  if(tid == 0)
  {
    int acc = 0, e = 0;
    for(int k = 0; k < N; k++)
    {
      const int mk = m[s * N + k];
      off[k] = acc;
      eidx[k] = mk > 0 ? e++ : -1;
      acc += mk > 0 ? mk : 0;
      if(mk < 0 || mk > mm) acc = n + 1;
    }
    off[N] = acc;
    s_bad = (acc != n || e != me) ? 1 : 0;
    if(s_bad) *err = 1;
  }
  __syncthreads();
  if(s_bad) return;
  for(int k = tid; k < N; k += kXyThreads)
  {
    double * A = Ad + k * 36;
    const double a = fz[s * N + k] / mass;
    for(int i = 0; i < 36; i++) A[i] = 0.0;
    for(int i = 0; i < 6; i++) A[i * 6 + i] = 1.0;
    A[0 * 6 + 1] = dt;
    A[2 * 6 + 3] = dt;
    A[4 * 6 + 2] = -(a * dt);
    A[5 * 6 + 0] = a * dt;
    A[4 * 6 + 3] = -(a * h2);
    A[5 * 6 + 1] = a * h2;
    if(eidx[k] >= 0) beq[s * me + eidx[k]] = fz[s * N + k];
  }
  __syncthreads();
  // A_seq: column c of block i is Ad_i times column c of block i - 1
  if(tid < 6)
  {
    double col[6];
    for(int r = 0; r < 6; r++) col[r] = Ad[r * 6 + tid];
    double * out = A_seq + (size_t)s * rows * 6;
    for(int i = 0; i < N; i++)
    {
      if(i > 0)
      {
        double nx[6];
        for(int r = 0; r < 6; r++) nx[r] = chain6(Ad + i * 36 + r * 6, col);
        for(int r = 0; r < 6; r++) col[r] = nx[r];
      }
      for(int r = 0; r < 6; r++) out[(i * 6 + r) * 6 + tid] = col[r];
    }
  }
  // B_seq and eq_mat: one thread per input column
  double * Bs = B_seq + (size_t)s * rows_pad * n;
  for(int c = tid; c < n; c += kXyThreads)
  {
    int i = 0;
    while(off[i + 1] <= c) i++;
    const int l = c - off[i];
    const double a = fz[s * N + i] / mass;
    const double ah3 = a * h3;
    const double cz = com_z[s * N + i];
    const double * rg = ridge + ((size_t)(s * N + i) * mm + l) * 3;
    const double * vt = vertex + ((size_t)(s * N + i) * mm + l) * 3;
    const double rx = rg[0], ry = rg[1], rz = rg[2];
    const double hz = vt[2] - cz;
    const double b4 = fma(vt[1], rz, -(hz * ry));
    const double b5 = fma(hz, rx, -(vt[0] * rz));
    double col[6] = {h2 * rx, dt * rx, h2 * ry, dt * ry, fma(dt, b4, (-ah3) * ry), fma(dt, b5, ah3 * rx)};
    for(int j = 0; j < i; j++)
      for(int r = 0; r < 6; r++) Bs[(size_t)(j * 6 + r) * n + c] = 0.0;
    for(int j = i; j < N; j++)
    {
      if(j > i)
      {
        double nx[6];
        for(int r = 0; r < 6; r++) nx[r] = chain6(Ad + j * 36 + r * 6, col);
        for(int r = 0; r < 6; r++) col[r] = nx[r];
      }
      for(int r = 0; r < 6; r++) Bs[(size_t)(j * 6 + r) * n + c] = col[r];
    }
    for(int r = rows; r < rows_pad; r++) Bs[(size_t)r * n + c] = 0.0;
    for(int e = 0; e < me; e++) Aeq[((size_t)s * me + e) * n + c] = e == eidx[i] ? rz : 0.0;
  }
}

/** Smallest leading dimension >= cols with ld % 16 == 4: the 4 x 4 doubles a half warp reads for an m8n8k4 fragment
 *  (4 consecutive columns of 4 consecutive rows) then fall into 16 different 8-byte bank slots. */
__host__ __device__ inline int xy_ld(int cols)
{
  return cols + ((4 - cols % 16) + 16) % 16;
}

__device__ __forceinline__ void dmma884(double & d0, double & d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

/** grid (S, parts): CTA (s, p) computes the (tile row, tile group) units u = p, p + parts, ... of schedule s.
 *  The rows of B_seq go through shared memory in chunks of `chunk` rows (a multiple of 4; one chunk at the
 *  reference's sizes); between chunks the accumulators rest in H itself, so the chain over r stays sequential. */
__global__ void __launch_bounds__(kXyThreads) xy_hessian_kernel(int n, int rows_pad, int chunk, double w0, double w1, double w2, double w3,
                                                                double w4, double w5, double w_force,
                                                                const double * __restrict__ B_seq, double * __restrict__ H)
{
  extern __shared__ __align__(16) double sm[];
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = (n + 7) / 8, ld = xy_ld(T * 8);
  const int groups = (T + kTilesPerWarp - 1) / kTilesPerWarp;
  const int units = T * groups;
  const double w6[6] = {w0, w1, w2, w3, w4, w5};
  const double * Bg = B_seq + (size_t)s * rows_pad * n;
  double * Hs = H + (size_t)s * n * n;
  double * wrow = sm + (size_t)chunk * ld; // [chunk] output weight of each staged row
  const int g = lane >> 2, t = lane & 3;
  for(int r0 = 0; r0 < rows_pad; r0 += chunk)
  {
    const int nr = rows_pad - r0 < chunk ? rows_pad - r0 : chunk;
    __syncthreads();
    for(int e = tid; e < nr * ld; e += kXyThreads)
    {
      const int r = e / ld, c = e % ld;
      sm[e] = c < n ? Bg[(size_t)(r0 + r) * n + c] : 0.0;
    }
    for(int r = tid; r < nr; r += kXyThreads)
    {
      const int k = (r0 + r) % 6;
      wrow[r] = k == 0 ? w6[0] : k == 1 ? w6[1] : k == 2 ? w6[2] : k == 3 ? w6[3] : k == 4 ? w6[4] : w6[5];
    }
    __syncthreads();
    const bool first = r0 == 0, last = r0 + chunk >= rows_pad;
    for(int u = blockIdx.y * (kXyThreads / 32) + warp; u < units; u += gridDim.y * (kXyThreads / 32))
    {
      const int ti = u / groups, tj0 = (u % groups) * kTilesPerWarp;
      const int i0 = ti * 8;
      double acc[kTilesPerWarp][2];
#pragma unroll
      for(int q = 0; q < kTilesPerWarp; q++)
      {
        const int row = i0 + g, col = (tj0 + q) * 8 + 2 * t;
        acc[q][0] = (!first && tj0 + q < T && row < n && col < n) ? Hs[(size_t)row * n + col] : 0.0;
        acc[q][1] = (!first && tj0 + q < T && row < n && col + 1 < n) ? Hs[(size_t)row * n + col + 1] : 0.0;
      }
      for(int k0 = 0; k0 < nr; k0 += 4)
      {
        const double * rowp = sm + (size_t)(k0 + t) * ld;
        const double a = wrow[k0 + t] * rowp[i0 + g];
#pragma unroll
        for(int q = 0; q < kTilesPerWarp; q++)
          if(tj0 + q < T) dmma884(acc[q][0], acc[q][1], a, rowp[(tj0 + q) * 8 + g]);
      }
#pragma unroll
      for(int q = 0; q < kTilesPerWarp; q++)
      {
        if(tj0 + q >= T) continue;
        const int row = i0 + g, col = (tj0 + q) * 8 + 2 * t;
        if(row < n && col < n) Hs[(size_t)row * n + col] = (last && row == col) ? acc[q][0] + w_force : acc[q][0];
        if(row < n && col + 1 < n) Hs[(size_t)row * n + col + 1] = (last && row == col + 1) ? acc[q][1] + w_force : acc[q][1];
      }
    }
  }
}

__global__ void __launch_bounds__(kXyThreads) xy_gradient_kernel(int N, int S, int n, int me, int rows_pad, double w0, double w1, double w2,
                                                                 double w3, double w4, double w5, double lo, double hi,
                                                                 const int * __restrict__ sched_id, const double * __restrict__ x0,
                                                                 const double * __restrict__ ref, const double * __restrict__ A_seq,
                                                                 const double * __restrict__ B_seq, const double * __restrict__ beq,
                                                                 double * __restrict__ gvec, double * __restrict__ bq,
                                                                 double * __restrict__ dvec, int * __restrict__ err)
{
  extern __shared__ __align__(16) double resid[]; // [6N]
  const int b = blockIdx.x, tid = threadIdx.x, rows = 6 * N;
  const int s = sched_id[b];
  if(s < 0 || s >= S)
  {
    if(tid == 0) *err = 2;
    return;
  }
  const double w6[6] = {w0, w1, w2, w3, w4, w5};
  double x[6];
  for(int c = 0; c < 6; c++) x[c] = x0[(size_t)b * 6 + c];
  for(int r = tid; r < rows; r += kXyThreads) resid[r] = ref[(size_t)s * rows + r] - chain6(A_seq + ((size_t)s * rows + r) * 6, x);
  __syncthreads();
  const double * Bs = B_seq + (size_t)s * rows_pad * n;
  for(int j = tid; j < n; j += kXyThreads)
  {
    double acc = 0.0;
    for(int r = 0; r < rows; r += 6)
    {
#pragma unroll
      for(int k = 0; k < 6; k++) acc = fma(w6[k] * Bs[(size_t)(r + k) * n + j], resid[r + k], acc);
    }
    gvec[(size_t)b * n + j] = -acc;
    dvec[(size_t)b * 2 * n + j] = -lo;
    dvec[(size_t)b * 2 * n + n + j] = hi;
  }
  for(int e = tid; e < me; e += kXyThreads) bq[(size_t)b * me + e] = beq[(size_t)s * me + e];
}

__global__ void xy_box_kernel(int n, double * __restrict__ C)
{
  for(size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)2 * n * n; e += (size_t)gridDim.x * blockDim.x)
  {
    const int i = (int)(e / n), j = (int)(e % n);
    C[e] = i < n ? (i == j ? -1.0 : 0.0) : (i - n == j ? 1.0 : 0.0);
  }
}

template<class T>
bool dev_alloc(T *& p, size_t n)
{
  return ccc_host::check(cudaMalloc(reinterpret_cast<void **>(&p), (n ? n : 1) * sizeof(T)), "cudaMalloc");
}

constexpr size_t kHessianSmem = 200 * 1024;
} // namespace

struct ccc_linear_mpc_xy_ws
{
  int N = 0, n = 0, me = 0, max_batch = 0, max_sched = 0, rows = 0, rows_pad = 0, launches = 0, n_sm = 148;
  ccc_qp_ws * qp = nullptr;
  cudaStream_t own_stream = nullptr;
  // staging of host-buffer calls
  int *d_m = nullptr, *d_sid = nullptr;
  double *d_ridge = nullptr, *d_vertex = nullptr, *d_comz = nullptr, *d_fz = nullptr, *d_ref = nullptr, *d_x0 = nullptr;
  double * d_u = nullptr;
  int *d_iters = nullptr, *d_status = nullptr, *d_nact = nullptr, *d_active = nullptr;
  // device-resident intermediates
  double *A_seq = nullptr, *B_seq = nullptr, *H = nullptr, *Aeq = nullptr, *beq = nullptr, *Cbox = nullptr;
  double *g = nullptr, *bq = nullptr, *dvec = nullptr;
  int * err = nullptr;
};

extern "C" {

ccc_linear_mpc_xy_ws_t * ccc_linear_mpc_xy_create(int32_t horizon_steps, int32_t n, int32_t n_eq, int32_t max_batch, int32_t max_sched)
{
  using ccc_host::check;
  if(horizon_steps <= 0 || horizon_steps > 512 || n <= 0 || n > 256 || n_eq < 0 || n_eq > horizon_steps || n_eq > n || max_batch <= 0
     || max_sched <= 0)
  {
    ccc_host::set_error("ccc_linear_mpc_xy_create: sizes outside the kernels' limits (N <= 512, n <= 256, n_eq <= min(N, n))");
    return nullptr;
  }
  auto * ws = new ccc_linear_mpc_xy_ws();
  ws->N = horizon_steps;
  ws->n = n;
  ws->me = n_eq;
  ws->max_batch = max_batch;
  ws->max_sched = max_sched;
  ws->rows = 6 * horizon_steps;
  ws->rows_pad = (ws->rows + 3) & ~3;
  ws->qp = ccc_host::qp_ws_create(n, n_eq, 2 * n, max_batch, max_sched, false);
  if(!ws->qp)
  {
    delete ws;
    return nullptr;
  }
  ws->n_sm = ws->qp->n_sm;
  const size_t N = horizon_steps, S = max_sched, B = max_batch, nn = n, ME = n_eq;
  bool ok = check(cudaStreamCreateWithFlags(&ws->own_stream, cudaStreamNonBlocking), "cudaStreamCreate");
  ok = ok && dev_alloc(ws->d_m, S * N) && dev_alloc(ws->d_sid, B) && dev_alloc(ws->d_ridge, S * N * kXyMmax * 3)
       && dev_alloc(ws->d_vertex, S * N * kXyMmax * 3) && dev_alloc(ws->d_comz, S * N) && dev_alloc(ws->d_fz, S * N)
       && dev_alloc(ws->d_ref, S * N * 6) && dev_alloc(ws->d_x0, B * 6);
  ok = ok && dev_alloc(ws->d_u, B * nn) && dev_alloc(ws->d_iters, B) && dev_alloc(ws->d_status, B) && dev_alloc(ws->d_nact, B)
       && dev_alloc(ws->d_active, B * nn);
  ok = ok && dev_alloc(ws->A_seq, S * ws->rows * 6) && dev_alloc(ws->B_seq, S * ws->rows_pad * nn) && dev_alloc(ws->H, S * nn * nn)
       && dev_alloc(ws->Aeq, S * ME * nn) && dev_alloc(ws->beq, S * ME) && dev_alloc(ws->Cbox, 2 * nn * nn);
  ok = ok && dev_alloc(ws->g, B * nn) && dev_alloc(ws->bq, B * ME) && dev_alloc(ws->dvec, B * 2 * nn) && dev_alloc(ws->err, 1);
  ok = ok && check(cudaFuncSetAttribute(xy_hessian_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHessianSmem), "cudaFuncSetAttribute");
  if(ok)
  {
    xy_box_kernel<<<64, 256, 0, ws->own_stream>>>(n, ws->Cbox);
    ok = check(cudaGetLastError(), "launch xy_box_kernel") && check(cudaStreamSynchronize(ws->own_stream), "cudaStreamSynchronize");
  }
  if(!ok)
  {
    ccc_linear_mpc_xy_destroy(ws);
    return nullptr;
  }
  return ws;
}

void ccc_linear_mpc_xy_destroy(ccc_linear_mpc_xy_ws_t * ws)
{
  if(!ws) return;
  void * ptrs[] = {ws->d_m,   ws->d_sid,    ws->d_ridge, ws->d_vertex, ws->d_comz, ws->d_fz, ws->d_ref, ws->d_x0, ws->d_u, ws->d_iters,
                   ws->d_status, ws->d_nact, ws->d_active, ws->A_seq,  ws->B_seq,  ws->H,    ws->Aeq,   ws->beq,  ws->Cbox, ws->g,
                   ws->bq,    ws->dvec,     ws->err};
  for(void * p : ptrs)
    if(p) cudaFree(p);
  if(ws->own_stream) cudaStreamDestroy(ws->own_stream);
  ccc_qp_destroy(ws->qp);
  delete ws;
}

int32_t ccc_linear_mpc_xy_solve(ccc_linear_mpc_xy_ws_t * ws,
                                const ccc_linear_mpc_xy_batch_t * bt,
                                ccc_linear_mpc_xy_result_t * res,
                                int32_t mem,
                                void * stream_v)
{
  using ccc_host::check;
  if(!ws || !bt || !res) return ccc_host::fail(CCC_ERR_INVALID, "null argument");
  const int N = bt->horizon_steps, B = bt->batch, S = bt->n_sched, mm = bt->m_max, n = ws->n, me = ws->me;
  if(N != ws->N) return ccc_host::fail(CCC_ERR_INVALID, "horizon_steps differs from the workspace's");
  if(B <= 0 || B > ws->max_batch || S <= 0 || S > ws->max_sched) return ccc_host::fail(CCC_ERR_ALLOC, "batch / n_sched exceeds workspace");
  if(mm <= 0 || mm > kXyMmax) return ccc_host::fail(CCC_ERR_INVALID, "m_max outside 1..64");
  if(!bt->sched_id || !bt->m || !bt->ridge || !bt->vertex || !bt->com_z || !bt->total_force_z || !bt->ref_output || !bt->x0)
    return ccc_host::fail(CCC_ERR_INVALID, "null input");
  const bool host = mem == CCC_MEM_HOST;
  cudaStream_t st = host ? ws->own_stream : reinterpret_cast<cudaStream_t>(stream_v);
  ws->launches = 0;
  ws->qp->launches = 0;
  const int *m = bt->m, *sid = bt->sched_id;
  const double *ridge = bt->ridge, *vertex = bt->vertex, *comz = bt->com_z, *fz = bt->total_force_z, *ref = bt->ref_output, *x0 = bt->x0;
#define CCC_H2D(dst, src, nbytes) \
  if(!check(cudaMemcpyAsync(dst, src, (nbytes), cudaMemcpyHostToDevice, st), "H2D")) return CCC_ERR_CUDA
#define CCC_OUT(dst, src, nbytes) \
  if((dst) && !check(cudaMemcpyAsync(dst, src, (nbytes), host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st), "copy out")) return CCC_ERR_CUDA
  if(host)
  {
    const size_t SN = (size_t)S * N;
    CCC_H2D(ws->d_m, m, sizeof(int) * SN);
    CCC_H2D(ws->d_sid, sid, sizeof(int) * B);
    CCC_H2D(ws->d_ridge, ridge, sizeof(double) * SN * mm * 3);
    CCC_H2D(ws->d_vertex, vertex, sizeof(double) * SN * mm * 3);
    CCC_H2D(ws->d_comz, comz, sizeof(double) * SN);
    CCC_H2D(ws->d_fz, fz, sizeof(double) * SN);
    CCC_H2D(ws->d_ref, ref, sizeof(double) * SN * 6);
    CCC_H2D(ws->d_x0, x0, sizeof(double) * B * 6);
    m = ws->d_m;
    sid = ws->d_sid;
    ridge = ws->d_ridge;
    vertex = ws->d_vertex;
    comz = ws->d_comz;
    fz = ws->d_fz;
    ref = ws->d_ref;
    x0 = ws->d_x0;
  }
  if(!check(cudaMemsetAsync(ws->err, 0, sizeof(int), st), "memset")) return CCC_ERR_CUDA;
  const size_t smem_c = sizeof(double) * N * 36 + sizeof(int) * (2 * N + 2);
  xy_condense_kernel<<<S, kXyThreads, smem_c, st>>>(N, n, me, mm, ws->rows_pad, bt->dt, bt->mass, m, ridge, vertex, comz, fz, ws->A_seq, ws->B_seq,
                                                   ws->Aeq, ws->beq, ws->err);
  const int ld = xy_ld(((n + 7) / 8) * 8);
  int chunk = (int)((kHessianSmem / sizeof(double) - 8) / (ld + 1)) & ~3;
  if(chunk > ws->rows_pad) chunk = ws->rows_pad;
  int parts = (ws->n_sm + S - 1) / S;
  parts = parts < 1 ? 1 : parts > 8 ? 8 : parts;
  const double * w = bt->w_output;
  xy_hessian_kernel<<<dim3(S, parts), kXyThreads, sizeof(double) * ((size_t)chunk * ld + chunk), st>>>(n, ws->rows_pad, chunk, w[0], w[1], w[2],
                                                                                                      w[3], w[4], w[5], bt->w_force, ws->B_seq,
                                                                                                      ws->H);
  xy_gradient_kernel<<<B, kXyThreads, sizeof(double) * 6 * N, st>>>(N, S, n, me, ws->rows_pad, w[0], w[1], w[2], w[3], w[4], w[5], bt->force_lo,
                                                                   bt->force_hi, sid, x0, ref, ws->A_seq, ws->B_seq, ws->beq, ws->g, ws->bq,
                                                                   ws->dvec, ws->err);
  ws->launches += 3;
  if(!check(cudaGetLastError(), "launch xy kernels")) return CCC_ERR_CUDA;
  int rc = ccc_host::qp_setup_launch(ws->qp, S, ws->H, ws->Aeq, ws->Cbox, st);
  if(rc != CCC_OK) return rc;
  ccc::QpParams P = ccc_host::qp_params(ws->qp, B, sid);
  P.c = ws->g;
  P.b = ws->bq;
  P.d = ws->dvec;
  P.out_x = host ? (res->u ? ws->d_u : nullptr) : res->u;
  P.out_iters = host ? (res->iters ? ws->d_iters : nullptr) : res->iters;
  P.out_status = host ? ws->d_status : res->status;
  P.out_n_active = host ? (res->n_active ? ws->d_nact : nullptr) : res->n_active;
  P.out_active = host ? (res->active ? ws->d_active : nullptr) : res->active;
  rc = ccc_host::qp_launch(ws->qp, P, st);
  if(rc != CCC_OK) return rc;
  ws->launches += ws->qp->launches;
  // optional intermediates (B_seq is stored with its row count padded to a multiple of 4)
  CCC_OUT(res->A_seq, ws->A_seq, sizeof(double) * S * ws->rows * 6);
  if(res->B_seq
     && !check(cudaMemcpy2DAsync(res->B_seq, sizeof(double) * ws->rows * n, ws->B_seq, sizeof(double) * ws->rows_pad * n,
                                 sizeof(double) * ws->rows * n, S, host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st),
               "copy out"))
    return CCC_ERR_CUDA;
  CCC_OUT(res->obj_mat, ws->H, sizeof(double) * S * n * n);
  CCC_OUT(res->obj_vec, ws->g, sizeof(double) * B * n);
  if(!host) return CCC_OK; // schedule shape / sched_id errors of a device-pointer call surface as status 3 / garbage: unchecked
  CCC_OUT(res->u, ws->d_u, sizeof(double) * B * n);
  CCC_OUT(res->iters, ws->d_iters, sizeof(int) * B);
  CCC_OUT(res->status, ws->d_status, sizeof(int) * B);
  CCC_OUT(res->n_active, ws->d_nact, sizeof(int) * B);
  CCC_OUT(res->active, ws->d_active, sizeof(int) * B * n);
  int err = 0;
  if(!check(cudaMemcpyAsync(&err, ws->err, sizeof(int), cudaMemcpyDeviceToHost, st), "D2H")) return CCC_ERR_CUDA;
  if(!check(cudaStreamSynchronize(st), "cudaStreamSynchronize")) return CCC_ERR_CUDA;
#undef CCC_H2D
#undef CCC_OUT
  if(err == 1) return ccc_host::fail(CCC_ERR_INVALID, "a schedule's total input dimension or number of contact stages differs from the workspace's (n, n_eq)");
  if(err == 2) return ccc_host::fail(CCC_ERR_INVALID, "sched_id out of range");
  return CCC_OK;
}

int32_t ccc_linear_mpc_xy_last_launches(const ccc_linear_mpc_xy_ws_t * ws)
{
  return ws ? ws->launches : 0;
}

} // extern "C"
