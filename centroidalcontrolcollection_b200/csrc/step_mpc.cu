// step_mpc.cu — ccc_step_mpc_plan (include/ccc_b200.h): CCC::StepMpc::planOnce for a batch, one thread per problem and axis.
// Replaces: StepMpc1d::planOnce (reference src/StepMpc.cpp:27-193) x 2 axes (:195-247).  The arithmetic is step_mpc_core.cuh.
#include "../../include/ccc_b200.h"
#include "common_host.cuh"
#include "step_mpc_core.cuh"

namespace
{
__global__ void step_mpc_kernel(ccc_step_mpc_batch_t in, ccc_step_mpc_result_t out)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if(idx >= 2 * in.batch) return;
  const int b = idx >> 1, a = idx & 1;
  const int p = in.plan_id[b], K = in.max_elements;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  double cur = nan, nxt = 0.0;
  bool has = false;
  if(p >= 0 && p < in.n_plans)
  {
    const int n = in.n_elements[p];
    if(n >= 1 && n <= K && K <= ccc_step::kStepMpcMax)
    {
      const ccc_step::Weights w = {in.w_free_zmp, in.w_fixed_zmp, in.w_double_support, in.w_pos, in.w_vel, in.w_capture_point_abs,
                                   in.w_capture_point_rel};
      if(!ccc_step::plan_1d(n, in.single + (size_t)p * K, in.zmp + (size_t)p * K * 2 + a, 2, in.end_time + (size_t)p * K, in.current_time[p],
                            in.com_height, w, in.x_pos[idx], in.x_vel[idx], cur, nxt, has))
        cur = nan;
    }
  }
  out.current_zmp[idx] = cur;
  if(out.next_foot_zmp) out.next_foot_zmp[idx] = nxt;
  if(out.has_next && a == 0) out.has_next[b] = has ? 1 : 0; // the phase pattern decides it, the same for both axes
}
} // namespace

extern "C" int32_t ccc_step_mpc_plan(const ccc_step_mpc_batch_t * bt, const ccc_step_mpc_result_t * res, int32_t mem, void * stream)
{
  using ccc_host::check;
  if(!bt || !res || !res->current_zmp) return ccc_host::fail(CCC_ERR_INVALID, "null argument");
  const int B = bt->batch, P = bt->n_plans, K = bt->max_elements;
  if(B <= 0 || P <= 0 || K < 1 || K > CCC_STEP_MPC_MAX_ELEMENTS || !(bt->com_height > 0))
    return ccc_host::fail(CCC_ERR_INVALID, "ccc_step_mpc_plan: sizes / com_height");
  if(!bt->plan_id || !bt->x_pos || !bt->x_vel || !bt->current_time || !bt->n_elements || !bt->single || !bt->zmp || !bt->end_time)
    return ccc_host::fail(CCC_ERR_INVALID, "null input");
  if(mem != CCC_MEM_HOST)
  {
    step_mpc_kernel<<<(2 * B + 63) / 64, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*bt, *res);
    return check(cudaGetLastError(), "launch step_mpc_kernel") ? CCC_OK : CCC_ERR_CUDA;
  }
  for(int b = 0; b < B; b++)
    if(bt->plan_id[b] < 0 || bt->plan_id[b] >= P) return ccc_host::fail(CCC_ERR_INVALID, "plan_id out of range");
  for(int p = 0; p < P; p++)
    if(bt->n_elements[p] < 1 || bt->n_elements[p] > K) return ccc_host::fail(CCC_ERR_INVALID, "n_elements outside [1, max_elements]");
  int ndev = 0;
  if(!check(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount") || ndev == 0)
    return ccc_host::fail(CCC_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
  const size_t nb_d = sizeof(double) * ((size_t)B * 8 + (size_t)P * (1 + 3 * (size_t)K));
  const size_t nb_i = sizeof(int) * ((size_t)B * 2 + (size_t)P * (1 + (size_t)K));
  char * base = nullptr;
  if(!check(cudaMalloc(reinterpret_cast<void **>(&base), nb_d + nb_i + 512), "cudaMalloc")) return CCC_ERR_CUDA;
  size_t used = 0;
  bool ok = true;
  auto push = [&](const void * src, size_t bytes) {
    used = (used + 15) & ~size_t(15);
    char * dst = base + used;
    used += bytes;
    if(ok && src && bytes) ok = check(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice), "H2D");
    return dst;
  };
  ccc_step_mpc_batch_t d = *bt;
  d.plan_id = reinterpret_cast<const int32_t *>(push(bt->plan_id, sizeof(int) * B));
  d.x_pos = reinterpret_cast<const double *>(push(bt->x_pos, sizeof(double) * B * 2));
  d.x_vel = reinterpret_cast<const double *>(push(bt->x_vel, sizeof(double) * B * 2));
  d.current_time = reinterpret_cast<const double *>(push(bt->current_time, sizeof(double) * P));
  d.n_elements = reinterpret_cast<const int32_t *>(push(bt->n_elements, sizeof(int) * P));
  d.single = reinterpret_cast<const int32_t *>(push(bt->single, sizeof(int) * (size_t)P * K));
  d.zmp = reinterpret_cast<const double *>(push(bt->zmp, sizeof(double) * (size_t)P * K * 2));
  d.end_time = reinterpret_cast<const double *>(push(bt->end_time, sizeof(double) * (size_t)P * K));
  ccc_step_mpc_result_t r;
  r.current_zmp = reinterpret_cast<double *>(push(nullptr, sizeof(double) * B * 2));
  r.next_foot_zmp = reinterpret_cast<double *>(push(nullptr, sizeof(double) * B * 2));
  r.has_next = reinterpret_cast<int32_t *>(push(nullptr, sizeof(int) * B));
  if(ok)
  {
    step_mpc_kernel<<<(2 * B + 63) / 64, 64>>>(d, r);
    ok = check(cudaGetLastError(), "launch step_mpc_kernel");
  }
  ok = ok && check(cudaMemcpy(res->current_zmp, r.current_zmp, sizeof(double) * B * 2, cudaMemcpyDeviceToHost), "D2H");
  if(ok && res->next_foot_zmp) ok = check(cudaMemcpy(res->next_foot_zmp, r.next_foot_zmp, sizeof(double) * B * 2, cudaMemcpyDeviceToHost), "D2H");
  if(ok && res->has_next) ok = check(cudaMemcpy(res->has_next, r.has_next, sizeof(int) * B, cudaMemcpyDeviceToHost), "D2H");
  cudaFree(base);
  return ok ? CCC_OK : CCC_ERR_CUDA;
}
