// closed_form.cu — ccc_dcm_tracking_plan / ccc_foot_guided_plan / ccc_singular_preview_plan (include/ccc_b200.h): the
// closed-form ZMP controllers of the reference, batched; one thread per problem (and axis).
//
// Replaces: CCC::DcmTracking::planOnce (reference src/DcmTracking.cpp:7-48) and CCC::FootGuidedControl1d::planOnce
// (src/FootGuidedControl.cpp:11-69) x 2 axes (:71-92).  The expressions are the reference's, term by term (the file is
// compiled with -fmad=false); exp() is CUDA's.
#include "../../include/ccc_b200.h"
#include "common_host.cuh"

#include <vector>

namespace
{
__global__ void dcm_tracking_kernel(ccc_dcm_tracking_batch_t in, double * __restrict__ out)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= in.batch) return;
  const int p = in.plan_id[b], K = in.max_knots;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  if(p < 0 || p >= in.n_plans)
  {
    out[2 * b] = out[2 * b + 1] = nan;
    return;
  }
  const int nk = in.n_knots[p];
  const double t = in.current_time[p];
  const double * kt = in.knot_time + (size_t)p * K;
  const double * kz = in.knot_zmp + (size_t)p * K * 2;
  bool bad = nk < 0 || nk > K;
  for(int i = 0; i < nk && !bad; i++) bad = kt[i] < t; // "ZMP switching time must be in the future" (:20-27)
  for(int a = 0; a < 2; a++)
  {
    const double cz = in.current_zmp[2 * p + a];
    double target = cz;
    if(nk > 0 && !bad)
    {
      // DCM at the last switch, carried back to the first one: equation (18) (:29-36)
      double dcm_switch = kz[(nk - 1) * 2 + a];
      for(int i = nk - 2; i >= 0; i--)
      {
        const double zmp_duration = kt[i + 1] - kt[i];
        dcm_switch = kz[i * 2 + a] + exp(-1 * in.omega * zmp_duration) * (dcm_switch - kz[i * 2 + a]);
      }
      // target DCM now: equation (19) (:39-42)
      target = cz + exp(in.omega * (t - kt[0])) * (dcm_switch - cz);
    }
    // control ZMP: equation (24) (:44-45)
    const double z = cz + (1.0 + in.feedback_gain / in.omega) * (in.dcm[2 * b + a] - target);
    out[2 * b + a] = bad ? nan : z;
  }
}

__global__ void foot_guided_kernel(ccc_foot_guided_batch_t in, double * __restrict__ out)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= in.batch) return;
  const int p = in.plan_id[b];
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  if(p < 0 || p >= in.n_plans)
  {
    out[2 * b] = out[2 * b + 1] = nan;
    return;
  }
  const double omega = in.omega, current_time = in.current_time[p];
  const double transit_start_time = in.transit_start_time[p], transit_duration = in.transit_duration[p];
  const double transit_end_time = transit_start_time + transit_duration;
  const double future_margin_duration = 1e-6;
  const bool bad = !(transit_duration >= 0) || !(transit_end_time >= current_time + future_margin_duration); // :17-29
  for(int a = 0; a < 2; a++)
  {
    const double capture_point = in.capture_point[2 * b + a];
    const double start_zmp = in.transit_start_zmp[2 * p + a], end_zmp = in.transit_end_zmp[2 * p + a];
    double planned_zmp;
    if(transit_duration == 0)
    {
      // equation (7) of Kojio's paper (:34-40)
      planned_zmp = start_zmp
                    + 2 * ((capture_point - start_zmp) - (end_zmp - start_zmp) * exp(-1 * omega * (transit_start_time - current_time)))
                          / (1.0 - exp(-2 * omega * (transit_start_time - current_time)));
    }
    else
    {
      const double zmp_transit_vel = (end_zmp - start_zmp) / transit_duration;
      if(current_time <= transit_start_time)
      {
        // equation (17) (:47-53)
        planned_zmp = start_zmp
                      + (2 * (capture_point - start_zmp)
                         + 2 * zmp_transit_vel / omega
                               * (exp(-1 * omega * (transit_end_time - current_time)) - exp(-1 * omega * (transit_start_time - current_time))))
                            / (1.0 - exp(-2 * omega * (transit_end_time - current_time)));
      }
      else
      {
        const double current_ref_zmp = start_zmp + zmp_transit_vel * (current_time - transit_start_time);
        // equation (17) (:57-63)
        planned_zmp = current_ref_zmp
                      + (2 * (capture_point - current_ref_zmp) + 2 * zmp_transit_vel / omega * (exp(-1 * omega * (transit_end_time - current_time)) - 1.0))
                            / (1.0 - exp(-2 * omega * (transit_end_time - current_time)));
      }
    }
    out[2 * b + a] = bad ? nan : planned_zmp;
  }
}

/** Feed-forward term of SingularPreviewControlZmp1d::procOnce (reference src/SingularPreviewControlZmp.cpp:38-48) for one
 *  (plan, axis): S0 by the backward recursion of equation (25), then u_ff of equation (26).  One thread per (plan, axis). */
__global__ void singular_preview_ff_kernel(int P, int N, double omega, double dt, const double * __restrict__ ref_zmp, double * __restrict__ u_ff)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if(idx >= 2 * P) return;
  const int p = idx >> 1, a = idx & 1;
  const double * seq = ref_zmp + (size_t)p * N * 2 + a; // stride 2
  const double b = 1 + omega * dt;
  double S0 = seq[(size_t)(N - 1) * 2] * (1 + omega * dt) / (omega * dt);
  for(int i = N - 2; i >= 1; i--) S0 = seq[(size_t)i * 2] + S0 / b;
  u_ff[idx] = seq[0] / dt - omega * (2 + omega * dt) * S0 / (b * b);
}

/** Feedback term (equation (16), :30-36), u = u_fb + u_ff (equation (9)) and the ZMP update (:52-55); one thread per
 *  (problem, axis). */
__global__ void singular_preview_kernel(ccc_singular_preview_batch_t in, const double * __restrict__ u_ff, double * __restrict__ out)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if(idx >= 2 * in.batch) return;
  const int b = idx >> 1, a = idx & 1;
  const int p = in.plan_id[b];
  if(p < 0 || p >= in.n_plans)
  {
    out[idx] = __longlong_as_double(0x7ff8000000000000LL);
    return;
  }
  const double omega = in.omega, dt = in.horizon_dt;
  const double * x = in.state + (size_t)idx * 3;
  const double K0 = (1 + omega * dt) / dt + omega / (1 + omega * dt);
  const double K1 = -1 * (2 + omega * dt) / dt;
  const double K2 = -1 * (2 + omega * dt) / (omega * dt);
  // Eigen's 3-vector dot product: K[0] x[0] + K[1] x[1] + K[2] x[2], left to right
  const double u_fb = -1 * ((K0 * x[0] + K1 * x[1]) + K2 * x[2]);
  const double u = u_fb + u_ff[2 * p + a];
  out[idx] = x[0] + in.control_dt * u;
}

/** Host-buffer staging: one device allocation for the doubles, one for the ints; `push` copies and returns the device address. */
struct Stage
{
  char * base = nullptr;
  size_t cap = 0, used = 0;
  bool ok = true;
  explicit Stage(size_t bytes) : cap(bytes) { ok = ccc_host::check(cudaMalloc(reinterpret_cast<void **>(&base), bytes ? bytes : 8), "cudaMalloc"); }
  ~Stage() { cudaFree(base); }
  template<class T>
  const T * push(const T * src, size_t n)
  {
    used = (used + 15) & ~size_t(15);
    T * dst = reinterpret_cast<T *>(base + used);
    used += n * sizeof(T);
    if(ok && n && used <= cap) ok = ccc_host::check(cudaMemcpy(dst, src, n * sizeof(T), cudaMemcpyHostToDevice), "H2D");
    return dst;
  }
  template<class T>
  T * room(size_t n)
  {
    used = (used + 15) & ~size_t(15);
    T * dst = reinterpret_cast<T *>(base + used);
    used += n * sizeof(T);
    return dst;
  }
};

bool have_device()
{
  int ndev = 0;
  return ccc_host::check(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount") && ndev > 0;
}
} // namespace

extern "C" {

int32_t ccc_dcm_tracking_plan(const ccc_dcm_tracking_batch_t * bt, double * control_zmp, int32_t mem, void * stream)
{
  using ccc_host::check;
  if(!bt || !control_zmp) return ccc_host::fail(CCC_ERR_INVALID, "null argument");
  const int B = bt->batch, P = bt->n_plans, K = bt->max_knots;
  if(B <= 0 || P <= 0 || K < 0 || !(bt->omega > 0)) return ccc_host::fail(CCC_ERR_INVALID, "ccc_dcm_tracking_plan: sizes / omega");
  if(!bt->plan_id || !bt->dcm || !bt->current_time || !bt->current_zmp || !bt->n_knots || (K && (!bt->knot_time || !bt->knot_zmp)))
    return ccc_host::fail(CCC_ERR_INVALID, "null input");
  if(mem != CCC_MEM_HOST)
  {
    dcm_tracking_kernel<<<(B + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*bt, control_zmp);
    return check(cudaGetLastError(), "launch dcm_tracking_kernel") ? CCC_OK : CCC_ERR_CUDA;
  }
  // the reference's exceptions (src/DcmTracking.cpp:20-27)
  for(int b = 0; b < B; b++)
    if(bt->plan_id[b] < 0 || bt->plan_id[b] >= P) return ccc_host::fail(CCC_ERR_INVALID, "plan_id out of range");
  for(int p = 0; p < P; p++)
  {
    if(bt->n_knots[p] < 0 || bt->n_knots[p] > K) return ccc_host::fail(CCC_ERR_INVALID, "n_knots out of range");
    for(int i = 0; i < bt->n_knots[p]; i++)
      if(bt->knot_time[(size_t)p * K + i] < bt->current_time[p]) return ccc_host::fail(CCC_ERR_INVALID, "ZMP switching time must be in the future");
  }
  if(!have_device()) return ccc_host::fail(CCC_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
  Stage st(sizeof(double) * ((size_t)B * 4 + (size_t)P * (3 + 3 * (size_t)K)) + sizeof(int) * ((size_t)B + P) + 256);
  ccc_dcm_tracking_batch_t d = *bt;
  d.plan_id = st.push(bt->plan_id, B);
  d.dcm = st.push(bt->dcm, (size_t)B * 2);
  d.current_time = st.push(bt->current_time, P);
  d.current_zmp = st.push(bt->current_zmp, (size_t)P * 2);
  d.n_knots = st.push(bt->n_knots, P);
  d.knot_time = st.push(bt->knot_time, (size_t)P * K);
  d.knot_zmp = st.push(bt->knot_zmp, (size_t)P * K * 2);
  double * out = st.room<double>((size_t)B * 2);
  if(!st.ok) return CCC_ERR_CUDA;
  dcm_tracking_kernel<<<(B + 127) / 128, 128>>>(d, out);
  if(!check(cudaGetLastError(), "launch dcm_tracking_kernel")) return CCC_ERR_CUDA;
  return check(cudaMemcpy(control_zmp, out, sizeof(double) * B * 2, cudaMemcpyDeviceToHost), "D2H") ? CCC_OK : CCC_ERR_CUDA;
}

int32_t ccc_foot_guided_plan(const ccc_foot_guided_batch_t * bt, double * planned_zmp, int32_t mem, void * stream)
{
  using ccc_host::check;
  if(!bt || !planned_zmp) return ccc_host::fail(CCC_ERR_INVALID, "null argument");
  const int B = bt->batch, P = bt->n_plans;
  if(B <= 0 || P <= 0 || !(bt->omega > 0)) return ccc_host::fail(CCC_ERR_INVALID, "ccc_foot_guided_plan: sizes / omega");
  if(!bt->plan_id || !bt->capture_point || !bt->current_time || !bt->transit_start_zmp || !bt->transit_end_zmp || !bt->transit_start_time
     || !bt->transit_duration)
    return ccc_host::fail(CCC_ERR_INVALID, "null input");
  if(mem != CCC_MEM_HOST)
  {
    foot_guided_kernel<<<(B + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*bt, planned_zmp);
    return check(cudaGetLastError(), "launch foot_guided_kernel") ? CCC_OK : CCC_ERR_CUDA;
  }
  // the reference's exceptions (src/FootGuidedControl.cpp:17-29)
  for(int b = 0; b < B; b++)
    if(bt->plan_id[b] < 0 || bt->plan_id[b] >= P) return ccc_host::fail(CCC_ERR_INVALID, "plan_id out of range");
  for(int p = 0; p < P; p++)
  {
    if(!(bt->transit_duration[p] >= 0)) return ccc_host::fail(CCC_ERR_INVALID, "Transition duration must be non-negative");
    if(!(bt->transit_start_time[p] + bt->transit_duration[p] >= bt->current_time[p] + 1e-6))
      return ccc_host::fail(CCC_ERR_INVALID, "Transition end time must be in the future with some margin");
  }
  if(!have_device()) return ccc_host::fail(CCC_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
  Stage st(sizeof(double) * ((size_t)B * 4 + (size_t)P * 7) + sizeof(int) * (size_t)B + 256);
  ccc_foot_guided_batch_t d = *bt;
  d.plan_id = st.push(bt->plan_id, B);
  d.capture_point = st.push(bt->capture_point, (size_t)B * 2);
  d.current_time = st.push(bt->current_time, P);
  d.transit_start_zmp = st.push(bt->transit_start_zmp, (size_t)P * 2);
  d.transit_end_zmp = st.push(bt->transit_end_zmp, (size_t)P * 2);
  d.transit_start_time = st.push(bt->transit_start_time, P);
  d.transit_duration = st.push(bt->transit_duration, P);
  double * out = st.room<double>((size_t)B * 2);
  if(!st.ok) return CCC_ERR_CUDA;
  foot_guided_kernel<<<(B + 127) / 128, 128>>>(d, out);
  if(!check(cudaGetLastError(), "launch foot_guided_kernel")) return CCC_ERR_CUDA;
  return check(cudaMemcpy(planned_zmp, out, sizeof(double) * B * 2, cudaMemcpyDeviceToHost), "D2H") ? CCC_OK : CCC_ERR_CUDA;
}

int32_t ccc_singular_preview_plan(const ccc_singular_preview_batch_t * bt, double * planned_zmp, int32_t mem, void * stream)
{
  using ccc_host::check;
  if(!bt || !planned_zmp) return ccc_host::fail(CCC_ERR_INVALID, "null argument");
  const int B = bt->batch, P = bt->n_plans, N = bt->horizon_steps;
  if(B <= 0 || P <= 0 || N < 1 || !(bt->omega > 0) || !(bt->horizon_dt > 0))
    return ccc_host::fail(CCC_ERR_INVALID, "ccc_singular_preview_plan: sizes / omega / horizon_dt");
  if(!bt->plan_id || !bt->state || !bt->ref_zmp) return ccc_host::fail(CCC_ERR_INVALID, "null input");
  if(mem != CCC_MEM_HOST)
  {
    // device pointers: the per-plan feed-forward terms need scratch of their own (freed on the stream's order)
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    double * u_ff = nullptr;
    if(!check(cudaMallocAsync(reinterpret_cast<void **>(&u_ff), sizeof(double) * 2 * (size_t)P, st), "cudaMallocAsync")) return CCC_ERR_CUDA;
    singular_preview_ff_kernel<<<(2 * P + 127) / 128, 128, 0, st>>>(P, N, bt->omega, bt->horizon_dt, bt->ref_zmp, u_ff);
    singular_preview_kernel<<<(2 * B + 127) / 128, 128, 0, st>>>(*bt, u_ff, planned_zmp);
    const bool ok = check(cudaGetLastError(), "launch singular_preview_kernel");
    cudaFreeAsync(u_ff, st);
    return ok ? CCC_OK : CCC_ERR_CUDA;
  }
  for(int b = 0; b < B; b++)
    if(bt->plan_id[b] < 0 || bt->plan_id[b] >= P) return ccc_host::fail(CCC_ERR_INVALID, "plan_id out of range");
  if(!have_device()) return ccc_host::fail(CCC_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
  Stage st(sizeof(double) * ((size_t)B * 8 + (size_t)P * (2 * (size_t)N + 2)) + sizeof(int) * (size_t)B + 256);
  ccc_singular_preview_batch_t d = *bt;
  d.plan_id = st.push(bt->plan_id, B);
  d.state = st.push(bt->state, (size_t)B * 6);
  d.ref_zmp = st.push(bt->ref_zmp, (size_t)P * N * 2);
  double * u_ff = st.room<double>((size_t)P * 2);
  double * out = st.room<double>((size_t)B * 2);
  if(!st.ok) return CCC_ERR_CUDA;
  singular_preview_ff_kernel<<<(2 * P + 127) / 128, 128>>>(P, N, bt->omega, bt->horizon_dt, d.ref_zmp, u_ff);
  singular_preview_kernel<<<(2 * B + 127) / 128, 128>>>(d, u_ff, out);
  if(!check(cudaGetLastError(), "launch singular_preview_kernel")) return CCC_ERR_CUDA;
  return check(cudaMemcpy(planned_zmp, out, sizeof(double) * B * 2, cudaMemcpyDeviceToHost), "D2H") ? CCC_OK : CCC_ERR_CUDA;
}

} // extern "C"
