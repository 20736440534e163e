// misc.cu — ABI version, device count, last error, nmpc_ddp default configuration.
#include "../../include/ccc_b200.h"
#include "common_host.cuh"

#include <cmath>

namespace
{
/** FP64 peak probe: every thread runs 8 independent fma chains (enough to cover the DFMA latency with 16 warps per
 *  scheduler resident), nothing else in the loop.  2 flops per fma. */
__global__ void __launch_bounds__(512) dfma_peak_kernel(double * out, int iters, double a, double b)
{
  double c[8];
  for(int i = 0; i < 8; i++) c[i] = (double)(threadIdx.x + i);
#pragma unroll 1
  for(int it = 0; it < iters; it++)
  {
#pragma unroll
    for(int r = 0; r < 8; r++)
#pragma unroll
      for(int i = 0; i < 8; i++) c[i] = __fma_rn(c[i], a, b);
  }
  double s = 0.0;
  for(int i = 0; i < 8; i++) s += c[i];
  if(s == 123456.789) out[0] = s; // never true: keeps the chains alive
}
} // namespace

extern "C" {

/* Measured FP64 fma throughput of the current device in TFLOP/s (2 flops per fma; best of `reps` timed launches of a
 * pure-DFMA kernel on `stream`): the denominator of bench.py's roofline.fp64, measured in the same run.  < 0: error. */
double ccc_fp64_peak_tflops(int32_t reps, void * stream_v)
{
  int dev = 0, n_sm = 0;
  if(cudaGetDevice(&dev) != cudaSuccess) return -1.0;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  double * out = nullptr;
  if(cudaMalloc(&out, sizeof(double)) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 4096, threads = 512, blocks = n_sm * 4;
  double best = -1.0;
  for(int r = 0; r < (reps < 1 ? 1 : reps) + 1; r++)
  {
    cudaEventRecord(e0, st);
    dfma_peak_kernel<<<blocks, threads, 0, st>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1, st);
    if(cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 64.0 * iters * (double)threads * blocks;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if(r > 0 && tf > best) best = tf; // launch 0 warms up
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

int32_t ccc_abi_version(void)
{
  return CCC_B200_ABI_VERSION;
}

int32_t ccc_device_count(void)
{
  int n = 0;
  if(cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char * ccc_last_error(void)
{
  return ccc_host::error_buf();
}

/* nmpc_ddp::DDPSolver<>::Configuration and nmpc_ddp::BoxQP<>::Configuration defaults
 * (external dependency of the reference, values as listed in SURVEY.md App. A). */
void ccc_ddp_config_default(ccc_ddp_config_t * c)
{
  if(!c) return;
  c->with_input_constraint = 0;
  c->max_iter = 500;
  c->reg_type = 1;
  c->n_alpha = 11;
  c->initial_lambda = 1e-4;
  c->initial_dlambda = 1.0;
  c->lambda_factor = 1.6;
  c->lambda_min = 1e-6;
  c->lambda_max = 1e10;
  c->k_rel_norm_thre = 1e-4;
  c->lambda_thre = 1e-5;
  c->cost_update_ratio_thre = 0.0;
  c->cost_update_thre = 1e-7;
  for(int i = 0; i < CCC_DDP_MAX_ALPHA; i++) c->alpha[i] = i < 11 ? std::pow(10.0, -3.0 * i / 10.0) : 0.0;
  c->boxqp_max_iter = 500;
  c->reserved0 = 0;
  c->boxqp_grad_thre = 1e-8;
  c->boxqp_rel_improve_thre = 1e-8;
  c->boxqp_step_factor = 0.6;
  c->boxqp_min_step = 1e-22;
  c->boxqp_armijo = 0.1;
}

} // extern "C"
