// preview.cu — ccc_preview_input (include/ccc_b200.h): batched u = -K x + F ref_seq.
//
// One warp per row: the N-long reference sequence is streamed with coalesced 256-byte reads (lane l takes
// entries l, l+32, ...), reduced with the pairwise-tree warp sum.  This path IS HBM bound: (N + 3 + 1) * 8
// bytes per row and ~2 flops per byte.  Evaluation order: oracle ccc_oracle_preview_input (bit-exact).
#include "../../include/ccc_b200.h"
#include "common_host.cuh"
#include "preview_core.cuh"

namespace
{
__global__ void __launch_bounds__(256) preview_kernel(int B, int N, const double * __restrict__ K, const double * __restrict__ F,
                                                      const double * __restrict__ x, const double * __restrict__ ref,
                                                      double * __restrict__ u)
{
  const int warps_per_grid = gridDim.x * (blockDim.x >> 5);
  for(int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < B; b += warps_per_grid)
  {
    const double v = ccc::preview_row(N, K, F, x + (size_t)b * 3, ref + (size_t)b * N);
    if((threadIdx.x & 31) == 0) u[b] = v;
  }
}
} // namespace

extern "C" int32_t ccc_preview_input(int32_t B, int32_t N, const double * K, const double * F, const double * x,
                                     const double * ref_seq, double * u, int32_t mem, void * stream_v)
{
  using ccc_host::check;
  if(B <= 0 || N <= 0 || !K || !F || !x || !ref_seq || !u) return ccc_host::fail(CCC_ERR_INVALID, "bad argument");
  int ndev = 0;
  if(!check(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount") || ndev == 0)
    return ccc_host::fail(CCC_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  double *dK = nullptr, *dF = nullptr, *dx = nullptr, *dref = nullptr, *du = nullptr;
  if(mem == CCC_MEM_HOST)
  {
    st = nullptr;
    bool ok = check(cudaMalloc(&dK, 3 * sizeof(double)), "cudaMalloc") && check(cudaMalloc(&dF, N * sizeof(double)), "cudaMalloc")
              && check(cudaMalloc(&dx, (size_t)B * 3 * sizeof(double)), "cudaMalloc")
              && check(cudaMalloc(&dref, (size_t)B * N * sizeof(double)), "cudaMalloc")
              && check(cudaMalloc(&du, (size_t)B * sizeof(double)), "cudaMalloc");
    ok = ok && check(cudaMemcpy(dK, K, 3 * sizeof(double), cudaMemcpyHostToDevice), "H2D")
         && check(cudaMemcpy(dF, F, N * sizeof(double), cudaMemcpyHostToDevice), "H2D")
         && check(cudaMemcpy(dx, x, (size_t)B * 3 * sizeof(double), cudaMemcpyHostToDevice), "H2D")
         && check(cudaMemcpy(dref, ref_seq, (size_t)B * N * sizeof(double), cudaMemcpyHostToDevice), "H2D");
    if(!ok)
    {
      cudaFree(dK), cudaFree(dF), cudaFree(dx), cudaFree(dref), cudaFree(du);
      return CCC_ERR_CUDA;
    }
    K = dK, F = dF, x = dx, ref_seq = dref;
  }
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  int grid = (B + 7) / 8;
  if(grid > n_sm * 8) grid = n_sm * 8; // 8 CTAs of 8 warps per SM: a whole number of waves
  preview_kernel<<<grid, 256, 0, st>>>(B, N, K, F, x, ref_seq, mem == CCC_MEM_HOST ? du : u);
  int rc = check(cudaGetLastError(), "launch preview_kernel") ? CCC_OK : CCC_ERR_CUDA;
  if(mem == CCC_MEM_HOST)
  {
    if(rc == CCC_OK && !check(cudaMemcpy(u, du, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost), "D2H")) rc = CCC_ERR_CUDA;
    cudaFree(dK), cudaFree(dF), cudaFree(dx), cudaFree(dref), cudaFree(du);
  }
  return rc;
}
