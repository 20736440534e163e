// preview_core.cuh — one warp computes one row of u = -K x + F ref_seq (portable: CUDA + warp emulator).
// Evaluation order: oracle ccc_oracle_preview_input (32 lane-strided fma chains + pairwise tree).
#pragma once
#include "warp_ctx.cuh"

namespace ccc
{
/** One warp, one row. */
CCC_DEV double preview_row(int N, const double * K, const double * F, const double * x3, const double * ref)
{
  const int lane = lane_id();
  double acc = 0.0;
  for(int i = lane; i < N; i += 32) acc = dfma(ldg(F + i), ref[i], acc);
  const double total = warp_sum(acc);
  double kx = 0.0;
  CCC_UNROLL
  for(int i = 0; i < 3; i++) kx = dfma(ldg(K + i), x3[i], kx);
  return (-kx) + total;
}
} // namespace ccc

