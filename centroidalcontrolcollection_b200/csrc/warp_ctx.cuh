// warp_ctx.cuh — the handful of warp-level primitives the solver cores are written against.
//
// On the GPU (the product build, nvcc, sm_100a) these are the raw CUDA intrinsics.
// When CCC_WARP_EMU is defined (tests/emu only, g++), the same names are provided by a
// 32-fibre lock-step warp emulator so that the *kernel source itself* can be executed and
// checked bit for bit against the oracle on a machine without a GPU.  The emulator is test
// infrastructure: libccc_b200.so is never built with CCC_WARP_EMU and has no CPU path.
#pragma once

#ifdef CCC_WARP_EMU
#  include <cmath>
#  include <cstdint>
namespace ccc_emu
{
int lane();
void syncwarp();
double shfl(double v, int src);
double shfl_xor(double v, int mask);
int shfl_i(int v, int src);
unsigned ballot(bool p);
} // namespace ccc_emu
#  define CCC_DEV inline
#  define CCC_DEV_NOINLINE
#  define CCC_UNROLL
#  define CCC_UNROLL_N(n)
#  define CCC_NOUNROLL
namespace ccc
{
inline int lane_id() { return ccc_emu::lane(); }
inline void warp_sync() { ccc_emu::syncwarp(); }
inline double warp_shfl(double v, int src) { return ccc_emu::shfl(v, src); }
inline double warp_shfl_xor(double v, int mask) { return ccc_emu::shfl_xor(v, mask); }
inline int warp_shfl_i(int v, int src) { return ccc_emu::shfl_i(v, src); }
inline unsigned warp_ballot(bool p) { return ccc_emu::ballot(p); }
inline double dfma(double a, double b, double c) { return std::fma(a, b, c); }
inline double dsqrt(double a) { return std::sqrt(a); }
inline double drcp(double a) { return 1.0 / a; }
inline double dabs(double a) { return std::fabs(a); }
template<class T>
inline T ldg(const T * p) { return *p; }
} // namespace ccc
#else
#  include <cuda_runtime.h>
#  include <stdint.h>
#  define CCC_DEV __device__ __forceinline__
#  define CCC_DEV_NOINLINE __device__ __noinline__
#  define CCC_UNROLL _Pragma("unroll")
#  define CCC_STR_(x) #x
#  define CCC_UNROLL_N(n) _Pragma(CCC_STR_(unroll n))
#  define CCC_NOUNROLL _Pragma("unroll 1")
namespace ccc
{
constexpr unsigned kFullMask = 0xffffffffu;
CCC_DEV int lane_id() { return static_cast<int>(threadIdx.x & 31u); }
CCC_DEV void warp_sync() { __syncwarp(); }
CCC_DEV double warp_shfl(double v, int src) { return __shfl_sync(kFullMask, v, src); }
CCC_DEV double warp_shfl_xor(double v, int mask) { return __shfl_xor_sync(kFullMask, v, mask); }
CCC_DEV int warp_shfl_i(int v, int src) { return __shfl_sync(kFullMask, v, src); }
CCC_DEV unsigned warp_ballot(bool p) { return __ballot_sync(kFullMask, p); }
CCC_DEV double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
CCC_DEV double dsqrt(double a) { return __dsqrt_rn(a); }
CCC_DEV double drcp(double a) { return __drcp_rn(a); } // correctly rounded 1/a == IEEE 1.0 / a
CCC_DEV double dabs(double a) { return fabs(a); }
template<class T>
CCC_DEV T ldg(const T * p) { return __ldg(p); }
} // namespace ccc
#endif

namespace ccc
{
/** Pairwise tree over the 32 lanes, strides 16, 8, 4, 2, 1 (xor butterfly): every lane ends
 *  with the same bits.  Inactive lanes must contribute +0.0.  This is the engine side of the
 *  oracle's tree_sum32 (oracle/num.hpp). */
CCC_DEV double warp_sum(double v)
{
  v = v + warp_shfl_xor(v, 16);
  v = v + warp_shfl_xor(v, 8);
  v = v + warp_shfl_xor(v, 4);
  v = v + warp_shfl_xor(v, 2);
  v = v + warp_shfl_xor(v, 1);
  return v;
}

/** N independent pairwise-tree sums advanced level by level (same bits as N warp_sum calls,
 *  but the 2N shuffles of a level are in flight together). */
template<int N>
CCC_DEV void warp_sum_n(double (&v)[N])
{
  CCC_UNROLL
  for(int off = 16; off >= 1; off >>= 1)
  {
    double t[N];
    CCC_UNROLL
    for(int i = 0; i < N; i++) t[i] = warp_shfl_xor(v[i], off);
    CCC_UNROLL
    for(int i = 0; i < N; i++) v[i] = v[i] + t[i];
  }
}

/** Eight pairwise-tree sums at once by halving the value set at each butterfly level
 *  ("transpose-reduce"): level 16 exchanges 4 of the 8 values, level 8 two, level 4 one, levels
 *  2 and 1 finish the single value a lane is left with — 9 shuffles instead of 40.  Every value
 *  still goes through the same (16, 8, 4, 2, 1) pairwise tree as warp_sum, so the bits are
 *  identical.  The 8 totals are then broadcast through `xch` (8 doubles of shared memory). */
CCC_DEV void warp_sum8(double (&v)[8], double * xch)
{
  const int lane = lane_id();
  {
    const bool up = (lane & 16) != 0;
    CCC_UNROLL
    for(int i = 0; i < 4; i++)
    {
      const double send = up ? v[i] : v[4 + i];
      const double keep = up ? v[4 + i] : v[i];
      v[i] = keep + warp_shfl_xor(send, 16);
    }
  }
  {
    const bool up = (lane & 8) != 0;
    CCC_UNROLL
    for(int i = 0; i < 2; i++)
    {
      const double send = up ? v[i] : v[2 + i];
      const double keep = up ? v[2 + i] : v[i];
      v[i] = keep + warp_shfl_xor(send, 8);
    }
  }
  {
    const bool up = (lane & 4) != 0;
    const double send = up ? v[0] : v[1];
    const double keep = up ? v[1] : v[0];
    v[0] = keep + warp_shfl_xor(send, 4);
  }
  v[0] = v[0] + warp_shfl_xor(v[0], 2);
  v[0] = v[0] + warp_shfl_xor(v[0], 1);
  // lane holds the total of value ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)
  warp_sync();
  if((lane & 3) == 0) xch[((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = v[0];
  warp_sync();
  CCC_UNROLL
  for(int i = 0; i < 8; i++) v[i] = xch[i];
}

CCC_DEV double warp_max(double v)
{
  CCC_UNROLL
  for(int off = 16; off >= 1; off >>= 1)
  {
    double o = warp_shfl_xor(v, off);
    v = v < o ? o : v;
  }
  return v;
}

CCC_DEV double clampd(double v, double lo, double hi)
{
  double r = v < lo ? lo : v;
  return r > hi ? hi : r;
}

/** cross product: one rounded product + one fma per component (oracle/num.hpp cross3). */
CCC_DEV void cross3(const double * a, const double * b, double * out)
{
  out[0] = dfma(a[1], b[2], -(a[2] * b[1]));
  out[1] = dfma(a[2], b[0], -(a[0] * b[2]));
  out[2] = dfma(a[0], b[1], -(a[1] * b[0]));
}
} // namespace ccc
