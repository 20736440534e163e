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
#  include <cstring>
namespace ccc_emu
{
int lane();
int tid();
void syncwarp();
void syncthreads();
double shfl(double v, int src);
double shfl_xor(double v, int mask);
int shfl_i(int v, int src);
unsigned ballot(bool p);
} // namespace ccc_emu
#  define CCC_DEV inline
#  define CCC_HD inline
#  define CCC_DEV_NOINLINE
#  define CCC_UNROLL
#  define CCC_UNROLL_N(n)
#  define CCC_NOUNROLL
namespace ccc
{
inline int lane_id() { return ccc_emu::lane(); }
inline void warp_sync() { ccc_emu::syncwarp(); }
inline int thread_id() { return ccc_emu::tid(); }
inline void cta_sync() { ccc_emu::syncthreads(); }
inline double warp_shfl(double v, int src) { return ccc_emu::shfl(v, src); }
inline double warp_shfl_xor(double v, int mask) { return ccc_emu::shfl_xor(v, mask); }
inline int warp_shfl_i(int v, int src) { return ccc_emu::shfl_i(v, src); }
inline unsigned warp_ballot(bool p) { return ccc_emu::ballot(p); }
inline double dfma(double a, double b, double c) { return std::fma(a, b, c); }
inline double dsqrt(double a) { return std::sqrt(a); }
inline double drcp(double a) { return 1.0 / a; }
inline double drint(double a) { return std::nearbyint(a); }
inline double dabs(double a) { return std::fabs(a); }
template<class T>
inline T ldg(const T * p) { return *p; }
inline double ldcg(const double * p) { return *p; }
inline void prefetch_l1(const void *) {}
// TMA / mbarrier stand-ins for the lock-step emulator: the copy is done at once by the issuing fibre and the wait is a
// warp barrier (every fibre of the warp waits at the same point), which is the ordering the mbarrier gives on the GPU
inline void mbar_init(unsigned long long *, int) {}
inline void mbar_arrive_expect_tx(unsigned long long *, unsigned) {}
inline void tma_bulk_g2s(void * dst, const void * src, unsigned bytes, unsigned long long *) { std::memcpy(dst, src, bytes); }
inline void tma_bulk_s2g(void * dst, const void * src, unsigned bytes) { std::memcpy(dst, src, bytes); }
inline void mbar_wait(unsigned long long *, unsigned) { ccc_emu::syncwarp(); }
inline void fence_proxy_async_smem() {}
inline void fence_proxy_async_all() {}
inline void bulk_commit() {}
inline void bulk_wait_read_all() {}
inline void bulk_wait_all() {}
} // namespace ccc
#  define CCC_HAS_TMA 0
#else
#  include <cuda_runtime.h>
#  include <stdint.h>
#  define CCC_DEV __device__ __forceinline__
#  define CCC_HD __host__ __device__ __forceinline__
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
CCC_DEV int thread_id() { return static_cast<int>(threadIdx.x); }
CCC_DEV void cta_sync() { __syncthreads(); }
CCC_DEV double warp_shfl(double v, int src) { return __shfl_sync(kFullMask, v, src); }
CCC_DEV double warp_shfl_xor(double v, int mask) { return __shfl_xor_sync(kFullMask, v, mask); }
CCC_DEV int warp_shfl_i(int v, int src) { return __shfl_sync(kFullMask, v, src); }
CCC_DEV unsigned warp_ballot(bool p) { return __ballot_sync(kFullMask, p); }
CCC_DEV double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
CCC_DEV double dsqrt(double a) { return __dsqrt_rn(a); }
CCC_DEV double drcp(double a) { return __drcp_rn(a); } // correctly rounded 1/a == IEEE 1.0 / a
CCC_DEV double drint(double a) { return rint(a); }
CCC_DEV double dabs(double a) { return fabs(a); }
template<class T>
CCC_DEV T ldg(const T * p) { return __ldg(p); }
/** Load that bypasses L1 (data written by the TMA engine, which does not update this SM's L1). */
CCC_DEV double ldcg(const double * p) { return __ldcg(p); }
/** Hint: bring the 128-byte line holding p into L1 (no register, no scoreboard wait). */
CCC_DEV void prefetch_l1(const void * p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// ---- TMA bulk copy global -> shared memory, completion on an mbarrier (sm_90+; SASS: UBLKCP / SYNCS) ----
CCC_DEV unsigned smem_u32(const void * p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
CCC_DEV void mbar_init(unsigned long long * bar, int arrivals)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
/** One arrival + the number of bytes the bulk copies of this phase will deliver. */
CCC_DEV void mbar_arrive_expect_tx(unsigned long long * bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
/** dst (shared, 16-byte aligned) <- src (global, 16-byte aligned), bytes a multiple of 16. */
CCC_DEV void tma_bulk_g2s(void * dst, const void * src, unsigned bytes, unsigned long long * bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
CCC_DEV void mbar_wait(unsigned long long * bar, unsigned parity)
{
  unsigned done = 0;
  while(!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
}
/** Orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes. */
CCC_DEV void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
/** The same for every state space (generic-proxy stores to global that a later bulk copy overwrites or reads). */
CCC_DEV void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
/** dst (global, 16-byte aligned) <- src (shared, 16-byte aligned), bytes a multiple of 16; completion is tracked by
 *  the issuing thread's bulk async-group (SASS: UBLKCP.S2G / UTMASTG-less bulk store). */
CCC_DEV void tma_bulk_s2g(void * dst, const void * src, unsigned bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
CCC_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
/** All of this thread's bulk stores have finished READING their shared-memory source (it may be overwritten). */
CCC_DEV void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
/** All of this thread's bulk stores are complete (their global writes performed). */
CCC_DEV void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
} // namespace ccc
#  define CCC_HAS_TMA 1
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

/** Two pairwise-tree sums with 6 shuffles instead of 10: level 16 leaves value 0 in the lower
 *  half-warp and value 1 in the upper one (each lane adds the partner's copy of the value it
 *  keeps, exactly the butterfly's pair), levels 8..1 finish both inside their halves, one more
 *  exchange hands every lane the other total.  Same tree, same bits as two warp_sum calls. */
CCC_DEV void warp_sum2(double (&v)[2])
{
  const bool up = (lane_id() & 16) != 0;
  const double send = up ? v[0] : v[1];
  double keep = up ? v[1] : v[0];
  keep = keep + warp_shfl_xor(send, 16);
  keep = keep + warp_shfl_xor(keep, 8);
  keep = keep + warp_shfl_xor(keep, 4);
  keep = keep + warp_shfl_xor(keep, 2);
  keep = keep + warp_shfl_xor(keep, 1);
  const double other = warp_shfl_xor(keep, 16);
  v[0] = up ? other : keep;
  v[1] = up ? keep : other;
}

/** Four pairwise-tree sums by the same halving, WITHOUT the final broadcast: 6 shuffles, after which a lane holds the
 *  total of value 2 * bit4(lane) + bit3(lane).  For per-stage accumulations whose totals are only needed at the end of
 *  a pass (dV, the small-gradient norm): every lane group accumulates its own quantity and the pass ends with one
 *  broadcast per quantity.  Same (16, 8, 4, 2, 1) tree and bits as warp_sum. */
CCC_DEV double warp_sum4_scattered(const double (&v)[4])
{
  const int lane = lane_id();
  double a[2];
  {
    const bool up = (lane & 16) != 0;
    CCC_UNROLL
    for(int i = 0; i < 2; i++)
    {
      const double send = up ? v[i] : v[2 + i];
      const double keep = up ? v[2 + i] : v[i];
      a[i] = keep + warp_shfl_xor(send, 16);
    }
  }
  double r;
  {
    const bool up = (lane & 8) != 0;
    const double send = up ? a[0] : a[1];
    const double keep = up ? a[1] : a[0];
    r = keep + warp_shfl_xor(send, 8);
  }
  r = r + warp_shfl_xor(r, 4);
  r = r + warp_shfl_xor(r, 2);
  r = r + warp_shfl_xor(r, 1);
  return r;
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

/** Every lane prefetches one 128-byte line of the span [p, p + bytes) (bytes is a compile-time
 *  constant at the call sites, so the loop is one or two predicated instructions). */
CCC_DEV void prefetch_span(const void * p, int bytes)
{
  const char * c = static_cast<const char *>(p);
  CCC_UNROLL
  for(int off = 0; off < bytes; off += 32 * 128)
  {
    const int o = off + lane_id() * 128;
    if(o < bytes) prefetch_l1(c + o);
  }
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

/** sin and cos from +,-,*,fma and rint only (oracle/num.hpp sincos_canon, same sequence). */
CCC_DEV void sincos_canon(double x, double & s_out, double & c_out)
{
  const double j = drint(x * 6.36619772367581382433e-01);
  double r = dfma(-j, 1.57079632673412561417e+00, x);
  r = dfma(-j, 6.07710050650619224932e-11, r);
  r = dfma(-j, 2.02226624879595063154e-21, r);
  const double z = r * r;
  double ps = 1.58969099521155010221e-10;
  ps = dfma(ps, z, -2.50507602534068634195e-08);
  ps = dfma(ps, z, 2.75573137070700676789e-06);
  ps = dfma(ps, z, -1.98412698298579493134e-04);
  ps = dfma(ps, z, 8.33333333332248946124e-03);
  ps = dfma(ps, z, -1.66666666666666324348e-01);
  const double sr = dfma(r * z, ps, r);
  double pc = -1.13596475577881948265e-11;
  pc = dfma(pc, z, 2.08757232129817482790e-09);
  pc = dfma(pc, z, -2.75573143513906633035e-07);
  pc = dfma(pc, z, 2.48015872894767294178e-05);
  pc = dfma(pc, z, -1.38888888888741095749e-03);
  pc = dfma(pc, z, 4.16666666666666019037e-02);
  const double cr = dfma(z * z, pc, dfma(-0.5, z, 1.0));
  const long long q = static_cast<long long>(j) & 3;
  s_out = q == 0 ? sr : q == 1 ? cr : q == 2 ? -sr : -cr;
  c_out = q == 0 ? cr : q == 1 ? -sr : q == 2 ? -cr : sr;
}

/** cross product: one rounded product + one fma per component (oracle/num.hpp cross3). */
CCC_DEV void cross3(const double * a, const double * b, double * out)
{
  out[0] = dfma(a[1], b[2], -(a[2] * b[1]));
  out[1] = dfma(a[2], b[0], -(a[0] * b[2]));
  out[2] = dfma(a[0], b[1], -(a[1] * b[0]));
}
} // namespace ccc
