// oracle/num.hpp — canonical FP64 primitives of the CPU oracle.
//
// TEST INFRASTRUCTURE ONLY (see oracle/README.md): nothing outside tests/, bench.py's
// cpu_baseline / --impl reference legs and __graft_entry__.smoke() may link this.
//
// The reference does its arithmetic through Eigen, whose evaluation order is not part of
// its contract and cannot be observed here (Eigen is absent).  The oracle therefore fixes
// one explicit, documented evaluation order ("canonical arithmetic"):
//   * every multiply-accumulate is an explicit IEEE fma(); the file is compiled with
//     -ffp-contract=off so the compiler adds or removes none;
//   * inner products over a state index (<= 13 terms) and matrix products run as one
//     sequential fma chain in ascending index order starting from +0.0; matrix-vector products
//     over the input index use four interleaved chains (dot4);
//   * scalar / 3-vector reductions over the ridge (input) index use a fixed 32-leaf
//     stride-halving pairwise tree (tree_sum32), zero padded;
//   * divisions and square roots are IEEE correctly rounded operations.
// Only +,-,*,fma,/ and sqrt are used on the DdpCentroidal / DdpZmp / QP paths, so a
// second implementation that follows the same order reproduces these results bit for bit.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace oracle
{
constexpr double kGravity = 9.80665; // reference include/CCC/Constants.h:10

// ---- arithmetic variant (compile time) -------------------------------------------------------
// The default build is the CANONICAL arithmetic above, the one the engine reproduces bit for bit.
// -DORACLE_TEXTBOOK builds the same algorithm in the arithmetic a straightforward Eigen/libm
// program would use (libccc_oracle_textbook.so): no fused multiply-add, plain left-to-right sums,
// Cholesky with square roots and divisions (Eigen::LLT), divided Armijo ratio, two-term BoxQP
// objective, gradient norm with its square root, P / m as a division, libm sin/cos, sequential
// Givens sweep in the QP.  tests/test_oracle_textbook.py compares the two builds: every place where
// the canonical form departs from the textbook one for the engine's sake is covered by that test.
#ifdef ORACLE_TEXTBOOK
constexpr bool kTextbook = true;
#else
constexpr bool kTextbook = false;
#endif

// ---- algorithmic choices that nmpc_ddp's source would pin (run time, study only) ------------------
// nmpc_ddp is not under /root/reference; where the recollection of its source leaves a choice open,
// the oracle's default is the recalled form and the alternative is selectable here so that
// tools/srb_robustness.py can measure what each choice does to the reference's closed-loop tests.
enum Choice : uint32_t
{
  kChoiceKRelElementwise = 1u << 0, // small-gradient test on max_ij |k_ij| / (|u_ij| + 1) (iLQG.m) instead of
                                    // max_i ||k_i|| / (||u_i|| + 1) (nmpc_ddp)
  kChoiceBoxQpColdStart = 1u << 1,  // BoxQP starts from 0 instead of the gain of the next stage
  kChoiceVxxRegularised = 1u << 2,  // dV / Vx / Vxx built with Quu + lambda I instead of Quu
  kChoiceDescentTol = 1u << 3,      // BoxQP 'no descent' test: s'g > 1e-10 instead of s'g >= 0
  kChoiceBoxQpWarmSameStage = 1u << 4, // BoxQP starts from this stage's gain of the previous backward pass
};
inline uint32_t & choiceBits()
{
  static uint32_t bits = 0;
  return bits;
}
inline bool choice(uint32_t c)
{
  return (choiceBits() & c) != 0;
}

/** a * b + c: one rounding (IEEE fma) in the canonical build, two in the textbook build. */
inline double fmad(double a, double b, double c)
{
  if(kTextbook) return a * b + c;
  return std::fma(a, b, c);
}

/** Sequential fma chain: sum_i a[i*sa] * b[i*sb], ascending i, starting from +0.0. */
inline double dot_seq(const double * a, int sa, const double * b, int sb, int n)
{
  double acc = 0.0;
  for(int i = 0; i < n; i++) acc = fmad(a[i * sa], b[i * sb], acc);
  return acc;
}

/** Four interleaved fma chains (term i goes to chain i % 4, ascending i, each from +0.0),
 *  combined as (s0 + s1) + (s2 + s3).  Used for the matrix-vector products over the input index
 *  (H x in BoxQP, Quu k in the backward pass): the usual 4-way unrolled dot product. */
inline double dot4(const double * a, int sa, const double * b, int sb, int n)
{
  if(kTextbook) return dot_seq(a, sa, b, sb, n);
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for(int i = 0; i < n; i++) s[i & 3] = fmad(a[i * sa], b[i * sb], s[i & 3]);
  return (s[0] + s[1]) + (s[2] + s[3]);
}

/** Pairwise tree over 32 zero-padded leaves: level strides 16, 8, 4, 2, 1. */
inline double tree_sum32(const double * p, int m)
{
  if(kTextbook)
  {
    double acc = 0.0;
    for(int i = 0; i < m; i++) acc = acc + p[i];
    return acc;
  }
  double t[32];
  for(int i = 0; i < 32; i++) t[i] = i < m ? p[i] : 0.0;
  for(int off = 16; off >= 1; off >>= 1)
    for(int i = 0; i < off; i++) t[i] = t[i] + t[i + off];
  return t[0];
}

/** Cross product with one rounded product and one fma per component. */
inline void cross3(const double * a, const double * b, double * out)
{
  if(kTextbook)
  {
    out[0] = a[1] * b[2] - a[2] * b[1];
    out[1] = a[2] * b[0] - a[0] * b[2];
    out[2] = a[0] * b[1] - a[1] * b[0];
    return;
  }
  out[0] = fmad(a[1], b[2], -(a[2] * b[1]));
  out[1] = fmad(a[2], b[0], -(a[0] * b[2]));
  out[2] = fmad(a[0], b[1], -(a[1] * b[0]));
}

/** sin and cos from +,-,*,fma and rint only, so that a GPU implementation of the same sequence
 *  gives the same bits (libm's sin/cos differ between glibc and CUDA by up to 1-2 ulp):
 *  Cody-Waite reduction by pi/2 in three parts, fdlibm's degree-13/14 minimax kernels on
 *  [-pi/4, pi/4], quadrant fix-up.  Error vs the true value is ~1 ulp for |x| < 1e5. */
inline void sincos_canon(double x, double * s_out, double * c_out)
{
  if(kTextbook)
  {
    *s_out = std::sin(x);
    *c_out = std::cos(x);
    return;
  }
  const double j = std::nearbyint(x * 6.36619772367581382433e-01);
  double r = fmad(-j, 1.57079632673412561417e+00, x);
  r = fmad(-j, 6.07710050650619224932e-11, r);
  r = fmad(-j, 2.02226624879595063154e-21, r);
  const double z = r * r;
  double ps = 1.58969099521155010221e-10;
  ps = fmad(ps, z, -2.50507602534068634195e-08);
  ps = fmad(ps, z, 2.75573137070700676789e-06);
  ps = fmad(ps, z, -1.98412698298579493134e-04);
  ps = fmad(ps, z, 8.33333333332248946124e-03);
  ps = fmad(ps, z, -1.66666666666666324348e-01);
  const double sr = fmad(r * z, ps, r);
  double pc = -1.13596475577881948265e-11;
  pc = fmad(pc, z, 2.08757232129817482790e-09);
  pc = fmad(pc, z, -2.75573143513906633035e-07);
  pc = fmad(pc, z, 2.48015872894767294178e-05);
  pc = fmad(pc, z, -1.38888888888741095749e-03);
  pc = fmad(pc, z, 4.16666666666666019037e-02);
  const double cr = fmad(z * z, pc, fmad(-0.5, z, 1.0));
  const long long q = static_cast<long long>(j) & 3;
  *s_out = q == 0 ? sr : q == 1 ? cr : q == 2 ? -sr : -cr;
  *c_out = q == 0 ? cr : q == 1 ? -sr : q == 2 ? -cr : sr;
}

inline double clampd(double v, double lo, double hi)
{
  // cwiseMax(lo).cwiseMin(hi)
  double r = v < lo ? lo : v;
  return r > hi ? hi : r;
}
} // namespace oracle
