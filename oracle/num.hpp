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

/** Sequential fma chain: sum_i a[i*sa] * b[i*sb], ascending i, starting from +0.0. */
inline double dot_seq(const double * a, int sa, const double * b, int sb, int n)
{
  double acc = 0.0;
  for(int i = 0; i < n; i++) acc = std::fma(a[i * sa], b[i * sb], acc);
  return acc;
}

/** Four interleaved fma chains (term i goes to chain i % 4, ascending i, each from +0.0),
 *  combined as (s0 + s1) + (s2 + s3).  Used for the matrix-vector products over the input index
 *  (H x in BoxQP, Quu k in the backward pass): the usual 4-way unrolled dot product. */
inline double dot4(const double * a, int sa, const double * b, int sb, int n)
{
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for(int i = 0; i < n; i++) s[i & 3] = std::fma(a[i * sa], b[i * sb], s[i & 3]);
  return (s[0] + s[1]) + (s[2] + s[3]);
}

/** Pairwise tree over 32 zero-padded leaves: level strides 16, 8, 4, 2, 1. */
inline double tree_sum32(const double * p, int m)
{
  double t[32];
  for(int i = 0; i < 32; i++) t[i] = i < m ? p[i] : 0.0;
  for(int off = 16; off >= 1; off >>= 1)
    for(int i = 0; i < off; i++) t[i] = t[i] + t[i + off];
  return t[0];
}

/** Cross product with one rounded product and one fma per component. */
inline void cross3(const double * a, const double * b, double * out)
{
  out[0] = std::fma(a[1], b[2], -(a[2] * b[1]));
  out[1] = std::fma(a[2], b[0], -(a[0] * b[2]));
  out[2] = std::fma(a[0], b[1], -(a[1] * b[0]));
}

inline double clampd(double v, double lo, double hi)
{
  // cwiseMax(lo).cwiseMin(hi)
  double r = v < lo ? lo : v;
  return r > hi ? hi : r;
}
} // namespace oracle
