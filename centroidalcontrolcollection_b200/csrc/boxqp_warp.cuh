// boxqp_warp.cuh — one-warp projected-Newton box QP and the per-stage Cholesky (square-root-free,
// L D L') it shares with the DDP backward pass.  Lane i owns input i (m <= 32): row i of the Hessian lives in 32 FP64
// registers, the Cholesky factor lives in the warp's shared-memory tile.
//
// Replaces: nmpc_ddp::BoxQP<InputDim>::solve as reached from the reference through
// ddp_solver_->config().with_input_constraint = true (reference src/DdpCentroidal.cpp:197,
// src/DdpSingleRigidBody.cpp:267); algorithm = Tassa et al. ICRA 2014 boxQP.m.
// Evaluation order = oracle/boxqp.hpp (DESIGN.md §4 "canonical arithmetic"): sequential fma
// chains over the input index, pairwise-tree warp sums, IEEE / and sqrt — bit-exact vs oracle.
//
// Code size is a first-class constraint here (profiles/r01_ncu_v1_icache.md): the first
// version unrolled the factorisation over register-indexed columns and inlined the
// matrix-vector product at six sites; the kernel spent 72 % of its issue slots waiting for
// instruction fetch.  This version keeps every hot loop either rolled or single-site:
//  * the free block is compacted (rank r <- r-th free input), so the factorisation is the
//    oracle's dense nf x nf L D L' with no skipped columns;
//  * the column loop is rolled: after column k every lane shifts its register row by one
//    (folded into the update fma), so the live column is always register 0;
//  * BoxQP has one objective-evaluation site, driven by a small phase variable;
//  * division and square root are out-of-line calls.
#pragma once
#include "warp_ctx.cuh"

namespace ccc
{
constexpr int kLda = 34; // row stride of the 32x32 tile A (even: 16-byte aligned rows)

struct d2
{
  double x, y;
};
CCC_DEV d2 ld2(const double * p)
{
#ifdef CCC_WARP_EMU
  return d2{p[0], p[1]};
#else
  double2 v = *reinterpret_cast<const double2 *>(p);
  return d2{v.x, v.y};
#endif
}

CCC_DEV void st2(double * p, double x, double y)
{
#ifdef CCC_WARP_EMU
  p[0] = x;
  p[1] = y;
#else
  *reinterpret_cast<double2 *>(p) = make_double2(x, y);
#endif
}

/** IEEE division (inline: an out-of-line routine far from its callers costs instruction-cache
 *  locality in the hot loops). */
CCC_DEV double ddiv(double a, double b)
{
  return a / b;
}

struct BoxQpCfg
{
  int max_iter;
  double grad_thre, rel_improve_thre, step_factor, min_step, armijo;
};

/** y_lane = sum_j H[j] * vb[j] as four interleaved fma chains (slot j % 4, ascending j, from
 *  +0.0) combined as (s0 + s1) + (s2 + s3): the oracle's dot4 (oracle/num.hpp).  vb: 32 doubles
 *  in smem; entries j >= m are +0.0 and H[j >= m] is finite, so only the 16-boundary is tested. */
CCC_DEV double matvec32(const double (&H)[32], const double * vb, int m)
{
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  CCC_UNROLL
  for(int q = 0; q < 8; q++)
  {
    if(q == 4 && m <= 16) break;
    d2 v01 = ld2(vb + 4 * q), v23 = ld2(vb + 4 * q + 2);
    a0 = dfma(H[4 * q], v01.x, a0);
    a1 = dfma(H[4 * q + 1], v01.y, a1);
    a2 = dfma(H[4 * q + 2], v23.x, a2);
    a3 = dfma(H[4 * q + 3], v23.y, a3);
  }
  return (a0 + a1) + (a2 + a3);
}

/** Publish one value per lane into a 32-double smem vector (inactive lanes publish +0.0). */
CCC_DEV void publish(double * vb, double v, bool active)
{
  warp_sync(); // earlier readers of vb are done
  vb[lane_id()] = active ? v : 0.0;
  warp_sync();
}

/** Reload row `lane` of the symmetric matrix from the full tile S (both triangles stored, row stride kLda:
 *  16 LDS.128).  No guards: entries outside the m x m block are whatever finite values the tile holds (it is
 *  zero-filled once per solve and only ever receives finite data); they meet +0.0 in matvec32. */
CCC_DEV void load_sym_row(double (&H)[32], const double * S, int m)
{
  const double * row = S + lane_id() * kLda;
  CCC_UNROLL
  for(int q = 0; q < 16; q++)
  {
    if(q == 8 && m <= 16) break;
    const d2 v = ld2(row + 2 * q);
    H[2 * q] = v.x;
    H[2 * q + 1] = v.y;
  }
}

/** Compact numbering of the free inputs (the oracle's free_idx list). */
struct FreeSet
{
  unsigned free_mask; // bit j = input j is free
  int nf;             // number of free inputs
  int idx;            // original index of compact row `lane` (valid for lane < nf)
  int rank;           // compact position of original input `lane` (valid if free)
};

CCC_DEV int popc32(unsigned v)
{
#ifdef CCC_WARP_EMU
  return __builtin_popcount(v);
#else
  return __popc(v);
#endif
}

/** idxbuf: 32 ints of smem receiving the free list (slots >= nf are 0). */
CCC_DEV FreeSet make_free_set(unsigned clamped, int m, int * idxbuf)
{
  const int lane = lane_id();
  const unsigned active_mask = m >= 32 ? 0xffffffffu : ((1u << m) - 1u);
  FreeSet fs;
  fs.free_mask = ~clamped & active_mask;
  fs.nf = popc32(fs.free_mask);
  fs.rank = popc32(fs.free_mask & ((1u << lane) - 1u));
  warp_sync();
  if(lane >= fs.nf) idxbuf[lane] = 0;
  if((fs.free_mask >> lane) & 1u) idxbuf[fs.rank] = lane; // inverse of rank: no find-nth-set-bit
  warp_sync();
  fs.idx = idxbuf[lane];
  return fs;
}

/** Gather compact row `lane` of H[free,free] from the full symmetric tile S: row idx(lane), columns idx(c)
 *  (unguarded: lanes and columns >= nf pick up finite tile entries that the factorisation never lets through). */
CCC_DEV void load_compact_row(double (&Hc)[32], const double * S, const int * idxbuf, const FreeSet & fs)
{
  const double * row = S + fs.idx * kLda;
  CCC_UNROLL
  for(int c = 0; c < 32; c++)
  {
    if((c & 7) == 0 && c >= fs.nf && c > 0) break;
    Hc[c] = row[idxbuf[c]];
  }
}

constexpr int kCbStride = 68; // one column buffer: 32 values (+1 shift) + zero tail up to index 63

/** Square-root-free Cholesky (L D L', L unit lower) of the compact nf x nf block, rows in
 *  registers, column loop rolled.  In: Hc = compact row `lane`.  Out: Hc destroyed; A's strict
 *  lower triangle holds the compact L (row r > col c at A[r][c]); invd_c = 1 / d_lane (compact
 *  numbering).  cb: 2 * kCbStride doubles of smem.  The *unscaled* column k (entries a_ik = L_ik d_k)
 *  is published at offset (k & 1) of buffer (k & 1), so that the trailing-update reads start at an
 *  even (16-byte aligned) index and pair up; publishing the unscaled column keeps the publish and
 *  the barrier off the pivot -> reciprocal -> scale -> update critical path.
 *  rhs_c (compact numbering): in b, out y = L^-1 b (not yet scaled by D^-1) — the forward
 *  substitution is carried along column by column in the oracle's order
 *  (acc_i = fma(-L[i][k], y_k, acc_i), k ascending).
 *  Returns false (warp-uniform) if a pivot is not > 0.  Operation order: oracle FreeLlt::compute. */
CCC_DEV bool llt_factor_compact(double (&Hc)[32], double * A, double * cb, int nf, double & invd_c, double & rhs_c)
{
  const int lane = lane_id();
  bool ok = true;
  double rhs = lane < nf ? rhs_c : 0.0;
  double y = 0.0;
  warp_sync(); // earlier readers of cb and of A's lower triangle are done
  cb[32 + lane] = 0.0;                 // buffer 0: indices 32..63
  cb[kCbStride + 33 + lane] = 0.0;     // buffer 1: indices 33..64 (index 32 is lane 31's slot)
  CCC_NOUNROLL
  for(int k = 0; k < nf; k++)
  {
    const bool below = lane > k && lane < nf;
    const double c = below ? Hc[0] : 0.0; // unscaled column entry a_ik
    const int odd = k & 1;
    double * cbuf = cb + odd * kCbStride;
    cbuf[lane + odd] = c;
    warp_sync(); // the one barrier of a column; nothing below it waits on another lane's store
    const double * col = cbuf + k + odd; // even index: col[2q], col[2q+1] load as one LDS.128
    const double c1 = col[1];            // feeds the next pivot: issued before the reciprocal chain
    const double piv = warp_shfl(Hc[0], k);
    const double yk = warp_shfl(rhs, k); // unit L: y_k is the accumulated right-hand side itself
    ok = piv > 0.0; // a failed pivot poisons this column's update only; the loop exits below
    const double inv = drcp(piv);
    if(lane == k)
    {
      invd_c = inv;
      y = yk;
    }
    const double l = c * inv;
    if(below)
    {
      rhs = dfma(-l, yk, rhs);
      A[lane * kLda + k] = l;
    }
    // trailing update folded with a shift by one register: entry (lane, k+j) moves to slot j-1
    const int rem = nf - k;
    Hc[0] = dfma(-l, c1, Hc[1]);
    CCC_UNROLL
    for(int j = 2; j < 32; j += 2)
    {
      if((j & 7) == 2 && j >= rem) break;
      const d2 c2 = ld2(col + j);
      Hc[j - 1] = dfma(-l, c2.x, Hc[j]);
      if(j + 1 < 32) Hc[j] = dfma(-l, c2.y, Hc[j + 1]);
    }
    if(!ok) break;
  }
  warp_sync();
  rhs_c = y;
  return ok;
}

/** Back substitution L' x = D^-1 y on the compact block (y, x in compact numbering, lane r). */
CCC_DEV double llt_back_compact(double y_c, const double * A, int nf, double invd_c)
{
  const int lane = lane_id();
  const bool in = lane < nf;
  double acc = in ? y_c * invd_c : 0.0;
  CCC_UNROLL_N(2)
  for(int j = nf - 1; j >= 0; j--)
  {
    const double xj = warp_shfl(acc, j);
    if(in && lane < j) acc = dfma(-A[j * kLda + lane], xj, acc);
  }
  return acc;
}

/** Forward substitution L y = b on the compact block (used when the factor is reused); y is not
 *  yet scaled by D^-1, like the one llt_factor_compact carries along. */
CCC_DEV double llt_fwd_compact(double rhs_c, const double * A, int nf)
{
  const int lane = lane_id();
  const bool in = lane < nf;
  double acc = in ? rhs_c : 0.0;
  CCC_UNROLL_N(2)
  for(int j = 0; j < nf; j++)
  {
    const double yj = warp_shfl(acc, j);
    if(in && lane > j) acc = dfma(-A[lane * kLda + j], yj, acc);
  }
  return acc;
}

/** (L D L') x = r for NR right-hand sides per lane (compact numbering).  The row that becomes final in a step is
 *  handed to the other lanes through shared memory (NR/2 STS.128 by its owner, one warp barrier, NR/2 broadcast
 *  LDS.128) instead of 2 NR shuffles; `xb`: 2 * (NR + NR % 2) doubles, two buffers used alternately so that one
 *  barrier per step is enough.  Same values, same order of operations as the shuffle form. */
template<int NR>
CCC_DEV void llt_solve_compactN(double (&r)[NR], const double * A, double * xb, int nf, double invd_c)
{
  constexpr int NRP = NR + (NR & 1);
  const int lane = lane_id();
  const bool in = lane < nf;
  CCC_UNROLL
  for(int c = 0; c < NR; c++) r[c] = in ? r[c] : 0.0;
  int flip = 0;
  warp_sync(); // earlier users of xb are done
  CCC_NOUNROLL
  for(int j = 0; j < nf; j++)
  {
    double * buf = xb + flip * NRP;
    flip ^= 1;
    if(lane == j)
    {
      CCC_UNROLL
      for(int c = 0; c < NR; c += 2) st2(buf + c, r[c], c + 1 < NR ? r[c + 1] : 0.0);
    }
    warp_sync();
    const bool upd = in && lane > j;
    const double lij = upd ? A[lane * kLda + j] : 0.0;
    CCC_UNROLL
    for(int c = 0; c < NR; c += 2)
    {
      const d2 y = ld2(buf + c);
      if(upd) r[c] = dfma(-lij, y.x, r[c]);
      if(c + 1 < NR && upd) r[c + 1] = dfma(-lij, y.y, r[c + 1]);
    }
  }
  CCC_UNROLL
  for(int c = 0; c < NR; c++) r[c] = in ? r[c] * invd_c : 0.0;
  CCC_NOUNROLL
  for(int j = nf - 1; j >= 0; j--)
  {
    double * buf = xb + flip * NRP;
    flip ^= 1;
    if(lane == j)
    {
      CCC_UNROLL
      for(int c = 0; c < NR; c += 2) st2(buf + c, r[c], c + 1 < NR ? r[c + 1] : 0.0);
    }
    warp_sync();
    const bool upd = in && lane < j;
    const double lji = upd ? A[j * kLda + lane] : 0.0;
    CCC_UNROLL
    for(int c = 0; c < NR; c += 2)
    {
      const d2 x = ld2(buf + c);
      if(upd) r[c] = dfma(-lji, x.x, r[c]);
      if(c + 1 < NR && upd) r[c + 1] = dfma(-lji, x.y, r[c + 1]);
    }
  }
  warp_sync();
}

struct BoxQpOut
{
  int retval;       // boxQP.m result code
  unsigned clamped; // bit j = input j clamped
  FreeSet fs;       // compact numbering matching the factor left in A (unless all clamped)
  double invd_c;    // 1 / d_r of that factor (L D L'), compact numbering
  int iters, nfactor, ls_steps;
};

/** One-warp BoxQP.  H: row `lane` of the symmetric Hessian, also stored in the upper triangle
 *  in full in the smem tile S (both triangles); on return H is intact again and A's strict lower triangle
 *  holds the compact factor of the final free block.  x (in: start point, out: solution), g,
 *  lo, hi: one value per lane.  vb: 2 * kCbStride + 32 doubles of smem (column buffers + publish
 *  vector); idxbuf: 32 ints. */
CCC_DEV BoxQpOut boxqp_warp(double (&H)[32],
                            const double * S,
                            double * A,
                            double * vb,
                            int * idxbuf,
                            double g,
                            double lo,
                            double hi,
                            double & x,
                            int m,
                            const BoxQpCfg & cfg)
{
  const int lane = lane_id();
  const bool active = lane < m;
  const unsigned active_mask = m >= 32 ? 0xffffffffu : ((1u << m) - 1u);
  double * pub = vb + 2 * kCbStride;
  BoxQpOut out;
  out.retval = 0;
  out.clamped = 0;
  out.invd_c = 1.0;
  out.nfactor = 0;
  out.ls_steps = 0;
  out.iters = 0;
  out.fs.free_mask = 0;
  out.fs.nf = 0;
  out.fs.idx = 0;
  out.fs.rank = 0;
  x = clampd(x, lo, hi);
  double xc = x, Hx = 0.0, obj = 0.0, old_obj = 0.0;
  double step = 1.0, sdotg = -1.0, search = 0.0;
  unsigned old_clamped = 0;
  int phase = 0; // 0: initial objective, 1: first Armijo trial, 2: later Armijo trials
  int iter = 0;
  CCC_NOUNROLL
  for(;;)
  {
    // ---- the single objective-evaluation site:  0.5 xc'H xc + g'xc  ----
    publish(pub, xc, active);
    const double Hxc = matvec32(H, pub, m);
    const double objc = warp_sum(active ? xc * dfma(0.5, Hxc, g) : 0.0); // one tree sum (oracle/boxqp.hpp objective)
    if(phase != 0)
    {
      const bool ls_fail = (phase == 2 && step < cfg.min_step);
      if(!ls_fail && (objc - old_obj) > cfg.armijo * (step * sdotg)) // Armijo ratio test, cross-multiplied (oracle/boxqp.hpp)
      {
        step = step * cfg.step_factor;
        out.ls_steps++;
        xc = clampd(dfma(step, search, x), lo, hi);
        phase = 2;
        continue;
      }
      x = xc;
      Hx = Hxc;
      obj = objc;
      if(ls_fail)
      {
        out.retval = 2;
        break;
      }
      if(iter >= cfg.max_iter)
      {
        out.retval = 1;
        break;
      }
    }
    else
    {
      Hx = Hxc;
      obj = objc;
      old_obj = objc;
    }
    iter++;
    out.iters = iter;
    // ---- top of a projected-Newton iteration ----
    if(iter > 1 && (old_obj - obj) < cfg.rel_improve_thre * dabs(old_obj))
    {
      out.retval = 4;
      break;
    }
    old_obj = obj;
    const double grad = g + Hx;
    const bool cl = active && ((x == lo && grad > 0) || (x == hi && grad < 0));
    const unsigned clamped = warp_ballot(cl) & active_mask;
    if(clamped == active_mask)
    {
      out.clamped = clamped;
      out.retval = 6;
      break;
    }
    // g + H (x .* clamped): needs H in registers, so it is evaluated before the factorisation
    // borrows them (it does not depend on the factor)
    publish(pub, x, cl);
    const double gc = g + matvec32(H, pub, m);
    double y_c;
    if(iter == 1 || clamped != old_clamped)
    {
      out.clamped = clamped;
      out.fs = make_free_set(clamped, m, idxbuf);
      load_compact_row(H, S, idxbuf, out.fs);
      y_c = warp_shfl(gc, out.fs.idx);
      const bool ok = llt_factor_compact(H, A, vb, out.fs.nf, out.invd_c, y_c);
      load_sym_row(H, S, m);
      if(!ok)
      {
        out.retval = -1;
        break;
      }
      out.nfactor++;
    }
    else
    {
      y_c = llt_fwd_compact(warp_shfl(gc, out.fs.idx), A, out.fs.nf);
    }
    old_clamped = clamped;
    const bool free_i = active && !cl;
    const double gnorm2 = warp_sum(free_i ? grad * grad : 0.0);
    if(gnorm2 < cfg.grad_thre * cfg.grad_thre) // squared form of boxQP.m's norm(grad(free)) < minGrad
    {
      out.retval = 5;
      break;
    }
    const double sol_c = llt_back_compact(y_c, A, out.fs.nf, out.invd_c);
    const double sol = warp_shfl(sol_c, out.fs.rank);
    search = free_i ? (-sol) - x : 0.0;
    sdotg = warp_sum(active ? search * grad : 0.0);
    if(sdotg >= 0)
    {
      out.retval = 0;
      break;
    }
    step = 1.0;
    xc = clampd(dfma(step, search, x), lo, hi);
    phase = 1;
  }
  return out;
}
} // namespace ccc
