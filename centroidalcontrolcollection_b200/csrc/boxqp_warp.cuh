// boxqp_warp.cuh — one-warp projected-Newton box QP and the per-stage LL^T it shares with the
// DDP backward pass.  Lane i owns input i (m <= 32): row i of the Hessian lives in 32 FP64
// registers, the Cholesky factor lives in the warp's shared-memory tile.
//
// Replaces: nmpc_ddp::BoxQP<InputDim>::solve as reached from the reference through
// ddp_solver_->config().with_input_constraint = true (reference src/DdpCentroidal.cpp:197,
// src/DdpSingleRigidBody.cpp:267); algorithm = Tassa et al. ICRA 2014 boxQP.m.
// Evaluation order = oracle/boxqp.hpp (DESIGN.md §4 "canonical arithmetic"): sequential fma
// chains over the input index, pairwise-tree warp sums, IEEE / and sqrt — bit-exact vs oracle.
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

struct BoxQpCfg
{
  int max_iter;
  double grad_thre, rel_improve_thre, step_factor, min_step, armijo;
};

/** y_lane = sum_j H[j] * vb[j], ascending j, one fma chain from +0.0 (vb: 32 doubles in smem). */
CCC_DEV double matvec32(const double (&H)[32], const double * vb, int m)
{
  double acc = 0.0;
  CCC_UNROLL
  for(int jj = 0; jj < 16; jj++)
  {
    if(2 * jj >= m) break;
    d2 v = ld2(vb + 2 * jj);
    acc = dfma(H[2 * jj], v.x, acc);
    if(2 * jj + 1 < m) acc = dfma(H[2 * jj + 1], v.y, acc);
  }
  return acc;
}

/** Publish one value per lane into a 32-double smem vector (inactive lanes publish +0.0). */
CCC_DEV void publish(double * vb, double v, bool active)
{
  warp_sync(); // earlier readers of vb are done
  vb[lane_id()] = active ? v : 0.0;
  warp_sync();
}

/** Reload row `lane` of the symmetric matrix stored in the upper triangle (+diag) of A. */
CCC_DEV void load_sym_row(double (&H)[32], const double * A, int m)
{
  const int lane = lane_id();
  CCC_UNROLL
  for(int j = 0; j < 32; j++)
  {
    if(j >= m) break;
    int r = j < lane ? j : lane, c = j < lane ? lane : j;
    H[j] = A[r * kLda + c];
  }
}

/** In-register right-looking LL^T of the free block (clamped rows/columns behave as identity).
 *  In:  H = row `lane` of the matrix.  Out: H destroyed; strict lower triangle of A holds L for
 *  the free rows/columns; invd = 1 / L[lane][lane].  cb0/cb1: two 32-double smem vectors.
 *  Returns false (warp-uniform) if a pivot is not > 0. */
CCC_DEV bool llt_factor(double (&H)[32], double * A, double * cb0, double * cb1, unsigned clamped, int m, double & invd)
{
  const int lane = lane_id();
  const bool free_i = lane < m && !((clamped >> lane) & 1u);
  bool ok = true;
  int par = 0; // alternates per processed column: a buffer is rewritten only two syncs later
  warp_sync(); // earlier readers of cb0/cb1 and of A's lower triangle are done
  CCC_UNROLL
  for(int k = 0; k < 32; k++)
  {
    if(k >= m) break;
    if((clamped >> k) & 1u) continue;
    double piv = warp_shfl(H[k], k);
    if(!(piv > 0.0))
    {
      ok = false;
      break;
    }
    double d = dsqrt(piv);
    double inv = 1.0 / d;
    if(lane == k) invd = inv;
    const bool below = free_i && lane > k;
    double l = below ? H[k] * inv : 0.0;
    double * cb = par ? cb1 : cb0;
    par ^= 1;
    cb[lane] = l;
    if(below) A[lane * kLda + k] = l;
    warp_sync();
    CCC_UNROLL
    for(int j = k + 1; j < 32; j++)
    {
      if(j >= m) break;
      H[j] = dfma(-l, cb[j], H[j]);
    }
  }
  warp_sync();
  return ok;
}

/** Solve (L L') s = rhs on the free block; one value per lane (clamped/inactive lanes: 0). */
CCC_DEV double llt_solve1(double rhs, const double * A, unsigned clamped, int m, double invd)
{
  const int lane = lane_id();
  const bool free_i = lane < m && !((clamped >> lane) & 1u);
  double acc = free_i ? rhs : 0.0;
  for(int j = 0; j < m; j++)
  {
    if((clamped >> j) & 1u) continue;
    double yj = warp_shfl(acc * invd, j);
    if(free_i && lane > j) acc = dfma(-A[lane * kLda + j], yj, acc);
  }
  acc = free_i ? acc * invd : 0.0;
  for(int j = m - 1; j >= 0; j--)
  {
    if((clamped >> j) & 1u) continue;
    double xj = warp_shfl(acc * invd, j);
    if(free_i && lane < j) acc = dfma(-A[j * kLda + lane], xj, acc);
  }
  return free_i ? acc * invd : 0.0;
}

/** Same for NR right-hand sides held per lane (row `lane` of an m x NR matrix). */
template<int NR>
CCC_DEV void llt_solveN(double (&r)[NR], const double * A, unsigned clamped, int m, double invd)
{
  const int lane = lane_id();
  const bool free_i = lane < m && !((clamped >> lane) & 1u);
  CCC_UNROLL
  for(int c = 0; c < NR; c++) r[c] = free_i ? r[c] : 0.0;
  for(int j = 0; j < m; j++)
  {
    if((clamped >> j) & 1u) continue;
    const bool upd = free_i && lane > j;
    double lij = upd ? A[lane * kLda + j] : 0.0;
    CCC_UNROLL
    for(int c = 0; c < NR; c++)
    {
      double yj = warp_shfl(r[c] * invd, j);
      if(upd) r[c] = dfma(-lij, yj, r[c]);
    }
  }
  CCC_UNROLL
  for(int c = 0; c < NR; c++) r[c] = free_i ? r[c] * invd : 0.0;
  for(int j = m - 1; j >= 0; j--)
  {
    if((clamped >> j) & 1u) continue;
    const bool upd = free_i && lane < j;
    double lji = upd ? A[j * kLda + lane] : 0.0;
    CCC_UNROLL
    for(int c = 0; c < NR; c++)
    {
      double xj = warp_shfl(r[c] * invd, j);
      if(upd) r[c] = dfma(-lji, xj, r[c]);
    }
  }
  CCC_UNROLL
  for(int c = 0; c < NR; c++) r[c] = free_i ? r[c] * invd : 0.0;
}

struct BoxQpOut
{
  int retval;       // boxQP.m result code
  unsigned clamped; // bit j = input j clamped (matches the factor left in A unless all clamped)
  double invd;      // 1 / L[lane][lane] of that factor
  int iters, nfactor, ls_steps;
};

/** 0.5 x'Hx + g'x with x published through vb; also returns Hx (row `lane`). */
CCC_DEV double boxqp_objective(const double (&H)[32], double g, double x, double * vb, int m, bool active, double & Hx)
{
  publish(vb, x, active);
  Hx = matvec32(H, vb, m);
  double t1 = active ? x * Hx : 0.0;
  double t2 = active ? x * g : 0.0;
  return dfma(0.5, warp_sum(t1), warp_sum(t2));
}

/** One-warp BoxQP.  H: row `lane` of the symmetric Hessian, also stored in the upper triangle
 *  (+diagonal) of the smem tile A; on return H is intact again and A's strict lower triangle
 *  holds the factor of the final free block.  x (in: start point, out: solution), g, lo, hi:
 *  one value per lane.  vb0..vb2: three 32-double smem vectors. */
CCC_DEV BoxQpOut boxqp_warp(double (&H)[32],
                            double * A,
                            double * vb0,
                            double * vb1,
                            double * vb2,
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
  BoxQpOut out;
  out.retval = 0;
  out.clamped = 0;
  out.invd = 1.0;
  out.nfactor = 0;
  out.ls_steps = 0;
  x = clampd(x, lo, hi);
  double Hx;
  double obj = boxqp_objective(H, g, x, vb2, m, active, Hx);
  double old_obj = obj;
  unsigned old_clamped = 0;
  int iter = 1;
  for(;; iter++)
  {
    out.iters = iter;
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
    // gradient with the free part of x removed: g + H (x .* clamped).  Needs H in registers, so it
    // is evaluated before the factorisation overwrites them (it does not depend on the factor).
    publish(vb2, x, cl);
    const double gc = g + matvec32(H, vb2, m);
    if(iter == 1 || clamped != old_clamped)
    {
      bool ok = llt_factor(H, A, vb0, vb1, clamped, m, out.invd);
      load_sym_row(H, A, m);
      out.clamped = clamped;
      if(!ok)
      {
        out.retval = -1;
        break;
      }
      out.nfactor++;
    }
    old_clamped = clamped;
    out.clamped = clamped;
    const bool free_i = active && !cl;
    const double gnorm = dsqrt(warp_sum(free_i ? grad * grad : 0.0));
    if(gnorm < cfg.grad_thre)
    {
      out.retval = 5;
      break;
    }
    const double sol = llt_solve1(gc, A, clamped, m, out.invd);
    const double search = free_i ? (-sol) - x : 0.0;
    const double sdotg = warp_sum(active ? search * grad : 0.0);
    if(sdotg >= 0)
    {
      out.retval = 0;
      break;
    }
    double step = 1.0;
    double xc = clampd(dfma(step, search, x), lo, hi);
    double Hxc;
    double objc = boxqp_objective(H, g, xc, vb2, m, active, Hxc);
    bool ls_fail = false;
    while((objc - old_obj) / (step * sdotg) < cfg.armijo)
    {
      step = step * cfg.step_factor;
      out.ls_steps++;
      xc = clampd(dfma(step, search, x), lo, hi);
      objc = boxqp_objective(H, g, xc, vb2, m, active, Hxc);
      if(step < cfg.min_step)
      {
        ls_fail = true;
        break;
      }
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
  return out;
}
} // namespace ccc
