// oracle/boxqp.hpp — CPU restatement of nmpc_ddp::BoxQP (projected-Newton box QP).
//
// TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: nmpc_ddp (isri-aist/NMPC, un-versioned
// dependency, reference CMakeLists.txt:25) is not in /root/reference, so this restates its
// published algorithm — Tassa, Mansard, Todorov, "Control-limited differential dynamic
// programming", ICRA 2014, boxQP.m — which nmpc_ddp ports.  It is reached from the reference
// through ddp_solver_->config().with_input_constraint = true (src/DdpCentroidal.cpp:197,
// src/DdpSingleRigidBody.cpp:267).
//
//   min 0.5 x'Hx + g'x   s.t.  lo <= x <= hi
//
// result codes (boxQP.m): -1 Hessian not PD, 0 no descent direction, 1 max iterations,
// 2 max line-search iterations, 4 small improvement, 5 small gradient, 6 all clamped.
#pragma once
#include "num.hpp"

namespace oracle
{
struct BoxQpConfig
{
  int max_iter = 500;
  double grad_thre = 1e-8;
  double rel_improve_thre = 1e-8;
  double step_factor = 0.6;
  double min_step = 1e-22;
  double armijo = 0.1;
};

/** Square-root-free Cholesky factor H[free,free] = L D L' of the free block (compact storage,
 *  row-major nf x nf, L unit lower).  Eigen::LLT (what nmpc_ddp::BoxQP calls) and L D L' are the
 *  same factorisation up to where the pivot's square root is taken: the positive-definiteness
 *  test (pivot > 0), the solution and the free/clamped logic are unchanged; only the rounding of
 *  individual entries differs (canonical arithmetic).  It is the form the engine
 *  runs because it keeps the square root and the column publish off the per-column critical path. */
struct FreeLlt
{
  int nf = 0;
  std::vector<int> idx;     // free indices, ascending
  std::vector<double> L;    // nf*nf, unit lower (strict part stored)
  std::vector<double> C;    // nf*nf, lower: C[i][k] = L[i][k] * d_k, the unscaled column entries (diagonal: d_k)
  std::vector<double> invd; // 1 / d_k

  /** Left-looking L D L' in the engine's operation order: entry (i,k) accumulates
   *  fma(-L[i][j], C[k][j], .) over j ascending; returns false if a pivot is not > 0. */
  bool compute(const double * H, int ld, const std::vector<int> & free_idx)
  {
    idx = free_idx;
    nf = static_cast<int>(idx.size());
    L.assign(static_cast<size_t>(nf) * nf, 0.0);
    C.assign(static_cast<size_t>(nf) * nf, 0.0);
    invd.assign(nf, 0.0);
    if(kTextbook)
    {
      // Eigen::LLT<Lower>, unblocked: L[k][k] = sqrt(pivot), column divided by it
      for(int k = 0; k < nf; k++)
      {
        double acc = H[idx[k] * ld + idx[k]];
        for(int j = 0; j < k; j++) acc = acc - L[k * nf + j] * L[k * nf + j];
        if(!(acc > 0.0)) return false;
        const double d = std::sqrt(acc);
        L[k * nf + k] = d;
        for(int i = k + 1; i < nf; i++)
        {
          double a = H[idx[i] * ld + idx[k]];
          for(int j = 0; j < k; j++) a = a - L[i * nf + j] * L[k * nf + j];
          L[i * nf + k] = a / d;
        }
      }
      return true;
    }
    for(int k = 0; k < nf; k++)
    {
      double d = H[idx[k] * ld + idx[k]];
      for(int j = 0; j < k; j++) d = fmad(-L[k * nf + j], C[k * nf + j], d);
      if(!(d > 0.0)) return false;
      const double inv = 1.0 / d;
      C[k * nf + k] = d;
      invd[k] = inv;
      for(int i = k + 1; i < nf; i++)
      {
        double a = H[idx[i] * ld + idx[k]]; // lower triangle of H, like Eigen::LLT<Lower>
        for(int j = 0; j < k; j++) a = fmad(-L[i * nf + j], C[k * nf + j], a);
        C[i * nf + k] = a;
        L[i * nf + k] = a * inv;
      }
    }
    return true;
  }

  /** Solve (L D L') y = b in place on a compact vector of length nf (stride sb). */
  void solve(double * b, int sb) const
  {
    if(kTextbook)
    {
      for(int i = 0; i < nf; i++)
      {
        double acc = b[i * sb];
        for(int j = 0; j < i; j++) acc = acc - L[i * nf + j] * b[j * sb];
        b[i * sb] = acc / L[i * nf + i];
      }
      for(int i = nf - 1; i >= 0; i--)
      {
        double acc = b[i * sb];
        for(int j = nf - 1; j > i; j--) acc = acc - L[j * nf + i] * b[j * sb];
        b[i * sb] = acc / L[i * nf + i];
      }
      return;
    }
    for(int i = 0; i < nf; i++)
    {
      double acc = b[i * sb];
      for(int j = 0; j < i; j++) acc = fmad(-L[i * nf + j], b[j * sb], acc);
      b[i * sb] = acc;
    }
    for(int i = 0; i < nf; i++) b[i * sb] = b[i * sb] * invd[i];
    for(int i = nf - 1; i >= 0; i--)
    {
      double acc = b[i * sb];
      for(int j = nf - 1; j > i; j--) acc = fmad(-L[j * nf + i], b[j * sb], acc);
      b[i * sb] = acc;
    }
  }
};

/** Plain unblocked LL' with the same interface (the 3x3 inertia solves of the single-rigid-body
 *  model, reference src/DdpSingleRigidBody.cpp:70,146-167 `inertia.llt().solve`). */
struct DenseLlt
{
  int nf = 0;
  std::vector<int> idx;
  std::vector<double> L;    // nf*nf, lower
  std::vector<double> invd; // 1 / L[k][k]

  bool compute(const double * H, int ld, const std::vector<int> & free_idx)
  {
    idx = free_idx;
    nf = static_cast<int>(idx.size());
    L.assign(static_cast<size_t>(nf) * nf, 0.0);
    invd.assign(nf, 0.0);
    for(int k = 0; k < nf; k++)
    {
      double acc = H[idx[k] * ld + idx[k]];
      for(int j = 0; j < k; j++) acc = fmad(-L[k * nf + j], L[k * nf + j], acc);
      if(!(acc > 0.0)) return false;
      double d = std::sqrt(acc);
      double inv = 1.0 / d;
      L[k * nf + k] = d;
      invd[k] = inv;
      for(int i = k + 1; i < nf; i++)
      {
        double a = H[idx[i] * ld + idx[k]];
        for(int j = 0; j < k; j++) a = fmad(-L[i * nf + j], L[k * nf + j], a);
        L[i * nf + k] = a * inv;
      }
    }
    return true;
  }

  void solve(double * b, int sb) const
  {
    for(int i = 0; i < nf; i++)
    {
      double acc = b[i * sb];
      for(int j = 0; j < i; j++) acc = fmad(-L[i * nf + j], b[j * sb], acc);
      b[i * sb] = acc * invd[i];
    }
    for(int i = nf - 1; i >= 0; i--)
    {
      double acc = b[i * sb];
      for(int j = nf - 1; j > i; j--) acc = fmad(-L[j * nf + i], b[j * sb], acc);
      b[i * sb] = acc * invd[i];
    }
  }
};

struct BoxQp
{
  BoxQpConfig cfg;
  int retval = 0;
  int iters = 0;
  int nfactor = 0;
  int ls_steps = 0; // Armijo backtracking steps (statistics only)
  std::vector<double> x;
  std::vector<uint8_t> clamped;
  FreeLlt llt; // factor matching `clamped` (valid unless all clamped)

  static double objective(const double * H, int m, const double * g, const double * x, double * Hx)
  {
    // 0.5 x'Hx + g'x as ONE tree sum of x_i (0.5 (Hx)_i + g_i): the engine's evaluation is a chain of dependent
    // shuffle levels, and one reduction is a level shorter than two interleaved ones
    double t[32];
    if(kTextbook)
    {
      // x.dot(g) + 0.5 * x.dot(H * x)
      double xg = 0.0, xHx = 0.0;
      for(int i = 0; i < m; i++)
      {
        Hx[i] = dot4(H + i * m, 1, x, 1, m);
        xg = xg + x[i] * g[i];
        xHx = xHx + x[i] * Hx[i];
      }
      return xg + 0.5 * xHx;
    }
    for(int i = 0; i < m; i++)
    {
      Hx[i] = dot4(H + i * m, 1, x, 1, m);
      t[i] = x[i] * fmad(0.5, Hx[i], g[i]);
    }
    return tree_sum32(t, m);
  }

  /** H: m x m row-major.  Returns retval. */
  int solve(const double * H, const double * g, const double * lo, const double * hi, const double * x0, int m)
  {
    x.assign(m, 0.0);
    clamped.assign(m, 0);
    std::vector<uint8_t> old_clamped(m, 0);
    std::vector<double> Hx(m), grad(m), gc(m), search(m), xc(m), Hxc(m), tmp(m);
    for(int i = 0; i < m; i++) x[i] = clampd(x0[i], lo[i], hi[i]);
    double obj = objective(H, m, g, x.data(), Hx.data());
    double old_obj = obj;
    retval = 0;
    nfactor = 0;
    int iter = 1;
    for(;; iter++)
    {
      iters = iter;
      if(iter > 1 && (old_obj - obj) < cfg.rel_improve_thre * std::fabs(old_obj))
      {
        retval = 4;
        break;
      }
      old_obj = obj;

      // gradient (Hx is H*x of the current x)
      for(int i = 0; i < m; i++) grad[i] = g[i] + Hx[i];

      // clamped dimensions
      old_clamped = clamped;
      bool all_clamped = true, changed = false;
      std::vector<int> free_idx;
      for(int i = 0; i < m; i++)
      {
        clamped[i] = ((x[i] == lo[i] && grad[i] > 0) || (x[i] == hi[i] && grad[i] < 0)) ? 1 : 0;
        if(!clamped[i])
        {
          all_clamped = false;
          free_idx.push_back(i);
        }
        if(clamped[i] != old_clamped[i]) changed = true;
      }
      if(all_clamped)
      {
        retval = 6;
        break;
      }

      // factorize if the clamped set changed
      if(iter == 1 || changed)
      {
        if(!llt.compute(H, m, free_idx))
        {
          retval = -1;
          break;
        }
        nfactor++;
      }

      // gradient norm over the free dimensions
      for(int i = 0; i < m; i++) tmp[i] = clamped[i] ? 0.0 : grad[i] * grad[i];
      // boxQP.m: norm(grad(free)) < minGrad, tested on the squares (no square root on the iteration's chain)
      double gnorm2 = tree_sum32(tmp.data(), m);
      if(kTextbook ? std::sqrt(gnorm2) < cfg.grad_thre : gnorm2 < cfg.grad_thre * cfg.grad_thre)
      {
        retval = 5;
        break;
      }

      // search direction
      for(int i = 0; i < m; i++) xc[i] = clamped[i] ? x[i] : 0.0;
      for(int i = 0; i < m; i++) gc[i] = g[i] + dot4(H + i * m, 1, xc.data(), 1, m);
      {
        int nf = llt.nf;
        std::vector<double> rhs(nf);
        for(int a = 0; a < nf; a++) rhs[a] = gc[llt.idx[a]];
        llt.solve(rhs.data(), 1);
        for(int i = 0; i < m; i++) search[i] = 0.0;
        for(int a = 0; a < nf; a++) search[llt.idx[a]] = (-rhs[a]) - x[llt.idx[a]];
      }

      // descent check
      for(int i = 0; i < m; i++) tmp[i] = search[i] * grad[i];
      double sdotg = tree_sum32(tmp.data(), m);
      if(choice(kChoiceDescentTol) ? sdotg > 1e-10 : sdotg >= 0) // should not happen
      {
        retval = 0;
        break;
      }

      // Armijo line search
      double step = 1.0;
      for(int i = 0; i < m; i++) xc[i] = clampd(fmad(step, search[i], x[i]), lo[i], hi[i]);
      double objc = objective(H, m, g, xc.data(), Hxc.data());
      bool ls_fail = false;
      // boxQP.m: while (vc - oldvalue) / (step * sdotg) < Armijo.  step * sdotg < 0 here, so the ratio
      // test is evaluated cross-multiplied (no division in the backtracking loop); the two forms differ
      // only if the ratio is within rounding of the threshold
      auto armijo_fails = [&]() {
        if(kTextbook) return (objc - old_obj) / (step * sdotg) < cfg.armijo;
        return (objc - old_obj) > cfg.armijo * (step * sdotg);
      };
      while(armijo_fails())
      {
        step = step * cfg.step_factor;
        ls_steps++;
        for(int i = 0; i < m; i++) xc[i] = clampd(fmad(step, search[i], x[i]), lo[i], hi[i]);
        objc = objective(H, m, g, xc.data(), Hxc.data());
        if(step < cfg.min_step)
        {
          ls_fail = true;
          break;
        }
      }
      // accept candidate
      x = xc;
      Hx = Hxc;
      obj = objc;
      if(ls_fail)
      {
        retval = 2;
        break;
      }
      if(iter >= cfg.max_iter)
      {
        retval = 1;
        break;
      }
    }
    return retval;
  }
};
} // namespace oracle
