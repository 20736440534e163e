// ddp_thread_zmp.cuh — CCC::DdpZmp, one THREAD per problem.
//
// DdpZmp has 6 states and 3 unconstrained inputs: under the warp-per-problem core (ddp_warp_core.cuh, lane = input)
// 3 of 32 lanes carry an input.  This header is the whole solver — nmpc_ddp::DDPSolver<6,3>::solve (unconstrained
// path: plain factorisation of Quu + lambda I) around CCC::DdpZmp::DdpProblem (reference src/DdpZmp.cpp:8-143) — as
// straight-line scalar code for one problem, so that a warp solves 32 problems.  Evaluation order: oracle/ddp.hpp
// (solve, procOnce, backwardPass, forwardPass), oracle/boxqp.hpp (FreeLlt) and oracle/zmp.hpp, operation by
// operation: sequential fma chains from +0.0, dot4 / tree_sum32 written out for three terms.
//
// The code is plain C++ (fma / sqrt only) and compiles for the host too: tests/emu/ runs it on the CPU against the
// oracle, bit for bit.  Per-problem arrays (trajectories, costs, gains) live in global memory with the problem
// index fastest (element e of problem b at base[e * stride + b]): the 32 threads of a warp read 32 consecutive
// doubles.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

#include "../../include/ccc_b200.h"

#ifdef __CUDACC__
#define CCC_TD __host__ __device__ __forceinline__
#else
#define CCC_TD inline
#endif

namespace ccc_thread
{
constexpr double kG = 9.80665; // reference include/CCC/Constants.h:10

struct ZmpModelParams
{
  int N;
  double dt, mass;
  double w_run_com_z, w_run_zmp, w_run_fz, w_term_xy, w_term_z, w_term_vel;
  const double * ref_zmp; // [N+1][3] of this problem's schedule
  const double * com_z;   // [N+1]
};

/** Per-problem strided storage. */
struct ZmpWork
{
  double * x[2];    // [N+1][6]
  double * u[2];    // [N][3]
  double * cost[2]; // [N+1]
  double * kl;      // [N][3]
  double * Kl;      // [N][3][6]
  size_t stride;
};

CCC_TD double tree3(double t0, double t1, double t2)
{
  // tree_sum32 over (t0, t1, t2, 0, ...): the zero leaves only turn -0 into +0
  return ((t0 + 0.0) + (t2 + 0.0)) + (t1 + 0.0);
}

struct ZmpThreadSolver
{
  ZmpModelParams P;
  ZmpWork W;
  const ccc_ddp_config_t * cfg;
  double lambda, dlambda, dV0, dV1;
  int cur; // index of the nominal trajectory buffers

  CCC_TD double & X(int buf, int k, int i) const { return W.x[buf][(size_t)(k * 6 + i) * W.stride]; }
  CCC_TD double & U(int buf, int k, int j) const { return W.u[buf][(size_t)(k * 3 + j) * W.stride]; }
  CCC_TD double & C(int buf, int k) const { return W.cost[buf][(size_t)k * W.stride]; }
  CCC_TD double & KL(int k, int j) const { return W.kl[(size_t)(k * 3 + j) * W.stride]; }
  CCC_TD double & KK(int k, int j, int c) const { return W.Kl[(size_t)((k * 3 + j) * 6 + c) * W.stride]; }

  // ---- CCC::DdpZmp::DdpProblem (oracle/zmp.hpp) ----
  CCC_TD void stateEq(int k, const double * x, const double * u, double * xn) const
  {
    const double zz = P.ref_zmp[3 * k + 2];
    double xdot[6];
    xdot[0] = x[1];
    xdot[1] = (x[0] - u[0]) * u[2] / (P.mass * (x[4] - zz));
    xdot[2] = x[3];
    xdot[3] = (x[2] - u[1]) * u[2] / (P.mass * (x[4] - zz));
    xdot[4] = x[5];
    xdot[5] = u[2] / P.mass - kG;
    for(int i = 0; i < 6; i++) xn[i] = fma(P.dt, xdot[i], x[i]);
  }

  CCC_TD double runningCost(int k, const double * x, const double * u) const
  {
    const double r[3] = {P.ref_zmp[3 * k], P.ref_zmp[3 * k + 1], P.mass * kG};
    const double w[3] = {P.w_run_zmp, P.w_run_zmp, P.w_run_fz};
    double t[3];
    for(int j = 0; j < 3; j++)
    {
      const double d = u[j] - r[j];
      t[j] = w[j] * (d * d);
    }
    const double wx[6] = {0, 0, 0, 0, P.w_run_com_z, 0}, rx[6] = {0, 0, 0, 0, P.com_z[k], 0};
    double q = 0.0;
    for(int a = 0; a < 6; a++)
    {
      const double d = x[a] - rx[a];
      q = fma(wx[a], d * d, q);
    }
    return fma(0.5 * 1.0, tree3(t[0], t[1], t[2]), 0.5 * q);
  }

  CCC_TD void termWeights(double * w, double * r) const
  {
    const int N = P.N;
    const double ww[6] = {P.w_term_xy, P.w_term_vel, P.w_term_xy, P.w_term_vel, P.w_term_z, P.w_term_vel};
    const double rr[6] = {P.ref_zmp[3 * N], 0, P.ref_zmp[3 * N + 1], 0, P.com_z[N], 0};
    for(int a = 0; a < 6; a++)
    {
      w[a] = ww[a];
      r[a] = rr[a];
    }
  }

  CCC_TD double terminalCost(const double * x) const
  {
    double w[6], r[6], q = 0.0;
    termWeights(w, r);
    for(int a = 0; a < 6; a++)
    {
      const double d = x[a] - r[a];
      q = fma(w[a], d * d, q);
    }
    return 0.5 * q;
  }

  CCC_TD void stateEqDeriv(int k, const double * x, const double * u, double * Fx, double * Fu) const
  {
    const double zz = P.ref_zmp[3 * k + 2], mass = P.mass, dt = P.dt;
    for(int i = 0; i < 36; i++) Fx[i] = 0.0;
    Fx[0 * 6 + 1] = 1;
    Fx[1 * 6 + 0] = u[2] / (mass * (x[4] - zz));
    Fx[1 * 6 + 4] = -1 * (x[0] - u[0]) * u[2] / (mass * ((x[4] - zz) * (x[4] - zz)));
    Fx[2 * 6 + 3] = 1;
    Fx[3 * 6 + 2] = u[2] / (mass * (x[4] - zz));
    Fx[3 * 6 + 4] = -1 * (x[2] - u[1]) * u[2] / (mass * ((x[4] - zz) * (x[4] - zz)));
    Fx[4 * 6 + 5] = 1;
    for(int i = 0; i < 36; i++) Fx[i] = Fx[i] * dt;
    for(int i = 0; i < 6; i++) Fx[i * 6 + i] = Fx[i * 6 + i] + 1.0;
    for(int i = 0; i < 18; i++) Fu[i] = 0.0;
    Fu[1 * 3 + 0] = -1 * u[2] / (mass * (x[4] - zz));
    Fu[1 * 3 + 2] = (x[0] - u[0]) / (mass * (x[4] - zz));
    Fu[3 * 3 + 1] = -1 * u[2] / (mass * (x[4] - zz));
    Fu[3 * 3 + 2] = (x[2] - u[1]) / (mass * (x[4] - zz));
    Fu[5 * 3 + 2] = 1 / mass;
    for(int i = 0; i < 18; i++) Fu[i] = Fu[i] * dt;
  }

  // ---- nmpc_ddp::DDPSolver (oracle/ddp.hpp) ----
  CCC_TD void increaseLambda()
  {
    const double f = cfg->lambda_factor;
    const double a = dlambda * f;
    dlambda = a < f ? f : a; // std::max(a, f)
    const double b = lambda * dlambda;
    lambda = b < cfg->lambda_min ? cfg->lambda_min : b;
  }
  CCC_TD void decreaseLambda()
  {
    const double f = cfg->lambda_factor;
    const double a = dlambda / f, c = 1.0 / f;
    dlambda = c < a ? c : a; // std::min(a, c)
    lambda = (lambda * dlambda) * (lambda > cfg->lambda_min ? 1.0 : 0.0);
  }

  CCC_TD double sumCost(int buf) const
  {
    double s = 0.0;
    for(int k = 0; k <= P.N; k++) s = s + C(buf, k);
    return s;
  }

  /** Unconstrained backward pass; false = Quu + lambda I not positive definite at some stage. */
  CCC_TD bool backwardPass()
  {
    const int N = P.N;
    double Vx[6], Vxx[36];
    {
      double xN[6], w[6], r[6];
      for(int i = 0; i < 6; i++) xN[i] = X(cur, N, i);
      termWeights(w, r);
      for(int a = 0; a < 6; a++) Vx[a] = w[a] * (xN[a] - r[a]);
      for(int i = 0; i < 36; i++) Vxx[i] = 0.0;
      for(int a = 0; a < 6; a++) Vxx[a * 6 + a] = w[a];
    }
    dV0 = dV1 = 0.0;
    for(int k = N - 1; k >= 0; k--)
    {
      double x[6], u[3], Fx[36], Fu[18];
      for(int i = 0; i < 6; i++) x[i] = X(cur, k, i);
      for(int j = 0; j < 3; j++) u[j] = U(cur, k, j);
      stateEqDeriv(k, x, u, Fx, Fu);
      // runningCostDeriv
      double Lx[6], Lu[3];
      const double wx[6] = {0, 0, 0, 0, P.w_run_com_z, 0}, rx[6] = {0, 0, 0, 0, P.com_z[k], 0};
      const double wu[3] = {P.w_run_zmp, P.w_run_zmp, P.w_run_fz};
      const double ur[3] = {P.ref_zmp[3 * k], P.ref_zmp[3 * k + 1], P.mass * kG};
      for(int a = 0; a < 6; a++) Lx[a] = wx[a] * (x[a] - rx[a]);
      for(int j = 0; j < 3; j++) Lu[j] = wu[j] * (u[j] - ur[j]);
      // Q-function
      double Qx[6], Qu[3], T[36], Qxx[36], Wm[18], Quu[9], Qux[18];
      for(int i = 0; i < 6; i++)
      {
        double acc = 0.0;
        for(int r = 0; r < 6; r++) acc = fma(Fx[r * 6 + i], Vx[r], acc);
        Qx[i] = Lx[i] + acc;
      }
      for(int j = 0; j < 3; j++)
      {
        double acc = 0.0;
        for(int r = 0; r < 6; r++) acc = fma(Fu[r * 3 + j], Vx[r], acc);
        Qu[j] = Lu[j] + acc;
      }
      for(int i = 0; i < 6; i++)
        for(int j = 0; j < 6; j++)
        {
          double acc = 0.0;
          for(int r = 0; r < 6; r++) acc = fma(Vxx[i * 6 + r], Fx[r * 6 + j], acc);
          T[i * 6 + j] = acc;
        }
      for(int i = 0; i < 6; i++)
        for(int j = 0; j < 6; j++)
        {
          double acc = 0.0;
          for(int r = 0; r < 6; r++) acc = fma(Fx[r * 6 + i], T[r * 6 + j], acc);
          Qxx[i * 6 + j] = (i == j ? wx[i] : 0.0) + acc;
        }
      for(int i = 0; i < 6; i++)
        for(int j = 0; j < 3; j++)
        {
          double acc = 0.0;
          for(int r = 0; r < 6; r++) acc = fma(Vxx[i * 6 + r], Fu[r * 3 + j], acc);
          Wm[i * 3 + j] = acc;
        }
      for(int i = 0; i < 3; i++)
        for(int j = 0; j <= i; j++)
        {
          double acc = 0.0;
          for(int r = 0; r < 6; r++) acc = fma(Fu[r * 3 + i], Wm[r * 3 + j], acc);
          Quu[i * 3 + j] = (i == j ? wu[i] : 0.0) + acc;
          Quu[j * 3 + i] = Quu[i * 3 + j];
        }
      for(int j = 0; j < 3; j++)
        for(int i = 0; i < 6; i++)
        {
          double acc = 0.0;
          for(int r = 0; r < 6; r++) acc = fma(Fu[r * 3 + j], T[r * 6 + i], acc);
          Qux[j * 6 + i] = 0.0 + acc;
        }
      // L D L' of Quu + lambda I (FreeLlt::compute, all indices free)
      double QuuF[9], L[9], Cc[9], invd[3];
      for(int i = 0; i < 9; i++) QuuF[i] = Quu[i];
      for(int j = 0; j < 3; j++) QuuF[j * 3 + j] = Quu[j * 3 + j] + lambda;
      for(int i = 0; i < 9; i++) L[i] = Cc[i] = 0.0;
      for(int c = 0; c < 3; c++)
      {
        double d = QuuF[c * 3 + c];
        for(int j = 0; j < c; j++) d = fma(-L[c * 3 + j], Cc[c * 3 + j], d);
        if(!(d > 0.0)) return false;
        const double inv = 1.0 / d;
        Cc[c * 3 + c] = d;
        invd[c] = inv;
        for(int i = c + 1; i < 3; i++)
        {
          double a = QuuF[i * 3 + c];
          for(int j = 0; j < c; j++) a = fma(-L[i * 3 + j], Cc[c * 3 + j], a);
          Cc[i * 3 + c] = a;
          L[i * 3 + c] = a * inv;
        }
      }
      auto llt_solve = [&](double * b) {
        for(int i = 0; i < 3; i++)
        {
          double acc = b[i];
          for(int j = 0; j < i; j++) acc = fma(-L[i * 3 + j], b[j], acc);
          b[i] = acc;
        }
        for(int i = 0; i < 3; i++) b[i] = b[i] * invd[i];
        for(int i = 2; i >= 0; i--)
        {
          double acc = b[i];
          for(int j = 2; j > i; j--) acc = fma(-L[j * 3 + i], b[j], acc);
          b[i] = acc;
        }
      };
      double kk[3], Kg[18];
      {
        double rhs[3] = {Qu[0], Qu[1], Qu[2]};
        llt_solve(rhs);
        for(int j = 0; j < 3; j++) kk[j] = -rhs[j];
      }
      for(int c = 0; c < 6; c++)
      {
        double rhs[3] = {Qux[0 * 6 + c], Qux[1 * 6 + c], Qux[2 * 6 + c]};
        llt_solve(rhs);
        for(int j = 0; j < 3; j++) Kg[j * 6 + c] = -rhs[j];
      }
      for(int j = 0; j < 3; j++)
      {
        KL(k, j) = kk[j];
        for(int c = 0; c < 6; c++) KK(k, j, c) = Kg[j * 6 + c];
      }
      // cost-to-go update (unregularised Quu)
      double Quuk[3], QuuK[18], t1[3], t2[3];
      for(int i = 0; i < 3; i++)
      {
        // dot4 over three terms
        const double s0 = fma(Quu[i * 3 + 0], kk[0], 0.0), s1 = fma(Quu[i * 3 + 1], kk[1], 0.0), s2 = fma(Quu[i * 3 + 2], kk[2], 0.0);
        Quuk[i] = (s0 + s1) + (s2 + 0.0);
        for(int c = 0; c < 6; c++)
        {
          double acc = 0.0;
          for(int j = 0; j < 3; j++) acc = fma(Quu[i * 3 + j], Kg[j * 6 + c], acc);
          QuuK[i * 6 + c] = acc;
        }
        t1[i] = kk[i] * Qu[i];
        t2[i] = kk[i] * Quuk[i];
      }
      dV0 = dV0 + tree3(t1[0], t1[1], t1[2]);
      dV1 = fma(0.5, tree3(t2[0], t2[1], t2[2]), dV1);
      for(int c = 0; c < 6; c++)
      {
        double a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for(int j = 0; j < 3; j++) a1 = fma(Kg[j * 6 + c], Quuk[j], a1);
        for(int j = 0; j < 3; j++) a2 = fma(Kg[j * 6 + c], Qu[j], a2);
        for(int j = 0; j < 3; j++) a3 = fma(Qux[j * 6 + c], kk[j], a3);
        Vx[c] = ((Qx[c] + a1) + a2) + a3;
      }
      double Vn[36];
      for(int a = 0; a < 6; a++)
        for(int b = 0; b < 6; b++)
        {
          double s1 = 0.0, s2 = 0.0, s3 = 0.0;
          for(int j = 0; j < 3; j++) s1 = fma(Kg[j * 6 + a], QuuK[j * 6 + b], s1);
          for(int j = 0; j < 3; j++) s2 = fma(Kg[j * 6 + a], Qux[j * 6 + b], s2);
          for(int j = 0; j < 3; j++) s3 = fma(Qux[j * 6 + a], Kg[j * 6 + b], s3);
          Vn[a * 6 + b] = ((Qxx[a * 6 + b] + s1) + s2) + s3;
        }
      for(int a = 0; a < 6; a++)
        for(int b = 0; b < 6; b++) Vxx[a * 6 + b] = 0.5 * (Vn[a * 6 + b] + Vn[b * 6 + a]);
    }
    return true;
  }

  CCC_TD void forwardPass(double alpha)
  {
    const int N = P.N, cand = cur ^ 1;
    double xc[6];
    for(int i = 0; i < 6; i++)
    {
      xc[i] = X(cur, 0, i);
      X(cand, 0, i) = xc[i];
    }
    // the gains and the nominal (x, u) of a stage do not depend on the stages before it: they are loaded one stage
    // ahead, before this stage's stores (which the compiler must assume to alias them), so that their latency hides
    // behind the stage's arithmetic
    double Kq[18], kq[3], uq[3], xq[6];
    for(int j = 0; j < 3; j++)
    {
      for(int c = 0; c < 6; c++) Kq[j * 6 + c] = KK(0, j, c);
      kq[j] = KL(0, j);
      uq[j] = U(cur, 0, j);
    }
    for(int c = 0; c < 6; c++) xq[c] = xc[c];
    for(int k = 0; k < N; k++)
    {
      double Kg[18], kk[3], un[3], dx[6], uc[3], xn[6];
      for(int e = 0; e < 18; e++) Kg[e] = Kq[e];
      for(int j = 0; j < 3; j++)
      {
        kk[j] = kq[j];
        un[j] = uq[j];
      }
      for(int c = 0; c < 6; c++) dx[c] = xc[c] - xq[c];
      if(k + 1 < N)
      {
        for(int j = 0; j < 3; j++)
        {
          for(int c = 0; c < 6; c++) Kq[j * 6 + c] = KK(k + 1, j, c);
          kq[j] = KL(k + 1, j);
          uq[j] = U(cur, k + 1, j);
        }
        for(int c = 0; c < 6; c++) xq[c] = X(cur, k + 1, c);
      }
      for(int j = 0; j < 3; j++)
      {
        double fb = 0.0;
        for(int c = 0; c < 6; c++) fb = fma(Kg[j * 6 + c], dx[c], fb);
        uc[j] = fma(alpha, kk[j], un[j]) + fb;
      }
      stateEq(k, xc, uc, xn);
      const double ck = runningCost(k, xc, uc);
      for(int j = 0; j < 3; j++) U(cand, k, j) = uc[j];
      C(cand, k) = ck;
      for(int i = 0; i < 6; i++)
      {
        xc[i] = xn[i];
        X(cand, k + 1, i) = xn[i];
      }
    }
    C(cand, N) = terminalCost(xc);
  }

  /** solve(): returns the last procOnce value (1 converged, 0 max_iter, -1 lambda_max); iterations in *iters_out.
   *  Trace slots (accepted alpha index / lambda per iteration) are written with stride trace_stride when non-null. */
  CCC_TD int solve(const double * x0, const double * u_init, size_t u_init_stride, int * iters_out, int8_t * alpha_idx, double * lambda_trace,
                   int trace_len, size_t trace_stride)
  {
    const int N = P.N;
    lambda = cfg->initial_lambda;
    dlambda = cfg->initial_dlambda;
    cur = 0;
    {
      double x[6], u[3], xn[6];
      for(int i = 0; i < 6; i++)
      {
        x[i] = x0[i];
        X(0, 0, i) = x[i];
      }
      for(int k = 0; k < N; k++)
      {
        for(int j = 0; j < 3; j++)
        {
          u[j] = u_init ? u_init[(size_t)(k * 3 + j) * u_init_stride] : 0.0;
          U(0, k, j) = u[j];
        }
        stateEq(k, x, u, xn);
        C(0, k) = runningCost(k, x, u);
        for(int i = 0; i < 6; i++)
        {
          x[i] = xn[i];
          X(0, k + 1, i) = xn[i];
        }
      }
      C(0, N) = terminalCost(x);
    }
    for(int i = 0; i < trace_len; i++)
    {
      if(alpha_idx) alpha_idx[(size_t)i * trace_stride] = -4;
      if(lambda_trace) lambda_trace[(size_t)i * trace_stride] = 0.0;
    }
    int retval = 0, iter = 0;
    for(iter = 1; iter <= cfg->max_iter; iter++)
    {
      int aidx = -4;
      retval = procOnce(aidx);
      if(iter - 1 < trace_len)
      {
        if(alpha_idx) alpha_idx[(size_t)(iter - 1) * trace_stride] = (int8_t)aidx;
        if(lambda_trace) lambda_trace[(size_t)(iter - 1) * trace_stride] = lambda;
      }
      if(retval != 0) break;
    }
    *iters_out = iter <= cfg->max_iter ? iter : cfg->max_iter;
    return retval;
  }

  CCC_TD int procOnce(int & aidx)
  {
    const int N = P.N;
    while(!backwardPass())
    {
      increaseLambda();
      if(lambda > cfg->lambda_max)
      {
        aidx = -3;
        return -1;
      }
    }
    double k_rel_norm = 0;
    for(int k = 0; k < N; k++)
    {
      const double k0 = KL(k, 0), k1 = KL(k, 1), k2 = KL(k, 2);
      const double u0 = U(cur, k, 0), u1 = U(cur, k, 1), u2 = U(cur, k, 2);
      const double kn = sqrt(tree3(k0 * k0, k1 * k1, k2 * k2));
      const double un = sqrt(tree3(u0 * u0, u1 * u1, u2 * u2));
      const double v = kn / (un + 1.0);
      k_rel_norm = k_rel_norm < v ? v : k_rel_norm; // std::max(a, b): a unless a < b
    }
    if(k_rel_norm < cfg->k_rel_norm_thre && lambda < cfg->lambda_thre)
    {
      decreaseLambda();
      aidx = -2;
      return 1;
    }
    bool success = false;
    double actual = 0;
    const double J = sumCost(cur);
    for(int a = 0; a < cfg->n_alpha; a++)
    {
      const double alpha = cfg->alpha[a];
      forwardPass(alpha);
      actual = J - sumCost(cur ^ 1);
      const double expected = -(alpha * fma(alpha, dV1, dV0));
      double ratio;
      if(expected > 0)
        ratio = actual / expected;
      else
        ratio = (double)((0 < actual) - (actual < 0));
      if(ratio > cfg->cost_update_ratio_thre)
      {
        success = true;
        aidx = a;
        break;
      }
    }
    int rv = 0;
    if(success)
    {
      decreaseLambda();
      cur ^= 1;
      if(actual < cfg->cost_update_thre) rv = 1;
    }
    else
    {
      aidx = -1;
      increaseLambda();
      if(lambda > cfg->lambda_max) rv = -1;
    }
    return rv;
  }
};

/** Solve problem b of `bt` on the per-problem storage `W` (already offset to problem b) and store what the C-ABI
 *  returns for it (ccc_ddp_result_t layout; every pointer may be null). */
CCC_TD void zmp_thread_run(const ccc_ddp_zmp_batch_t & bt, const ccc_ddp_config_t & cfg, const ccc_ddp_result_t & res, const ZmpWork & W, int b)
{
  const int N = bt.horizon_steps;
  const int s = bt.sched_id[b];
  ZmpThreadSolver sv;
  sv.P.N = N;
  sv.P.dt = bt.dt;
  sv.P.mass = bt.mass;
  sv.P.w_run_com_z = bt.w[0];
  sv.P.w_run_zmp = bt.w[1];
  sv.P.w_run_fz = bt.w[2];
  sv.P.w_term_xy = bt.w[3];
  sv.P.w_term_z = bt.w[4];
  sv.P.w_term_vel = bt.w[5];
  sv.P.ref_zmp = bt.ref_zmp + (size_t)s * (N + 1) * 3;
  sv.P.com_z = bt.com_z + (size_t)s * (N + 1);
  sv.W = W;
  sv.cfg = &cfg;
  int iters = 0;
  const int tl = res.trace_len;
  const int rv = sv.solve(bt.x0 + (size_t)b * 6, bt.u_init ? bt.u_init + (size_t)b * N * 3 : nullptr, 1, &iters,
                          res.alpha_idx ? res.alpha_idx + (size_t)b * tl : nullptr, res.lambda_trace ? res.lambda_trace + (size_t)b * tl : nullptr,
                          tl, 1);
  // copies out of the interleaved storage: a block of loads, then its stores (the loads of one block are in flight
  // together instead of one round trip per element)
  if(res.x)
    for(int k = 0; k <= N; k += 2)
    {
      double v[12];
      const int nk = k + 1 <= N ? 2 : 1;
      for(int e = 0; e < 6 * nk; e++) v[e] = sv.X(sv.cur, k + e / 6, e % 6);
      for(int e = 0; e < 6 * nk; e++) res.x[((size_t)b * (N + 1) + k) * 6 + e] = v[e];
    }
  if(res.u)
    for(int k = 0; k < N; k += 4)
    {
      double v[12];
      const int nk = N - k < 4 ? N - k : 4;
      for(int e = 0; e < 3 * nk; e++) v[e] = sv.U(sv.cur, k + e / 3, e % 3);
      for(int e = 0; e < 3 * nk; e++) res.u[((size_t)b * N + k) * 3 + e] = v[e];
    }
  if(res.cost) res.cost[b] = sv.sumCost(sv.cur);
  if(res.iters) res.iters[b] = iters;
  if(res.status) res.status[b] = rv;
  if(res.clamped)
    for(int k = 0; k < N; k++) res.clamped[(size_t)b * N + k] = 0u;
}

/** Doubles of per-problem storage (ZmpWork) per problem. */
CCC_TD size_t zmp_work_doubles(int N)
{
  return (size_t)2 * (N + 1) * 6 + (size_t)2 * N * 3 + (size_t)2 * (N + 1) + (size_t)N * 3 + (size_t)N * 18;
}

/** Carve the arrays of problem b out of a slab of zmp_work_doubles(N) * stride doubles. */
CCC_TD ZmpWork zmp_work_at(double * slab, int N, size_t stride, size_t b)
{
  ZmpWork W;
  double * p = slab + b;
  W.x[0] = p;
  p += (size_t)(N + 1) * 6 * stride;
  W.x[1] = p;
  p += (size_t)(N + 1) * 6 * stride;
  W.u[0] = p;
  p += (size_t)N * 3 * stride;
  W.u[1] = p;
  p += (size_t)N * 3 * stride;
  W.cost[0] = p;
  p += (size_t)(N + 1) * stride;
  W.cost[1] = p;
  p += (size_t)(N + 1) * stride;
  W.kl = p;
  p += (size_t)N * 3 * stride;
  W.Kl = p;
  W.stride = stride;
  return W;
}
} // namespace ccc_thread
