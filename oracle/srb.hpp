// oracle/srb.hpp — CPU restatement of CCC::DdpSingleRigidBody::DdpProblem.
//
// TEST INFRASTRUCTURE ONLY.  Follows reference src/DdpSingleRigidBody.cpp:
//   matAngularVelToEulerDot :26-38     stateEq :52-91     runningCost / terminalCost :93-113
//   calcStateEqDeriv :115-185 (SymPy-derived blocks transcribed verbatim)   cost derivs :187-245
// State x = (c, ZYX Euler angles, v, omega) — 12 states (BASELINE.json says "13-state"; the
// reference is 12, include/CCC/DdpSingleRigidBody.h:110).  Scalar formulas are written exactly as
// in the reference and compiled without contraction; sin/cos come from sincos_canon (num.hpp) so
// that the engine can reproduce them bit for bit; the 3x3 inertia solves use a plain
// LL^T (DenseLlt, boxqp.hpp), as the reference's inertia.llt().solve does.
#pragma once
#include "ddp.hpp"

namespace oracle
{
struct SrbProblem : public DdpProblem
{
  double dt = 0, mass = 0;
  int m_max = 0;
  const int32_t * m_tab = nullptr;  // [N]
  const double * ridge = nullptr;   // [N][m_max][3]
  const double * vertex = nullptr;  // [N][m_max][3]
  const double * inertia = nullptr; // [N][9] row-major
  const double * ref = nullptr;     // [N+1][6] pos, ori
  double w_run[13];
  double w_term[12];
  double u_lo = 0, u_hi = 0;

  SrbProblem() { nx = 12; }

  int inputDim(int k) const override { return m_tab[k]; }

  static void eulerMat(const double * ori, double * E)
  {
    double sa, ca, sb, cb;
    sincos_canon(ori[0], &sa, &ca);
    sincos_canon(ori[1], &sb, &cb);
    E[0] = (ca * sb) / cb;
    E[1] = (sb * sa) / cb;
    E[2] = 1.0;
    E[3] = -1 * sa;
    E[4] = ca;
    E[5] = 0.0;
    E[6] = ca / cb;
    E[7] = sa / cb;
    E[8] = 0.0;
  }

  DenseLlt inertiaLlt(int k) const
  {
    DenseLlt llt;
    llt.compute(inertia + 9 * k, 3, {0, 1, 2});
    return llt;
  }

  void wrench(int k, const double * x, const double * u, double * f, double * n) const
  {
    const int m = m_tab[k];
    double pf[3][32], pn[3][32];
    for(int j = 0; j < m; j++)
    {
      const double * rho = ridge + (static_cast<size_t>(k) * m_max + j) * 3;
      const double * vtx = vertex + (static_cast<size_t>(k) * m_max + j) * 3;
      double d[3] = {vtx[0] - x[0], vtx[1] - x[1], vtx[2] - x[2]};
      double cr[3];
      cross3(d, rho, cr);
      for(int a = 0; a < 3; a++)
      {
        pf[a][j] = u[j] * rho[a];
        pn[a][j] = u[j] * cr[a];
      }
    }
    for(int a = 0; a < 3; a++)
    {
      f[a] = tree_sum32(pf[a], m);
      n[a] = tree_sum32(pn[a], m);
    }
  }

  void stateEq(int k, const double * x, const double * u, double * xn) const override
  {
    const double * I = inertia + 9 * k;
    const double * w = x + 9;
    double f[3], n[3], E[9], xdot[12];
    wrench(k, x, u, f, n);
    eulerMat(x + 3, E);
    for(int a = 0; a < 3; a++)
    {
      xdot[a] = x[6 + a];
      xdot[3 + a] = dot_seq(E + 3 * a, 1, w, 1, 3);
      xdot[6 + a] = f[a] / mass;
    }
    xdot[8] = f[2] / mass + (-1 * kGravity);
    double Iw[3], cw[3], rhs[3];
    for(int a = 0; a < 3; a++) Iw[a] = dot_seq(I + 3 * a, 1, w, 1, 3);
    cross3(w, Iw, cw);
    for(int a = 0; a < 3; a++) rhs[a] = (-cw[a]) + n[a];
    inertiaLlt(k).solve(rhs, 1);
    for(int a = 0; a < 3; a++) xdot[9 + a] = rhs[a];
    for(int i = 0; i < 12; i++) xn[i] = fmad(dt, xdot[i], x[i]);
  }

  static double quad12(const double * w, const double * x, const double * r)
  {
    double c = 0.0;
    for(int a = 0; a < 6; a++)
    {
      double d = x[a] - r[a];
      c = fmad(w[a], d * d, c);
    }
    for(int a = 6; a < 12; a++) c = fmad(w[a], x[a] * x[a], c);
    return c;
  }

  double runningCost(int k, const double * x, const double * u) const override
  {
    const int m = m_tab[k];
    double usq[32];
    for(int j = 0; j < m; j++) usq[j] = u[j] * u[j];
    return fmad(0.5 * w_run[12], tree_sum32(usq, m), 0.5 * quad12(w_run, x, ref + 6 * k));
  }

  double terminalCost(const double * x) const override { return 0.5 * quad12(w_term, x, ref + 6 * N); }

  void stateEqDeriv(int k, const double * x, const double * u, double * Fx, double * Fu) const override
  {
    const int m = m_tab[k];
    const double * I = inertia + 9 * k;
    DenseLlt llt = inertiaLlt(k);
    double f[3], n[3], E[9];
    wrench(k, x, u, f, n);
    for(int i = 0; i < 144; i++) Fx[i] = 0.0;
    for(int a = 0; a < 3; a++) Fx[a * 12 + 6 + a] = 1.0;
    eulerMat(x + 3, E);
    for(int a = 0; a < 3; a++)
      for(int c = 0; c < 3; c++) Fx[(3 + a) * 12 + 9 + c] = E[3 * a + c];

    // SymPy-derived blocks, reference src/DdpSingleRigidBody.cpp:136-161 (verbatim expressions)
    double w1 = x[9], w2 = x[10], w3 = x[11];
    double sin_alpha, cos_alpha, sin_beta, cos_beta;
    sincos_canon(x[3], &sin_alpha, &cos_alpha);
    sincos_canon(x[4], &sin_beta, &cos_beta);
    double cos_beta_2 = cos_beta * cos_beta;
    double sin_beta_2 = sin_beta * sin_beta;
    double I11 = I[0], I12 = I[1], I13 = I[2], I22 = I[4], I23 = I[5], I33 = I[8];
    Fx[3 * 12 + 3] = -w1 * sin_alpha * sin_beta / cos_beta + w2 * sin_beta * cos_alpha / cos_beta;
    Fx[4 * 12 + 3] = -w1 * cos_alpha - w2 * sin_alpha;
    Fx[5 * 12 + 3] = -w1 * sin_alpha / cos_beta + w2 * cos_alpha / cos_beta;
    Fx[3 * 12 + 4] = w1 * sin_beta_2 * cos_alpha / cos_beta_2 + w1 * cos_alpha + w2 * sin_alpha * sin_beta_2 / cos_beta_2
                     + w2 * sin_alpha;
    Fx[4 * 12 + 4] = 0.0;
    Fx[5 * 12 + 4] = w1 * sin_beta * cos_alpha / cos_beta_2 + w2 * sin_alpha * sin_beta / cos_beta_2;
    double Mw[9];
    Mw[0] = I12 * w3 - I13 * w2;
    Mw[1] = -I13 * w1 + I22 * w3 - 2 * I23 * w2 - I33 * w3;
    Mw[2] = I12 * w1 + I22 * w2 + 2 * I23 * w3 - I33 * w2;
    Mw[3] = -I11 * w3 + 2 * I13 * w1 + I23 * w2 + I33 * w3;
    Mw[4] = -I12 * w3 + I23 * w1;
    Mw[5] = -I11 * w1 - I12 * w2 - 2 * I13 * w3 + I33 * w1;
    Mw[6] = I11 * w2 - 2 * I12 * w1 - I22 * w2 - I23 * w3;
    Mw[7] = I11 * w1 + 2 * I12 * w2 + I13 * w3 - I22 * w1;
    Mw[8] = I13 * w2 - I23 * w1;
    double col[3];
    for(int c = 0; c < 3; c++)
    {
      for(int a = 0; a < 3; a++) col[a] = Mw[3 * a + c];
      llt.solve(col, 1);
      for(int a = 0; a < 3; a++) Fx[(9 + a) * 12 + 9 + c] = col[a];
    }
    // I^-1 crossMat(totalForce)
    const double cm[9] = {0, -f[2], f[1], f[2], 0, -f[0], -f[1], f[0], 0};
    for(int c = 0; c < 3; c++)
    {
      for(int a = 0; a < 3; a++) col[a] = cm[3 * a + c];
      llt.solve(col, 1);
      for(int a = 0; a < 3; a++) Fx[(9 + a) * 12 + c] = col[a];
    }
    for(int i = 0; i < 144; i++) Fx[i] = Fx[i] * dt;
    for(int i = 0; i < 12; i++) Fx[i * 12 + i] = Fx[i * 12 + i] + 1.0;

    for(int i = 0; i < 12 * m; i++) Fu[i] = 0.0;
    for(int j = 0; j < m; j++)
    {
      const double * rho = ridge + (static_cast<size_t>(k) * m_max + j) * 3;
      const double * vtx = vertex + (static_cast<size_t>(k) * m_max + j) * 3;
      double d[3] = {vtx[0] - x[0], vtx[1] - x[1], vtx[2] - x[2]};
      double cr[3];
      cross3(d, rho, cr);
      llt.solve(cr, 1);
      for(int a = 0; a < 3; a++)
      {
        Fu[(6 + a) * m + j] = (rho[a] / mass) * dt;
        Fu[(9 + a) * m + j] = cr[a] * dt;
      }
    }
  }

  void runningCostDeriv(int k,
                        const double * x,
                        const double * u,
                        double * Lx,
                        double * Lu,
                        double * Lxx,
                        double * Luu,
                        double * Lxu) const override
  {
    const int m = m_tab[k];
    const double * r = ref + 6 * k;
    for(int a = 0; a < 6; a++) Lx[a] = w_run[a] * (x[a] - r[a]);
    for(int a = 6; a < 12; a++) Lx[a] = w_run[a] * x[a];
    for(int j = 0; j < m; j++) Lu[j] = w_run[12] * u[j];
    for(int i = 0; i < 144; i++) Lxx[i] = 0.0;
    for(int i = 0; i < 12; i++) Lxx[i * 12 + i] = w_run[i];
    for(int i = 0; i < m * m; i++) Luu[i] = 0.0;
    for(int j = 0; j < m; j++) Luu[j * m + j] = w_run[12];
    for(int i = 0; i < 12 * m; i++) Lxu[i] = 0.0;
  }

  void terminalCostDeriv(const double * x, double * Vx, double * Vxx) const override
  {
    const double * r = ref + 6 * N;
    for(int a = 0; a < 6; a++) Vx[a] = w_term[a] * (x[a] - r[a]);
    for(int a = 6; a < 12; a++) Vx[a] = w_term[a] * x[a];
    for(int i = 0; i < 144; i++) Vxx[i] = 0.0;
    for(int i = 0; i < 12; i++) Vxx[i * 12 + i] = w_term[i];
  }

  void inputLimits(int k, double * lo, double * hi) const override
  {
    for(int j = 0; j < m_tab[k]; j++)
    {
      lo[j] = u_lo;
      hi[j] = u_hi;
    }
  }
};
} // namespace oracle
