// oracle/centroidal.hpp — CPU restatement of CCC::DdpCentroidal::DdpProblem.
//
// TEST INFRASTRUCTURE ONLY.  Follows reference src/DdpCentroidal.cpp:
//   stateEq              :32-64      runningCost / terminalCost   :66-83
//   calcStateEqDeriv     :85-121     calcRunningCostDeriv         :123-154
//   calcTerminalCostDeriv:156-177    input limits lambda          :202-210
// The contact geometry (ForceColl::Contact::vertexWithRidgeList_) is consumed in its flat
// form: one (vertex, ridge) pair per input, in the loop order of :49-60.
#pragma once
#include "ddp.hpp"

namespace oracle
{
struct CentroidalProblem : public DdpProblem
{
  double dt = 0, mass = 0;
  int m_max = 0;
  const int32_t * m_tab = nullptr;  // [N]
  const double * ridge = nullptr;   // [N][m_max][3]
  const double * vertex = nullptr;  // [N][m_max][3]
  const double * ref_pos = nullptr; // [N+1][3]
  double w_run[10];
  double w_term[9];
  double u_lo = 0, u_hi = 0;

  CentroidalProblem() { nx = 9; }

  int inputDim(int k) const override { return m_tab[k]; }

  /** total force  sum_j u_j rho_j  and moment  sum_j u_j (p_j - c) x rho_j  about the CoM c. */
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
    double f[3], n[3], xdot[9];
    wrench(k, x, u, f, n);
    // P / m as P * (1 / m): the same rounded 1 / m that calcStateEqDeriv puts into Fx (:100), so that
    // the rollout and its linearisation agree, and no division sits on the per-stage critical path
    const double inv_mass = 1 / mass;
    for(int a = 0; a < 3; a++)
    {
      xdot[a] = kTextbook ? x[3 + a] / mass : x[3 + a] * inv_mass; // reference :44 divides
      xdot[3 + a] = f[a];
      xdot[6 + a] = n[a];
    }
    xdot[5] = f[2] + (-1 * mass * kGravity);
    for(int i = 0; i < 9; i++) xn[i] = fmad(dt, xdot[i], x[i]);
  }

  static double quad9(const double * w, const double * x, const double * ref)
  {
    double c = 0.0;
    for(int a = 0; a < 3; a++)
    {
      double d = x[a] - ref[a];
      c = fmad(w[a], d * d, c);
    }
    for(int a = 3; a < 9; a++) c = fmad(w[a], x[a] * x[a], c);
    return c;
  }

  double runningCost(int k, const double * x, const double * u) const override
  {
    const int m = m_tab[k];
    double usq[32];
    for(int j = 0; j < m; j++) usq[j] = u[j] * u[j];
    double c = quad9(w_run, x, ref_pos + 3 * k);
    return fmad(0.5 * w_run[9], tree_sum32(usq, m), 0.5 * c);
  }

  double terminalCost(const double * x) const override { return 0.5 * quad9(w_term, x, ref_pos + 3 * N); }

  void stateEqDeriv(int k, const double * x, const double * u, double * Fx, double * Fu) const override
  {
    const int m = m_tab[k];
    double f[3], n[3];
    wrench(k, x, u, f, n);
    for(int i = 0; i < 81; i++) Fx[i] = 0.0;
    for(int i = 0; i < 9; i++) Fx[i * 9 + i] = 1.0;
    const double inv_m_dt = (1 / mass) * dt;
    for(int a = 0; a < 3; a++) Fx[a * 9 + 3 + a] = inv_m_dt;
    // crossMat(totalForce) * dt in rows 6..8, cols 0..2
    Fx[6 * 9 + 1] = -f[2] * dt;
    Fx[6 * 9 + 2] = f[1] * dt;
    Fx[7 * 9 + 0] = f[2] * dt;
    Fx[7 * 9 + 2] = -f[0] * dt;
    Fx[8 * 9 + 0] = -f[1] * dt;
    Fx[8 * 9 + 1] = f[0] * dt;
    for(int i = 0; i < 9 * m; i++) Fu[i] = 0.0;
    for(int j = 0; j < m; j++)
    {
      const double * rho = ridge + (static_cast<size_t>(k) * m_max + j) * 3;
      const double * vtx = vertex + (static_cast<size_t>(k) * m_max + j) * 3;
      double d[3] = {vtx[0] - x[0], vtx[1] - x[1], vtx[2] - x[2]};
      double cr[3];
      cross3(d, rho, cr);
      for(int a = 0; a < 3; a++)
      {
        Fu[(3 + a) * m + j] = rho[a] * dt;
        Fu[(6 + a) * m + j] = cr[a] * dt;
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
    const double * ref = ref_pos + 3 * k;
    for(int a = 0; a < 3; a++) Lx[a] = w_run[a] * (x[a] - ref[a]);
    for(int a = 3; a < 9; a++) Lx[a] = w_run[a] * x[a];
    for(int j = 0; j < m; j++) Lu[j] = w_run[9] * u[j];
    for(int i = 0; i < 81; i++) Lxx[i] = 0.0;
    for(int i = 0; i < 9; i++) Lxx[i * 9 + i] = w_run[i];
    for(int i = 0; i < m * m; i++) Luu[i] = 0.0;
    for(int j = 0; j < m; j++) Luu[j * m + j] = w_run[9];
    for(int i = 0; i < 9 * m; i++) Lxu[i] = 0.0;
  }

  void terminalCostDeriv(const double * x, double * Vx, double * Vxx) const override
  {
    const double * ref = ref_pos + 3 * N;
    for(int a = 0; a < 3; a++) Vx[a] = w_term[a] * (x[a] - ref[a]);
    for(int a = 3; a < 9; a++) Vx[a] = w_term[a] * x[a];
    for(int i = 0; i < 81; i++) Vxx[i] = 0.0;
    for(int i = 0; i < 9; i++) Vxx[i * 9 + i] = w_term[i];
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
