// oracle/zmp.hpp — CPU restatement of CCC::DdpZmp::DdpProblem.
//
// TEST INFRASTRUCTURE ONLY.  Follows reference src/DdpZmp.cpp: stateEq :8-18, runningCost :20-27,
// terminalCost :29-41, calcStateEqDeriv :43-72, calcRunningCostDeriv :74-110, calcTerminalCostDeriv :112-143.
// 6 states (c_x, v_x, c_y, v_y, c_z, v_z), 3 inputs (zmp_x, zmp_y, f_z), no input limits (DdpZmp's
// constructor only sets horizon_steps, include/CCC/DdpZmp.h:277-282).  Scalar formulas verbatim, compiled
// without contraction; the three input-cost terms go through tree_sum32 like every reduction over inputs.
#pragma once
#include "ddp.hpp"

namespace oracle
{
struct ZmpProblem : public DdpProblem
{
  double dt = 0, mass = 0;
  const double * ref_zmp = nullptr; // [N+1][3]
  const double * com_z = nullptr;   // [N+1]
  double w_run_com_z = 0, w_run_zmp = 0, w_run_fz = 0, w_term_xy = 0, w_term_z = 0, w_term_vel = 0;

  ZmpProblem() { nx = 6; }
  int inputDim(int) const override { return 3; }

  void stateEq(int k, const double * x, const double * u, double * xn) const override
  {
    const double zz = ref_zmp[3 * k + 2];
    double xdot[6];
    xdot[0] = x[1];
    xdot[1] = (x[0] - u[0]) * u[2] / (mass * (x[4] - zz));
    xdot[2] = x[3];
    xdot[3] = (x[2] - u[1]) * u[2] / (mass * (x[4] - zz));
    xdot[4] = x[5];
    xdot[5] = u[2] / mass - kGravity;
    for(int i = 0; i < 6; i++) xn[i] = fmad(dt, xdot[i], x[i]);
  }

  void uref(int k, double * r) const
  {
    r[0] = ref_zmp[3 * k];
    r[1] = ref_zmp[3 * k + 1];
    r[2] = mass * kGravity;
  }

  double runningCost(int k, const double * x, const double * u) const override
  {
    double r[3], t[3];
    uref(k, r);
    const double w[3] = {w_run_zmp, w_run_zmp, w_run_fz};
    for(int j = 0; j < 3; j++)
    {
      const double d = u[j] - r[j];
      t[j] = w[j] * (d * d);
    }
    // state part: only c_z is weighted; written as the generic diagonal quadratic with zero weights
    const double wx[6] = {0, 0, 0, 0, w_run_com_z, 0}, rx[6] = {0, 0, 0, 0, com_z[k], 0};
    double q = 0.0;
    for(int a = 0; a < 6; a++)
    {
      const double d = x[a] - rx[a];
      q = fmad(wx[a], d * d, q);
    }
    return fmad(0.5 * 1.0, tree_sum32(t, 3), 0.5 * q);
  }

  void termWeights(double * w, double * r) const
  {
    const double ww[6] = {w_term_xy, w_term_vel, w_term_xy, w_term_vel, w_term_z, w_term_vel};
    const double rr[6] = {ref_zmp[3 * N], 0, ref_zmp[3 * N + 1], 0, com_z[N], 0};
    for(int a = 0; a < 6; a++)
    {
      w[a] = ww[a];
      r[a] = rr[a];
    }
  }

  double terminalCost(const double * x) const override
  {
    double w[6], r[6], q = 0.0;
    termWeights(w, r);
    for(int a = 0; a < 6; a++)
    {
      const double d = x[a] - r[a];
      q = fmad(w[a], d * d, q);
    }
    return 0.5 * q;
  }

  void stateEqDeriv(int k, const double * x, const double * u, double * Fx, double * Fu) const override
  {
    const double zz = ref_zmp[3 * k + 2];
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

  void runningCostDeriv(int k, const double * x, const double * u, double * Lx, double * Lu, double * Lxx, double * Luu,
                        double * Lxu) const override
  {
    double r[3];
    uref(k, r);
    const double w[3] = {w_run_zmp, w_run_zmp, w_run_fz};
    const double wx[6] = {0, 0, 0, 0, w_run_com_z, 0}, rx[6] = {0, 0, 0, 0, com_z[k], 0};
    for(int a = 0; a < 6; a++) Lx[a] = wx[a] * (x[a] - rx[a]);
    for(int j = 0; j < 3; j++) Lu[j] = w[j] * (u[j] - r[j]);
    for(int i = 0; i < 36; i++) Lxx[i] = 0.0;
    for(int a = 0; a < 6; a++) Lxx[a * 6 + a] = wx[a];
    for(int i = 0; i < 9; i++) Luu[i] = 0.0;
    for(int j = 0; j < 3; j++) Luu[j * 3 + j] = w[j];
    for(int i = 0; i < 18; i++) Lxu[i] = 0.0;
  }

  void terminalCostDeriv(const double * x, double * Vx, double * Vxx) const override
  {
    double w[6], r[6];
    termWeights(w, r);
    for(int a = 0; a < 6; a++) Vx[a] = w[a] * (x[a] - r[a]);
    for(int i = 0; i < 36; i++) Vxx[i] = 0.0;
    for(int a = 0; a < 6; a++) Vxx[a * 6 + a] = w[a];
  }
};
} // namespace oracle
