// oracle/xy.hpp — CPU restatement of CCC::LinearMpcXY::planOnce after callback sampling, for a sweep of schedules.
//
// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Follows reference src/LinearMpcXY.cpp:59-83 (Model), :102-110 (per-stage discretisation), :116-181 (procOnce),
// include/CCC/VariantSequentialExtension.h:110-186 (condensing) and include/CCC/StateSpaceModel.h:164-216
// (calcDiscMatrix).  The model's A is nilpotent (A^3 = 0) and has no offset vector, so the matrix exponential the
// reference evaluates numerically is the finite polynomial written out below; the numpy / scipy
// `expm` path of centroidalcontrolcollection_b200/linear_models.py is the independent check of that statement
// (tests/test_linear_mpc_xy_cpu.py).
//
// Canonical arithmetic, which the engine's kernels reproduce:
//   a = total_force_z / mass; h2 = (dt dt) 0.5; h3 = ((dt dt) dt) / 6;
//   B column of ridge (rx, ry, rz) at vertex (vx, vy, vz): hz = vz - com_z,
//     (0, rx, 0, ry, fma(vy, rz, -(hz ry)), fma(hz, rx, -(vx rz)));
//   Bd column = (h2 rx, dt rx, h2 ry, dt ry, fma(dt, b4, (-(a h3)) ry), fma(dt, b5, (a h3) rx));
//   every 6 x 6 product (Ad_j x, Ad_i A_seq_{i-1}) is a dense sequential fma chain over k = 0..5 from +0.0;
//   obj_mat(i, j) = chain over r = 0..6N-1 of (w[r % 6] B_seq(r, i)) B_seq(r, j), then + w_force on the diagonal;
//   obj_vec(j) = -(chain over r of (w[r % 6] B_seq(r, j)) resid(r)), resid(r) = ref(r) - chain_c A_seq(r, c) x0(c).
#pragma once
#include <vector>

#include "../include/ccc_b200.h"
#include "num.hpp"
#include "qp.hpp"

namespace oracle
{
struct XySchedule
{
  int N = 0, n = 0, n_eq = 0;
  std::vector<int> m, off;           // [N] input dimension and first column of each stage
  std::vector<double> Ad;            // [N][6][6]
  std::vector<double> A_seq, B_seq;  // [6N][6], [6N][n]
  std::vector<double> H;             // [n][n]
  std::vector<double> Aeq, beq;      // [n_eq][n], [n_eq]

  /** Sample tables of schedule s of `bt`. */
  void build(const ccc_linear_mpc_xy_batch_t & bt, int s)
  {
    N = bt.horizon_steps;
    const int mm = bt.m_max;
    const double dt = bt.dt;
    const double h2 = (dt * dt) * 0.5, h3 = ((dt * dt) * dt) / 6.0;
    m.assign(N, 0);
    off.assign(N, 0);
    n = 0;
    n_eq = 0;
    for(int k = 0; k < N; k++)
    {
      m[k] = bt.m[s * N + k];
      off[k] = n;
      n += m[k];
      if(m[k] > 0) n_eq++;
    }
    Ad.assign(static_cast<size_t>(N) * 36, 0.0);
    for(int k = 0; k < N; k++)
    {
      double * A = Ad.data() + k * 36;
      const double a = bt.total_force_z[s * N + k] / bt.mass;
      for(int i = 0; i < 6; i++) A[i * 6 + i] = 1.0;
      A[0 * 6 + 1] = dt;
      A[2 * 6 + 3] = dt;
      A[4 * 6 + 2] = -(a * dt);
      A[5 * 6 + 0] = a * dt;
      A[4 * 6 + 3] = -(a * h2);
      A[5 * 6 + 1] = a * h2;
    }
    const int rows = 6 * N;
    A_seq.assign(static_cast<size_t>(rows) * 6, 0.0);
    B_seq.assign(static_cast<size_t>(rows) * n, 0.0);
    for(int i = 0; i < N; i++)
    {
      const double * A = Ad.data() + i * 36;
      for(int r = 0; r < 6; r++)
        for(int c = 0; c < 6; c++)
          A_seq[(i * 6 + r) * 6 + c] = i == 0 ? A[r * 6 + c] : dot_seq(A + r * 6, 1, A_seq.data() + ((i - 1) * 6) * 6 + c, 6, 6);
    }
    for(int i = 0; i < N; i++)
    {
      const double a = bt.total_force_z[s * N + i] / bt.mass;
      const double ah3 = a * h3;
      const double cz = bt.com_z[s * N + i];
      for(int l = 0; l < m[i]; l++)
      {
        const double * rg = bt.ridge + (static_cast<size_t>(s * N + i) * mm + l) * 3;
        const double * vt = bt.vertex + (static_cast<size_t>(s * N + i) * mm + l) * 3;
        const double hz = vt[2] - cz;
        const double b4 = fmad(vt[1], rg[2], -(hz * rg[1]));
        const double b5 = fmad(hz, rg[0], -(vt[0] * rg[2]));
        double col[6] = {h2 * rg[0], dt * rg[0], h2 * rg[1], dt * rg[1], fmad(dt, b4, (-ah3) * rg[1]), fmad(dt, b5, ah3 * rg[0])};
        const int c = off[i] + l;
        for(int j = i; j < N; j++)
        {
          if(j > i)
          {
            double nx[6];
            for(int r = 0; r < 6; r++) nx[r] = dot_seq(Ad.data() + j * 36 + r * 6, 1, col, 1, 6);
            for(int r = 0; r < 6; r++) col[r] = nx[r];
          }
          for(int r = 0; r < 6; r++) B_seq[static_cast<size_t>(j * 6 + r) * n + c] = col[r];
        }
      }
    }
    // obj_mat
    H.assign(static_cast<size_t>(n) * n, 0.0);
    std::vector<double> WB(static_cast<size_t>(rows) * n);
    for(int r = 0; r < rows; r++)
      for(int j = 0; j < n; j++) WB[static_cast<size_t>(r) * n + j] = bt.w_output[r % 6] * B_seq[static_cast<size_t>(r) * n + j];
    for(int i = 0; i < n; i++)
      for(int j = 0; j < n; j++)
      {
        double acc = dot_seq(WB.data() + i, n, B_seq.data() + j, n, rows);
        if(i == j) acc = acc + bt.w_force;
        H[static_cast<size_t>(i) * n + j] = acc;
      }
    // equalities: total vertical force of every contact stage (src :149-176)
    Aeq.assign(static_cast<size_t>(n_eq) * n, 0.0);
    beq.assign(n_eq, 0.0);
    int e = 0;
    for(int k = 0; k < N; k++)
    {
      if(m[k] == 0) continue;
      for(int l = 0; l < m[k]; l++) Aeq[static_cast<size_t>(e) * n + off[k] + l] = bt.ridge[(static_cast<size_t>(s * N + k) * mm + l) * 3 + 2];
      beq[e] = bt.total_force_z[s * N + k];
      e++;
    }
  }

  /** obj_vec of one initial state (src :145-146). */
  void objVec(const ccc_linear_mpc_xy_batch_t & bt, int s, const double * x0, double * g) const
  {
    const int rows = 6 * N;
    std::vector<double> resid(rows);
    for(int r = 0; r < rows; r++) resid[r] = bt.ref_output[static_cast<size_t>(s) * rows + r] - dot_seq(A_seq.data() + r * 6, 1, x0, 1, 6);
    for(int j = 0; j < n; j++)
    {
      double acc = 0.0;
      for(int r = 0; r < rows; r++) acc = fmad(bt.w_output[r % 6] * B_seq[static_cast<size_t>(r) * n + j], resid[r], acc);
      g[j] = -acc;
    }
  }
};
} // namespace oracle
