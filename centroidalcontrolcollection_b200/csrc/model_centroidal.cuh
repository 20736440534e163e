// model_centroidal.cuh — CCC::DdpCentroidal::DdpProblem as a model policy of the warp DDP core.
//
// state x = (c, P, L) (CoM, linear momentum, angular momentum), input u = ridge force scales.
// Replaces (reference src/DdpCentroidal.cpp): stateEq :32-64, calcStateEqDeriv :85-121.
// Evaluation order: oracle/centroidal.hpp.
#pragma once
#include "ddp_warp_core.cuh"

namespace ccc
{
struct CentroidalModel
{
  static constexpr int NX = 9;       // states
  static constexpr int STAGE_UNROLL = 9; // once-per-stage state-sized loops (ddp_warp_core.cuh): fully unrolled
  static constexpr int NXP = 10;     // even row stride of the K / QuuK / Qux staging buffers
  static constexpr int R0 = 3;       // Fu is non-zero in rows 3..8
  static constexpr int NREF = 3;     // referenced states: CoM position
  static constexpr int TAB_ROWS = 6; // stage table rows: ridge xyz, vertex xyz
  struct Params
  {
    double dt, mass;
  };

  /** Input cost: 0.5 w_force |u|^2, so Lu = w_force u and Luu = w_force I. */
  template<class W>
  CCC_DEV static double lu(const W & w, int, double u)
  {
    return w.P.w_run[NX] * u;
  }
  template<class W>
  CCC_DEV static double luu(const W & w)
  {
    return w.P.w_run[NX];
  }

  /** Stage-independent part of Fx: identity + (1/mass) dt on the (c, P) block (:100, :118-119). */
  template<class W>
  CCC_DEV static void init_Fx(W & w)
  {
    CCC_NOUNROLL
    for(int e = w.lane; e < W::sm::NN; e += 32)
    {
      const int i = e / NX, j = e - NX * i;
      double v = (i == j && e < NX * NX) ? 1.0 : 0.0;
      if(i < 3 && j == i + 3) v = ddiv(1, w.P.mp.mass) * w.P.mp.dt;
      w.s[W::sm::FX + e] = v;
    }
  }

  /** x <- x + dt * xdot(x, u) (:32-64); returns sum_j u_j^2 for the running cost. */
  template<class W>
  CCC_DEV static double step(W & w, int k, int m, double (&x)[NX], double u)
  {
    const int lane = w.lane;
    const bool active = lane < m;
    const double * tb = w.stage_tab(k);
    double rho[3], d[3], cr[3];
    CCC_UNROLL
    for(int a = 0; a < 3; a++)
    {
      rho[a] = active ? ldg(tb + a * 32 + lane) : 0.0;
      d[a] = (active ? ldg(tb + (3 + a) * 32 + lane) : 0.0) - x[a];
    }
    cross3(d, rho, cr);
    // seven ridge reductions in one transpose-reduce: force (3), moment (3), |u|^2
    double r7[8];
    CCC_UNROLL
    for(int a = 0; a < 3; a++)
    {
      r7[a] = active ? u * rho[a] : 0.0;
      r7[3 + a] = active ? u * cr[a] : 0.0;
    }
    r7[6] = active ? u * u : 0.0;
    r7[7] = 0.0;
    warp_sum8(r7, w.s + W::sm::S2);
    const double mass = w.P.mp.mass, dt = w.P.mp.dt;
    const double inv_mass = ddiv(1, mass); // loop invariant; the same rounded 1 / m as init_Fx
    double xdot[9];
    CCC_UNROLL
    for(int a = 0; a < 3; a++)
    {
      xdot[a] = x[3 + a] * inv_mass;
      xdot[3 + a] = r7[a];
      xdot[6 + a] = r7[3 + a];
    }
    xdot[5] = r7[2] + (-1 * mass * 9.80665);
    CCC_UNROLL
    for(int i = 0; i < 9; i++) x[i] = dfma(dt, xdot[i], x[i]);
    return r7[6];
  }

  /** This lane's column of Fu (rows 3..8: rho dt, ((p - c) x rho) dt) and the stage-dependent block
   *  of Fx: crossMat(total force) dt in rows 6..8, columns 0..2 (:85-121). */
  template<class W>
  CCC_DEV static void lane_derivs(W & w, int k, int m, const double * xn, double u, double (&Fu)[6])
  {
    const int lane = w.lane;
    const bool active = lane < m;
    const double dt = w.P.mp.dt;
    double x3[3];
    CCC_UNROLL
    for(int i = 0; i < 3; i++) x3[i] = xn[i];
    const double * tb = w.stage_tab(k);
    double rho[3], d[3], cr[3];
    CCC_UNROLL
    for(int a = 0; a < 3; a++)
    {
      rho[a] = active ? ldg(tb + a * 32 + lane) : 0.0;
      d[a] = (active ? ldg(tb + (3 + a) * 32 + lane) : 0.0) - x3[a];
    }
    cross3(d, rho, cr);
    double f[3];
    CCC_UNROLL
    for(int a = 0; a < 3; a++) f[a] = active ? u * rho[a] : 0.0;
    warp_sum_n<3>(f);
    warp_sync();
    if(lane == 0)
    {
      double * Fx = w.s + W::sm::FX;
      Fx[6 * 9 + 1] = -f[2] * dt;
      Fx[6 * 9 + 2] = f[1] * dt;
      Fx[7 * 9 + 0] = f[2] * dt;
      Fx[7 * 9 + 2] = -f[0] * dt;
      Fx[8 * 9 + 0] = -f[1] * dt;
      Fx[8 * 9 + 1] = f[0] * dt;
    }
    warp_sync();
    CCC_UNROLL
    for(int a = 0; a < 3; a++)
    {
      Fu[a] = rho[a] * dt;
      Fu[3 + a] = cr[a] * dt;
    }
  }
};
} // namespace ccc
