// model_zmp.cuh — CCC::DdpZmp::DdpProblem as a model policy of the warp DDP core.
//
// state x = (c_x, v_x, c_y, v_y, c_z, v_z), input u = (zmp_x, zmp_y, f_z): 3 inputs, no limits
// (the unconstrained path of the core).  Replaces (reference src/DdpZmp.cpp): stateEq :8-18, costs :20-41,
// calcStateEqDeriv :43-72, cost derivatives :74-143.  Evaluation order: oracle/zmp.hpp (scalar formulas
// verbatim, no contraction).  Only 3 of the 32 lanes carry an input: this instantiation exists for
// coverage of the class, it is not a throughput path.
//
// Stage table row 6, lanes 0..3: u_ref = (ref zmp x, ref zmp y, mass g) and ref zmp z of the stage.
#pragma once
#include "ddp_warp_core.cuh"

namespace ccc
{
struct ZmpModel
{
  static constexpr int NX = 6;
  static constexpr int STAGE_UNROLL = 6; // once-per-stage state-sized loops (ddp_warp_core.cuh): fully unrolled
  static constexpr int NXP = 8;
  static constexpr int R0 = 0;       // Fu column = all six rows
  static constexpr int NREF = 6;     // every state has a (possibly zero-weighted) reference
  static constexpr int TAB_ROWS = 7;
  struct Params
  {
    double dt, mass;
    double w_u[3]; // running_zmp, running_zmp, running_force_z
  };

  template<class W>
  CCC_DEV static double lu(const W & w, int k, double u)
  {
    const int lane = w.lane;
    if(lane >= 3) return 0.0;
    const double uref = ldg(w.stage_tab(k) + 6 * 32 + lane);
    const double wj = lane == 0 ? w.P.mp.w_u[0] : lane == 1 ? w.P.mp.w_u[1] : w.P.mp.w_u[2];
    return wj * (u - uref);
  }
  template<class W>
  CCC_DEV static double luu(const W & w)
  {
    const int lane = w.lane;
    return lane == 0 ? w.P.mp.w_u[0] : lane == 1 ? w.P.mp.w_u[1] : lane == 2 ? w.P.mp.w_u[2] : 0.0;
  }

  /** Stage-independent part of Fx: identity + dt on (0,1), (2,3), (4,5) (:52-57, :60-61). */
  template<class W>
  CCC_DEV static void init_Fx(W & w)
  {
    CCC_NOUNROLL
    for(int e = w.lane; e < W::sm::NN; e += 32)
    {
      const int i = e / NX, j = e - NX * i;
      double v = (i == j && e < NX * NX) ? 1.0 : 0.0;
      if((i == 0 || i == 2 || i == 4) && j == i + 1) v = 1 * w.P.mp.dt;
      w.s[W::sm::FX + e] = v;
    }
  }

  /** x <- x + dt xdot (:8-18); returns sum_j w_j (u_j - u_ref_j)^2 (the input part of the running cost). */
  template<class W>
  CCC_DEV static double step(W & w, int k, int m, double (&x)[NX], double u)
  {
    const int lane = w.lane;
    const double * row = w.stage_tab(k) + 6 * 32;
    const double u0 = warp_shfl(u, 0), u1 = warp_shfl(u, 1), u2 = warp_shfl(u, 2);
    const double zmp_z = ldg(row + 3);
    double term = 0.0;
    if(lane < 3 && lane < m)
    {
      const double dd = u - ldg(row + lane);
      const double wj = lane == 0 ? w.P.mp.w_u[0] : lane == 1 ? w.P.mp.w_u[1] : w.P.mp.w_u[2];
      term = wj * (dd * dd);
    }
    const double S = warp_sum(term);
    const double mass = w.P.mp.mass, dt = w.P.mp.dt;
    double xdot[6];
    xdot[0] = x[1];
    xdot[1] = (x[0] - u0) * u2 / (mass * (x[4] - zmp_z));
    xdot[2] = x[3];
    xdot[3] = (x[2] - u1) * u2 / (mass * (x[4] - zmp_z));
    xdot[4] = x[5];
    xdot[5] = u2 / mass - 9.80665;
    CCC_UNROLL
    for(int i = 0; i < 6; i++) x[i] = dfma(dt, xdot[i], x[i]);
    return S;
  }

  /** Fu column of this lane and the state-dependent entries of Fx (:43-72). */
  template<class W>
  CCC_DEV static void lane_derivs(W & w, int k, int m, const double * xn, double u, double (&Fu)[6])
  {
    const int lane = w.lane;
    const double dt = w.P.mp.dt, mass = w.P.mp.mass;
    double x[6];
    CCC_UNROLL
    for(int i = 0; i < 6; i++) x[i] = xn[i];
    const double u0 = warp_shfl(u, 0), u1 = warp_shfl(u, 1), u2 = warp_shfl(u, 2);
    const double zmp_z = ldg(w.stage_tab(k) + 6 * 32 + 3);
    const double f10 = u2 / (mass * (x[4] - zmp_z));
    const double f14 = -1 * (x[0] - u0) * u2 / (mass * ((x[4] - zmp_z) * (x[4] - zmp_z)));
    const double f34 = -1 * (x[2] - u1) * u2 / (mass * ((x[4] - zmp_z) * (x[4] - zmp_z)));
    warp_sync();
    if(lane == 0)
    {
      double * Fx = w.s + W::sm::FX;
      Fx[1 * 6 + 0] = f10 * dt;
      Fx[1 * 6 + 4] = f14 * dt;
      Fx[3 * 6 + 2] = f10 * dt;
      Fx[3 * 6 + 4] = f34 * dt;
    }
    warp_sync();
    CCC_UNROLL
    for(int i = 0; i < 6; i++) Fu[i] = 0.0;
    if(lane == 0) Fu[1] = (-1 * u2 / (mass * (x[4] - zmp_z))) * dt;
    if(lane == 1) Fu[3] = (-1 * u2 / (mass * (x[4] - zmp_z))) * dt;
    if(lane == 2)
    {
      Fu[1] = ((x[0] - u0) / (mass * (x[4] - zmp_z))) * dt;
      Fu[3] = ((x[2] - u1) / (mass * (x[4] - zmp_z))) * dt;
      Fu[5] = (1 / mass) * dt;
    }
    (void)m;
  }
};
} // namespace ccc
