// model_srb.cuh — CCC::DdpSingleRigidBody::DdpProblem as a model policy of the warp DDP core.
//
// state x = (c, ZYX Euler angles, v, omega) — 12 states; input u = ridge force scales.
// Replaces (reference src/DdpSingleRigidBody.cpp): matAngularVelToEulerDot :26-38, stateEq :52-91,
// calcStateEqDeriv :115-185.  Evaluation order: oracle/srb.hpp (scalar formulas verbatim, no
// contraction; sincos_canon; the 3x3 inertia LL^T is factorised once per schedule stage by
// srb_pack_inertia_kernel and stored in row 6 of the stage table).
#pragma once
#include "ddp_warp_core.cuh"

namespace ccc
{
/** Per-stage inertia constants, lanes 0..14 of table row 6: I (9, row-major), l10, l20, l21,
 *  1/l00, 1/l11, 1/l22 of its LL^T. */
struct Inertia3
{
  double I[9], l10, l20, l21, i0, i1, i2;

  CCC_DEV void load(const double * row)
  {
    CCC_UNROLL
    for(int i = 0; i < 9; i++) I[i] = ldg(row + i);
    l10 = ldg(row + 9);
    l20 = ldg(row + 10);
    l21 = ldg(row + 11);
    i0 = ldg(row + 12);
    i1 = ldg(row + 13);
    i2 = ldg(row + 14);
  }

  /** LL^T factor in the order of oracle DenseLlt::compute (nf = 3; the 3x3 inertia keeps the plain LL^T). */
  CCC_DEV void factor()
  {
    i0 = drcp(dsqrt(I[0]));
    l10 = I[3] * i0;
    l20 = I[6] * i0;
    i1 = drcp(dsqrt(dfma(-l10, l10, I[4])));
    l21 = dfma(-l20, l10, I[7]) * i1;
    i2 = drcp(dsqrt(dfma(-l21, l21, dfma(-l20, l20, I[8]))));
  }

  /** b <- I^-1 b in the order of oracle FreeLlt::solve. */
  CCC_DEV void solve(double (&b)[3]) const
  {
    const double y0 = b[0] * i0;
    const double y1 = dfma(-l10, y0, b[1]) * i1;
    const double y2 = dfma(-l21, y1, dfma(-l20, y0, b[2])) * i2;
    const double x2 = y2 * i2;
    const double x1 = dfma(-l21, x2, y1) * i1;
    const double x0 = dfma(-l10, x1, dfma(-l20, x2, y0)) * i0;
    b[0] = x0;
    b[1] = x1;
    b[2] = x2;
  }
};

struct SrbModel
{
  static constexpr int NX = 12;
#ifndef CCC_SRB_STAGE_UNROLL
#  define CCC_SRB_STAGE_UNROLL 3
#endif
  static constexpr int STAGE_UNROLL = CCC_SRB_STAGE_UNROLL; // once-per-stage state-sized loops (ddp_warp_core.cuh), rolled three at a time: instruction-cache footprint
  static constexpr int NXP = 14;     // even row stride of the K / QuuK / Qux staging buffers: 12 gains + k
  static constexpr int R0 = 6;       // Fu is non-zero in rows 6..11
  static constexpr int NREF = 6;     // referenced states: position and orientation
  static constexpr int TAB_ROWS = 7; // ridge xyz, vertex xyz, inertia constants
  struct Params
  {
    double dt, mass;
  };

  /** matAngularVelToEulerDot (:26-38).  (Inline at both call sites: as an out-of-line function shared by the rollouts and
   *  the backward pass it measured 9 % slower, profiles/r02g_ab_srb_euler.txt.) */
  CCC_DEV static void euler_mat(double o0, double o1, double (&E)[9], double & sa, double & ca, double & sb, double & cb)
  {
    sincos_canon(o0, sa, ca);
    sincos_canon(o1, sb, cb);
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

  /** Stage-independent part of Fx: identity + dt on the (c, v) block (:131, :177-178). */
  template<class W>
  CCC_DEV static void init_Fx(W & w)
  {
    CCC_NOUNROLL
    for(int e = w.lane; e < W::sm::NN; e += 32)
    {
      const int i = e / NX, j = e - NX * i;
      double v = (i == j && e < NX * NX) ? 1.0 : 0.0;
      if(i < 3 && j == i + 6) v = 1.0 * w.P.mp.dt;
      w.s[W::sm::FX + e] = v;
    }
  }

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
    Inertia3 in;
    in.load(tb + 6 * 32);
    double E[9], sa, ca, sb, cb;
    euler_mat(x[3], x[4], E, sa, ca, sb, cb);
    double xdot[12];
    CCC_UNROLL
    for(int a = 0; a < 3; a++)
    {
      xdot[a] = x[6 + a];
      double acc = 0.0;
      CCC_UNROLL
      for(int c = 0; c < 3; c++) acc = dfma(E[3 * a + c], x[9 + c], acc);
      xdot[3 + a] = acc;
      xdot[6 + a] = ddiv(r7[a], mass);
    }
    xdot[8] = ddiv(r7[2], mass) + (-1 * 9.80665);
    double Iw[3], cw[3], rhs[3];
    CCC_UNROLL
    for(int a = 0; a < 3; a++)
    {
      double acc = 0.0;
      CCC_UNROLL
      for(int c = 0; c < 3; c++) acc = dfma(in.I[3 * a + c], x[9 + c], acc);
      Iw[a] = acc;
    }
    const double wv[3] = {x[9], x[10], x[11]};
    cross3(wv, Iw, cw);
    CCC_UNROLL
    for(int a = 0; a < 3; a++) rhs[a] = (-cw[a]) + r7[3 + a];
    in.solve(rhs);
    CCC_UNROLL
    for(int a = 0; a < 3; a++) xdot[9 + a] = rhs[a];
    CCC_UNROLL
    for(int i = 0; i < 12; i++) x[i] = dfma(dt, xdot[i], x[i]);
    return r7[6];
  }

  /** This lane's column of Fu (rows 6..8: rho / mass dt; rows 9..11: I^-1 ((p - c) x rho) dt) and
   *  the stage-dependent entries of Fx (blocks (3,3), (3,9), (9,0), (9,9)) (:115-185). */
  template<class W>
  CCC_DEV static void lane_derivs(W & w, int k, int m, const double * xn, double u, double (&Fu)[6])
  {
    const int lane = w.lane;
    const bool active = lane < m;
    const double dt = w.P.mp.dt, mass = w.P.mp.mass;
    double x[12];
    CCC_UNROLL
    for(int i = 0; i < 12; i++) x[i] = xn[i];
    const double * tb = w.stage_tab(k);
    double rho[3], d[3], cr[3];
    CCC_UNROLL
    for(int a = 0; a < 3; a++)
    {
      rho[a] = active ? ldg(tb + a * 32 + lane) : 0.0;
      d[a] = (active ? ldg(tb + (3 + a) * 32 + lane) : 0.0) - x[a];
    }
    cross3(d, rho, cr);
    double f[3];
    CCC_UNROLL
    for(int a = 0; a < 3; a++) f[a] = active ? u * rho[a] : 0.0;
    warp_sum_n<3>(f);
    Inertia3 in;
    in.load(tb + 6 * 32);

    double E[9], sin_alpha, cos_alpha, sin_beta, cos_beta;
    euler_mat(x[3], x[4], E, sin_alpha, cos_alpha, sin_beta, cos_beta);
    // SymPy-derived blocks, reference src/DdpSingleRigidBody.cpp:136-161 (verbatim expressions)
    const double w1 = x[9], w2 = x[10], w3 = x[11];
    const double cos_beta_2 = cos_beta * cos_beta;
    const double sin_beta_2 = sin_beta * sin_beta;
    const double I11 = in.I[0], I12 = in.I[1], I13 = in.I[2], I22 = in.I[4], I23 = in.I[5], I33 = in.I[8];
    const double d33 = -w1 * sin_alpha * sin_beta / cos_beta + w2 * sin_beta * cos_alpha / cos_beta;
    const double d43 = -w1 * cos_alpha - w2 * sin_alpha;
    const double d53 = -w1 * sin_alpha / cos_beta + w2 * cos_alpha / cos_beta;
    const double d34 = w1 * sin_beta_2 * cos_alpha / cos_beta_2 + w1 * cos_alpha + w2 * sin_alpha * sin_beta_2 / cos_beta_2
                       + w2 * sin_alpha;
    const double d54 = w1 * sin_beta * cos_alpha / cos_beta_2 + w2 * sin_alpha * sin_beta / cos_beta_2;
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
    const double cm[9] = {0, -f[2], f[1], f[2], 0, -f[0], -f[1], f[0], 0};
    double b99[9], b90[9];
    CCC_UNROLL
    for(int c = 0; c < 3; c++)
    {
      double col[3] = {Mw[c], Mw[3 + c], Mw[6 + c]};
      in.solve(col);
      b99[c] = col[0];
      b99[3 + c] = col[1];
      b99[6 + c] = col[2];
      double col2[3] = {cm[c], cm[3 + c], cm[6 + c]};
      in.solve(col2);
      b90[c] = col2[0];
      b90[3 + c] = col2[1];
      b90[6 + c] = col2[2];
    }
    warp_sync();
    if(lane == 0)
    {
      double * Fx = w.s + W::sm::FX;
      Fx[3 * 12 + 3] = d33 * dt + 1.0;
      Fx[4 * 12 + 3] = d43 * dt;
      Fx[5 * 12 + 3] = d53 * dt;
      Fx[3 * 12 + 4] = d34 * dt;
      Fx[4 * 12 + 4] = 0.0 * dt + 1.0;
      Fx[5 * 12 + 4] = d54 * dt;
      CCC_UNROLL
      for(int a = 0; a < 3; a++)
        CCC_UNROLL
        for(int c = 0; c < 3; c++)
        {
          Fx[(3 + a) * 12 + 9 + c] = E[3 * a + c] * dt;
          Fx[(9 + a) * 12 + c] = b90[3 * a + c] * dt;
          Fx[(9 + a) * 12 + 9 + c] = a == c ? b99[3 * a + c] * dt + 1.0 : b99[3 * a + c] * dt;
        }
    }
    warp_sync();
    in.solve(cr);
    CCC_UNROLL
    for(int a = 0; a < 3; a++)
    {
      Fu[a] = ddiv(rho[a], mass) * dt;
      Fu[3 + a] = cr[a] * dt;
    }
  }
};
} // namespace ccc
