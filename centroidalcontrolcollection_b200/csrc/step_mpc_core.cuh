// step_mpc_core.cuh — StepMpc1d::planOnce (reference src/StepMpc.cpp:27-193) for one problem and one axis as plain scalar
// code, shared by the kernel (step_mpc.cu) and the CPU shim of the C++ tests (tests/cpp/oracle_engine_shim.cpp compiles it
// for the host).  n unknowns (one per support phase), n <= kStepMpcMax.
//
// The reference condenses the phases with VariantSequentialExtension<2>(model_list, true) into A_seq (3n x 2) and B_seq
// (3n x n) — outputs (position, velocity, capture point) after each phase — and adds, per objective term, B_seq' S' S B_seq
// to eq_mat with a 0 / 1 selection matrix S.  Every such term is a sum over the selected output rows r of w r r' (and w r a
// for eq_vec, a = the row's A_seq x0 minus its reference), so the system is accumulated row by row while the step model is
// propagated phase by phase; nothing of size 3n x n is stored.
#pragma once
#ifdef __CUDACC__
#  define CCC_STEP_HD __host__ __device__
#else
#  include <cmath>
#  define CCC_STEP_HD
#endif

namespace ccc_step
{
constexpr int kStepMpcMax = 16;

struct Weights
{
  double free_zmp, fixed_zmp, double_support, pos, vel, capture_point_abs, capture_point_rel;
};

/** eq_mat += w r r', eq_vec += w r a  (row r over the n unknowns). */
CCC_STEP_HD inline void rank_one(double * M, double * v, int n, const double * r, double w, double a)
{
  for(int i = 0; i < n; i++)
  {
    const double wr = w * r[i];
    for(int j = 0; j < n; j++) M[i * kStepMpcMax + j] += wr * r[j];
    v[i] += wr * a;
  }
}

/** Solves the 1-D problem; returns sol[0] in current_zmp and, if a future single-support phase exists, its unknown in
 *  next_foot_zmp (has_next).  single / zmp / end_time: the n elements (zmp with stride zmp_stride).  false: singular system. */
CCC_STEP_HD inline bool plan_1d(int n,
                                const int * single,
                                const double * zmp,
                                int zmp_stride,
                                const double * end_time,
                                double current_time,
                                double com_height,
                                const Weights & w,
                                double x0_pos,
                                double x0_vel,
                                double & current_zmp,
                                double & next_foot_zmp,
                                bool & has_next)
{
  double M[kStepMpcMax * kStepMpcMax], v[kStepMpcMax];
  for(int i = 0; i < n; i++)
  {
    v[i] = 0.0;
    for(int j = 0; j < n; j++) M[i * kStepMpcMax + j] = 0.0;
  }
  // ZMP terms (:55-64)
  for(int i = 0; i < n; i++)
  {
    double d = w.free_zmp;
    if(i == 0 || (i == 1 && n > 1 && !single[0])) d = w.fixed_zmp;
    M[i * kStepMpcMax + i] += d;
    v[i] += -1 * d * zmp[i * zmp_stride];
  }
  // double-support smoothness (:66-80)
  for(int i = 1; i + 1 < n; i++)
    if(single[i - 1] && !single[i] && single[i + 1])
    {
      const double blk[3] = {1.0, -2.0, 1.0};
      for(int a = 0; a < 3; a++)
        for(int b = 0; b < 3; b++) M[(i - 1 + a) * kStepMpcMax + (i - 1 + b)] += w.double_support * (blk[a] * blk[b]);
    }
  int n_future_single = 0;
  for(int i = 1; i < n; i++) n_future_single += single[i] ? 1 : 0;
  // propagate the step model: x_i = Ad_i x_{i-1} + Bd_i u_i, as (A_x(i) x0, B_x(i))
  const double omega = sqrt(9.80665 / com_height);
  double ax[2] = {x0_pos, x0_vel};
  double bx[2][kStepMpcMax];
  double cp_prev[kStepMpcMax], cp_prev_a = 0.0; // capture-point output row of the previous phase and its A_seq x0
  for(int j = 0; j < n; j++) bx[0][j] = bx[1][j] = cp_prev[j] = 0.0;
  for(int i = 0; i < n; i++)
  {
    const double dur = end_time[i] - (i == 0 ? current_time : end_time[i - 1]);
    const double e = exp(omega * dur), ei = 1.0 / e;
    const double a00 = 0.5 * (e + ei), a01 = 0.5 * (e - ei) / omega, a10 = 0.5 * omega * (e - ei), a11 = 0.5 * (e + ei);
    const double b0 = 1.0 - 0.5 * (e + ei), b1 = 0.5 * omega * (ei - e);
    const double n0 = a00 * ax[0] + a01 * ax[1], n1 = a10 * ax[0] + a11 * ax[1];
    ax[0] = n0;
    ax[1] = n1;
    for(int j = 0; j < i; j++)
    {
      const double m0 = a00 * bx[0][j] + a01 * bx[1][j], m1 = a10 * bx[0][j] + a11 * bx[1][j];
      bx[0][j] = m0;
      bx[1][j] = m1;
    }
    bx[0][i] = b0;
    bx[1][i] = b1;
    // outputs of phase i: position row bx[0], velocity row bx[1], capture point row bx[0] + bx[1] / omega (C of :21-22)
    double cp[kStepMpcMax];
    for(int j = 0; j < n; j++) cp[j] = bx[0][j] + (1.0 / omega) * bx[1][j];
    const double cp_a = ax[0] + (1.0 / omega) * ax[1];
    if(w.pos > 0.0) rank_one(M, v, n, bx[0], w.pos, ax[0] - zmp[(i + 1 < n ? i + 1 : n - 1) * zmp_stride]); // :82-96
    if(w.vel > 0.0) rank_one(M, v, n, bx[1], w.vel, ax[1]);                                                   // :98-110
    if(w.capture_point_abs > 0.0) // :112-149
    {
      if(n_future_single == 0)
      {
        if(i >= 1 || !single[i]) rank_one(M, v, n, cp, w.capture_point_abs, cp_a - zmp[i * zmp_stride]);
      }
      else if(i >= 1 && single[i])
        rank_one(M, v, n, cp_prev, w.capture_point_abs, cp_prev_a - zmp[i * zmp_stride]);
    }
    if(w.capture_point_rel > 0.0 && n_future_single >= 1 && i >= 1 && single[i]) // :151-178
    {
      double q[kStepMpcMax];
      for(int j = 0; j < n; j++) q[j] = cp_prev[j] - (j == i ? 1.0 : 0.0);
      rank_one(M, v, n, q, w.capture_point_rel, cp_prev_a);
    }
    for(int j = 0; j < n; j++) cp_prev[j] = cp[j];
    cp_prev_a = cp_a;
  }
  // eq_mat sol = -eq_vec by Gaussian elimination with partial pivoting (:181)
  double sol[kStepMpcMax];
  for(int i = 0; i < n; i++) sol[i] = -1 * v[i];
  for(int c = 0; c < n; c++)
  {
    int piv = c;
    double best = fabs(M[c * kStepMpcMax + c]);
    for(int r = c + 1; r < n; r++)
      if(fabs(M[r * kStepMpcMax + c]) > best)
      {
        best = fabs(M[r * kStepMpcMax + c]);
        piv = r;
      }
    if(!(best > 0.0)) return false;
    if(piv != c)
    {
      for(int j = 0; j < n; j++)
      {
        const double t = M[c * kStepMpcMax + j];
        M[c * kStepMpcMax + j] = M[piv * kStepMpcMax + j];
        M[piv * kStepMpcMax + j] = t;
      }
      const double t = sol[c];
      sol[c] = sol[piv];
      sol[piv] = t;
    }
    for(int r = c + 1; r < n; r++)
    {
      const double f = M[r * kStepMpcMax + c] / M[c * kStepMpcMax + c];
      for(int j = c; j < n; j++) M[r * kStepMpcMax + j] -= f * M[c * kStepMpcMax + j];
      sol[r] -= f * sol[c];
    }
  }
  for(int r = n - 1; r >= 0; r--)
  {
    double acc = sol[r];
    for(int j = r + 1; j < n; j++) acc -= M[r * kStepMpcMax + j] * sol[j];
    sol[r] = acc / M[r * kStepMpcMax + r];
  }
  current_zmp = sol[0];
  has_next = false;
  next_foot_zmp = 0.0;
  for(int i = 1; i < n; i++)
    if(single[i])
    {
      next_foot_zmp = sol[i];
      has_next = true;
      break;
    }
  return true;
}
} // namespace ccc_step
