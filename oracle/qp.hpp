// oracle/qp.hpp — CPU restatement of the dense QP solve behind QpSolverCollection::QpSolver::solve.
//
// TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED on the iteration path: QpSolverCollection
// (isri-aist/QpSolverCollection, un-versioned find_package at reference CMakeLists.txt:23) and the
// backend it dispatches to (QLD in the reference's CI, .github/workflows/ci-colcon.yaml:102) are not
// under /root/reference.  QLD is Schittkowski's implementation of the Goldfarb-Idnani dual
// active-set method (Goldfarb & Idnani, "A numerically stable dual method for solving strictly
// convex quadratic programs", Math. Programming 27, 1983); this file restates that published
// algorithm.  Every QP on the hot path is strictly convex (H = I, reference src/LinearMpcZmp.cpp:23;
// w_vel I + w_zmp P'P, src/IntrinsicallyStableMpc.cpp:30-33), so the minimiser — and with it the
// final active set of a non-degenerate problem — is unique and solver independent; the order in
// which constraints enter is defined by this restatement (most violated first, lowest index on ties).
//
//   min 0.5 x'Qx + c'x   s.t.  A x = b,  C x <= d          (QpCoeff: obj_mat_, obj_vec_, eq_*, ineq_*)
//
// Call sites: src/LinearMpcZmp.cpp:69 (n = N, 0 eq, 2N ineq), src/IntrinsicallyStableMpc.cpp:93
// (n = N, 1 eq, 2N ineq).  Bounds x_min/x_max = -+1e10 (src/LinearMpcZmp.cpp:26-27) are inactive and
// not represented.  src/LinearMpcXY.cpp:181 (n = sum of ridge counts <= 256, one equality per contact
// stage, no inequalities, finite bounds x_min <= x <= x_max): the caller appends the bounds as
// inequality rows (-e_j x <= -x_min, e_j x <= x_max).  Q, A, C are shared by a batch; c, b, d are per problem.
//
// Canonical arithmetic (num.hpp): sequential fma chains for matrix-vector products, a fixed
// 128-leaf (n <= 128) or 256-leaf pairwise tree for the two scalar products of a step, IEEE / and sqrt.
#pragma once
#include "num.hpp"

#include <limits>

namespace oracle
{
/** Number of leaves of the scalar-product tree: 128 up to n = 128, else 256 (= threads of the CTA that
 *  solves the problem in the engine). */
inline int qp_tree_leaves(int n)
{
  return n <= 128 ? 128 : 256;
}

/** Pairwise tree over `leaves` zero-padded leaves: strides leaves/2, ..., 1. */
inline double tree_sum_leaves(const double * p, int n, int leaves)
{
  if(kTextbook)
  {
    double acc = 0.0;
    for(int i = 0; i < n; i++) acc = acc + p[i];
    return acc;
  }
  double t[256];
  for(int i = 0; i < leaves; i++) t[i] = i < n ? p[i] : 0.0;
  for(int off = leaves / 2; off >= 1; off >>= 1)
    for(int i = 0; i < off; i++) t[i] = t[i] + t[i + off];
  return t[0];
}

inline double tree_sum128(const double * p, int n)
{
  return tree_sum_leaves(p, n, 128);
}

/** sqrt(a^2 + b^2) without overflow, as used for the Givens rotations. */
inline double givens_hypot(double a, double b)
{
  const double a1 = std::fabs(a), b1 = std::fabs(b);
  if(a1 > b1)
  {
    const double t = b1 / a1;
    return a1 * std::sqrt(fmad(t, t, 1.0));
  }
  if(b1 > a1)
  {
    const double t = a1 / b1;
    return b1 * std::sqrt(fmad(t, t, 1.0));
  }
  return a1 * std::sqrt(2.0);
}

struct DenseQpShared
{
  int n = 0, me = 0, mi = 0;
  std::vector<double> Q, A, C; // row-major n x n, me x n, mi x n
  std::vector<double> J;       // L^-T, row-major n x n  (Q = L L')
  bool ok = false;

  /** Factorise Q and form J = L^-T (shared by every problem of the batch). */
  bool setup()
  {
    std::vector<double> L(static_cast<size_t>(n) * n, 0.0), invd(n);
    for(int k = 0; k < n; k++)
    {
      double acc = Q[k * n + k];
      for(int j = 0; j < k; j++) acc = fmad(-L[k * n + j], L[k * n + j], acc);
      if(!(acc > 0.0)) return ok = false;
      const double d = std::sqrt(acc);
      L[k * n + k] = d;
      invd[k] = 1.0 / d;
      for(int i = k + 1; i < n; i++)
      {
        double a = Q[i * n + k];
        for(int j = 0; j < k; j++) a = fmad(-L[i * n + j], L[k * n + j], a);
        L[i * n + k] = a * invd[k];
      }
    }
    // row i of J = L^-1 e_i  (forward substitution)
    J.assign(static_cast<size_t>(n) * n, 0.0);
    std::vector<double> z(n);
    for(int i = 0; i < n; i++)
    {
      for(int r = 0; r < n; r++)
      {
        double acc = r == i ? 1.0 : 0.0;
        for(int j = 0; j < r; j++) acc = fmad(-L[r * n + j], z[j], acc);
        z[r] = acc * invd[r];
      }
      for(int j = 0; j < n; j++) J[i * n + j] = z[j];
    }
    return ok = true;
  }
};

struct DenseQpResult
{
  std::vector<double> x;
  std::vector<int> active; // constraint ids in activation order: 0..me-1 equalities, me + i inequality i
  int iters = 0;
  int status = 0; // 0 solved, 1 infeasible, 2 iteration limit, 3 setup failed, 4 too many active constraints
};

struct DenseQpSolver
{
  const DenseQpShared & S;
  int max_iter = 1000;
  double viol_tol = 1e-10;
  explicit DenseQpSolver(const DenseQpShared & s) : S(s) {}

  /** normal of constraint id (equalities: row of A; inequalities: -row of C) */
  double normal(int id, int j) const { return id < S.me ? S.A[id * S.n + j] : -S.C[(id - S.me) * S.n + j]; }

  DenseQpResult solve(const double * c, const double * b, const double * dvec) const
  {
    const int n = S.n, me = S.me, mi = S.mi;
    const double eps = std::numeric_limits<double>::epsilon();
    DenseQpResult res;
    res.x.assign(n, 0.0);
    if(!S.ok)
    {
      res.status = 3;
      return res;
    }
    std::vector<double> J = S.J, R(static_cast<size_t>(n) * n, 0.0), x(n), z(n), d(n), np(n), r(n), u(n + 1, 0.0), tmp(n);
    std::vector<int> A(n + 1, -1);
    std::vector<char> is_active(me + mi, 0);
    int q = 0;
    double R_norm = 1.0;
    std::vector<double> Rinv(n, 0.0);

    // unconstrained minimiser x = -Q^-1 c = -J (J' c)
    for(int j = 0; j < n; j++)
    {
      double acc = 0.0;
      if(c)
        for(int i = 0; i < n; i++) acc = fmad(J[i * n + j], c[i], acc);
      tmp[j] = acc;
    }
    for(int i = 0; i < n; i++)
    {
      double acc = 0.0;
      for(int j = 0; j < n; j++) acc = fmad(J[i * n + j], tmp[j], acc);
      x[i] = -acc;
    }

    auto slack = [&](int id) {
      // n_id . x + offset  (>= 0 when satisfied; equalities: = 0)
      double acc = 0.0;
      for(int j = 0; j < n; j++) acc = fmad(normal(id, j), x[j], acc);
      return id < me ? acc - b[id] : acc + dvec[id - me];
    };
    auto compute_dzr = [&]() {
      // four interleaved fma chains per product (dot4, num.hpp): the engine's threads run one product each and
      // a single chain of n dependent fma would be pure latency
      for(int j = 0; j < n; j++) d[j] = dot4(J.data() + j, n, np.data(), 1, n);
      for(int i = 0; i < n; i++) z[i] = dot4(J.data() + i * n + q, 1, d.data() + q, 1, n - q);
      for(int i = q - 1; i >= 0; i--)
      {
        double acc = d[i];
        for(int j = q - 1; j > i; j--) acc = fmad(-R[i * n + j], r[j], acc);
        // reciprocal of the diagonal kept by add / delete: no division in the sweep
        r[i] = kTextbook ? acc / R[i * n + i] : acc * Rinv[i];
      }
    };
    // Givens sweep that folds d[q..n-1] into d[q] and rotates the columns of J (Goldfarb-Idnani's
    // "add constraint").  The textbook sweep runs j = n-1 .. q+1 with h_j = hypot(d[j-1], running value):
    // a chain of n-q-1 dependent hypot / division steps.  |running value at j| is sqrt(sum_{i>=j} d_i^2), so the
    // rotation parameters of ALL steps follow from the suffix sums of squares T_j and can be evaluated
    // independently of each other; that is the form used here (and by the engine, where one thread computes
    // one step's parameters).  T_j is summed by a Kogge-Stone scan (fixed association, zeros are exact), the
    // sign of the running value at j is the sign of d[j] (the sweep normalises cc >= 0).
    auto add_constraint = [&]() -> bool {
      if(kTextbook)
      {
        // the sweep as Goldfarb-Idnani implementations write it (QuadProg++ add_constraint): one dependent
        // hypot / division step per column
        for(int j = n - 1; j >= q + 1; j--)
        {
          double cc = d[j - 1], ss = d[j];
          const double h = givens_hypot(cc, ss);
          if(std::fabs(h) < eps) continue;
          d[j] = 0.0;
          ss = ss / h;
          cc = cc / h;
          if(cc < 0.0)
          {
            cc = -cc;
            ss = -ss;
            d[j - 1] = -h;
          }
          else
            d[j - 1] = h;
          const double xny = ss / (1.0 + cc);
          for(int k = 0; k < n; k++)
          {
            const double t1 = J[k * n + j - 1], t2 = J[k * n + j];
            J[k * n + j - 1] = t1 * cc + t2 * ss;
            J[k * n + j] = xny * (t1 + J[k * n + j - 1]) - t2;
          }
        }
        q++;
        for(int i = 0; i < q; i++) R[i * n + q - 1] = d[i];
        Rinv[q - 1] = 1.0 / d[q - 1];
        if(std::fabs(d[q - 1]) <= eps * R_norm) return false;
        R_norm = std::max(R_norm, std::fabs(d[q - 1]));
        return true;
      }
      std::vector<double> Ta(n + 1, 0.0), Tb(n + 1, 0.0), gc(n, -1.0), gs(n, 0.0), gx(n, 0.0);
      for(int j = q; j < n; j++) Ta[j] = d[j] * d[j];
      for(int off = 1; off < n - q; off <<= 1)
      {
        for(int j = 0; j < n; j++) Tb[j] = Ta[j] + (j + off < n ? Ta[j + off] : 0.0);
        Ta.swap(Tb);
      }
      for(int j = q + 1; j <= n - 1; j++)
      {
        const double aa = d[j - 1];
        const double bmag = j == n - 1 ? std::fabs(d[n - 1]) : std::sqrt(Ta[j]);
        const bool bneg = d[j] < 0.0;
        const double h = std::sqrt(Ta[j - 1]);
        if(h < eps) continue; // gc[j] = -1: no rotation at this step
        const double cc = std::fabs(aa) / h;
        double ss = bmag / h;
        if((aa < 0.0) != bneg) ss = -ss;
        gc[j] = cc;
        gs[j] = ss;
        gx[j] = ss / (1.0 + cc);
      }
      const double dq = n - 1 == q ? d[q] : (d[q] < 0.0 ? -std::sqrt(Ta[q]) : std::sqrt(Ta[q]));
      for(int k = 0; k < n; k++)
        for(int j = n - 1; j >= q + 1; j--)
        {
          const double cc = gc[j];
          if(cc < 0.0) continue;
          const double t1 = J[k * n + j - 1], t2 = J[k * n + j];
          const double a = fmad(t2, gs[j], t1 * cc);
          J[k * n + j - 1] = a;
          J[k * n + j] = fmad(gx[j], t1 + a, -t2);
        }
      d[q] = dq;
      q++;
      for(int i = 0; i < q; i++) R[i * n + q - 1] = d[i];
      Rinv[q - 1] = 1.0 / d[q - 1];
      if(std::fabs(d[q - 1]) <= eps * R_norm) return false; // linearly dependent
      R_norm = std::max(R_norm, std::fabs(d[q - 1]));
      return true;
    };
    auto delete_constraint = [&](int l) {
      int qq = -1;
      for(int i = me; i < q; i++)
        if(A[i] == l)
        {
          qq = i;
          break;
        }
      if(qq < 0) return;
      for(int i = qq; i < q - 1; i++)
      {
        A[i] = A[i + 1];
        u[i] = u[i + 1];
        for(int j = 0; j < n; j++) R[j * n + i] = R[j * n + i + 1];
      }
      A[q - 1] = A[q];
      u[q - 1] = u[q];
      A[q] = -1;
      u[q] = 0.0;
      for(int j = 0; j < q; j++) R[j * n + q - 1] = 0.0;
      q--;
      if(q == 0) return;
      for(int j = qq; j < q; j++)
      {
        double cc = R[j * n + j], ss = R[(j + 1) * n + j];
        const double h = givens_hypot(cc, ss);
        if(std::fabs(h) < eps) continue;
        cc = cc / h;
        ss = ss / h;
        R[(j + 1) * n + j] = 0.0;
        if(cc < 0.0)
        {
          R[j * n + j] = -h;
          cc = -cc;
          ss = -ss;
        }
        else
          R[j * n + j] = h;
        const double xny = ss / (1.0 + cc);
        for(int k = j + 1; k < q; k++)
        {
          const double t1 = R[j * n + k], t2 = R[(j + 1) * n + k];
          const double a = fmad(t2, ss, t1 * cc);
          R[j * n + k] = a;
          R[(j + 1) * n + k] = fmad(xny, t1 + a, -t2);
        }
        for(int k = 0; k < n; k++)
        {
          const double t1 = J[k * n + j], t2 = J[k * n + j + 1];
          const double a = fmad(t2, ss, t1 * cc);
          J[k * n + j] = a;
          J[k * n + j + 1] = fmad(xny, t1 + a, -t2);
        }
      }
      for(int j = qq; j < q; j++) Rinv[j] = 1.0 / R[j * n + j];
    };
    auto dot_tree = [&](const std::vector<double> & a, const std::vector<double> & bb) {
      double p[256];
      for(int i = 0; i < n; i++) p[i] = a[i] * bb[i];
      return tree_sum_leaves(p, n, qp_tree_leaves(n));
    };

    // equality constraints: always active, full step each
    for(int e = 0; e < me; e++)
    {
      for(int j = 0; j < n; j++) np[j] = normal(e, j);
      compute_dzr();
      const double zz = dot_tree(z, z);
      double t2 = 0.0;
      if(std::fabs(zz) > eps) t2 = (-slack(e)) / dot_tree(z, np);
      for(int i = 0; i < n; i++) x[i] = fmad(t2, z[i], x[i]);
      u[q] = t2;
      for(int k = 0; k < q; k++) u[k] = fmad(-t2, r[k], u[k]);
      A[q] = e;
      is_active[e] = 1;
      if(!add_constraint())
      {
        res.status = 1;
        res.x = x;
        return res;
      }
    }

    int iter = 0;
    int ip = -1;
    bool need_pick = true;
    double s_ip = 0.0;
    for(;;)
    {
      if(need_pick)
      {
        // most violated inactive inequality, lowest index on ties
        ip = -1;
        double worst = -viol_tol;
        for(int i = 0; i < mi; i++)
        {
          if(is_active[me + i]) continue;
          const double s = slack(me + i);
          if(s < worst)
          {
            worst = s;
            ip = me + i;
          }
        }
        if(ip < 0) break; // optimal
        // (q == n with a violated constraint left: z = J[:, q:] d[q:] is the empty sum, so the step below is the dual-only
        // step that drops a blocking constraint, as in Goldfarb-Idnani; A and u have n + 1 entries for the pending one)
        iter++;
        if(iter > max_iter)
        {
          res.status = 2;
          break;
        }
        s_ip = worst;
        for(int j = 0; j < n; j++) np[j] = normal(ip, j);
        u[q] = 0.0;
        A[q] = ip;
      }
      // step direction in primal (z) and dual (r) space
      compute_dzr();
      int l = -1;
      double t1 = std::numeric_limits<double>::infinity();
      for(int k = me; k < q; k++)
        if(r[k] > 0.0)
        {
          const double ratio = u[k] / r[k];
          if(ratio < t1)
          {
            t1 = ratio;
            l = A[k];
          }
        }
      const double zz = dot_tree(z, z);
      double t2 = std::numeric_limits<double>::infinity();
      double znp = 0.0;
      if(std::fabs(zz) > eps)
      {
        znp = dot_tree(z, np);
        t2 = (-s_ip) / znp;
      }
      const double t = t1 < t2 ? t1 : t2;
      if(!(t < std::numeric_limits<double>::infinity()))
      {
        res.status = 1; // infeasible
        break;
      }
      if(!(t2 < std::numeric_limits<double>::infinity()))
      {
        // step in dual space only, drop constraint l
        for(int k = 0; k < q; k++) u[k] = fmad(-t, r[k], u[k]);
        u[q] = u[q] + t;
        is_active[l] = 0;
        delete_constraint(l);
        need_pick = false;
        continue;
      }
      for(int i = 0; i < n; i++) x[i] = fmad(t, z[i], x[i]);
      for(int k = 0; k < q; k++) u[k] = fmad(-t, r[k], u[k]);
      u[q] = u[q] + t;
      if(t == t2)
      {
        // full step: the constraint becomes active
        if(!add_constraint())
        {
          res.status = 1; // dependent constraints: treated as infeasible (does not occur on this path)
          break;
        }
        is_active[ip] = 1;
        need_pick = true;
      }
      else
      {
        // partial step: drop the blocking constraint and try again towards ip
        is_active[l] = 0;
        delete_constraint(l);
        s_ip = slack(ip);
        need_pick = false;
      }
    }
    res.x = x;
    res.iters = iter;
    res.active.assign(A.begin(), A.begin() + q);
    return res;
  }
};
} // namespace oracle
