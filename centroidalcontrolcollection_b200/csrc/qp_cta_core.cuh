// qp_cta_core.cuh — one CTA (NT = 128 or 256 threads) solves one strictly convex dense QP with the
// Goldfarb-Idnani dual active-set method; Q, A, C are shared by the batch.
//
//   min 0.5 x'Qx + c'x   s.t.  A x = b,  C x <= d        n <= NT, n_eq + n_ineq <= 4 NT
//
// Replaces: QpSolverCollection::QpSolver::solve(QpCoeff &) as called at reference
// src/LinearMpcZmp.cpp:69, src/IntrinsicallyStableMpc.cpp:93 and src/LinearMpcXY.cpp:181.  Algorithm and
// evaluation order: oracle/qp.hpp (bit-exact): thread i owns variable i / row i of J = L^-T Q_givens,
// thread j column sums, sequential fma chains in the oracle's order; scalar products through a fixed
// NT-leaf tree; the small triangular solve R r = d runs as a column sweep on warp 0.
//
// J and R (n x ld doubles each) live in shared memory when they fit (kGlobal = false; ld odd so that
// row- and column-wise accesses are conflict free; 176 KB at n = 100, one CTA per SM), else in a
// per-CTA slab of global memory that stays L2 resident (kGlobal = true; LinearMpcXY, n = 240).
// kPackedR: R only ever holds a q x q upper triangle (q = active constraints), so it is stored packed by columns with
// room for P.rcap columns — at n = 100 J (80.8 KB) + packed R (62 columns, 15.6 KB) + vectors fit TWICE on an SM, and
// two CTAs per SM hide each other's barrier and dependency stalls.  A problem whose active set outgrows the cap is
// recorded in P.ovf_list and solved again from scratch by the full-R kernel (same arithmetic, so the same bits).
#pragma once
#include "warp_ctx.cuh"

namespace ccc
{
constexpr int kQpThreads = 256; // setup kernel; the solve kernel runs NT = 128 or 256 threads

struct QpParams
{
  int n, me, mi, B, ld;
  const double * J0; // [n][n]  L^-T of Q (setup kernel)
  const double * J0s; // [n][ld] the same with the solve kernel's row stride: one TMA bulk copy per problem (or null)
  const double * At; // [n][me] transposed equality matrix
  const double * Ct; // [n][mi] transposed inequality matrix
  const double * c;  // [B][n] or null
  const double * b;  // [B][me]
  const double * d;  // [B][mi]
  const int * setup_ok; // [0]: Q positive definite; [1]: inequality rows i and i + mi/2 are exact negatives of each other;
                        // [2]: C is exactly [-I; I] (the box x_min <= x <= x_max of a QpCoeff, src/LinearMpcXY.cpp:177-178)
  // matrix groups (LinearMpcXY over a sweep of schedules: one Q / A per schedule): problem b uses group grp[b] and the
  // per-group arrays are gs_* doubles apart (0: shared by all groups); grp == nullptr: one group
  const int * grp = nullptr;
  size_t gs_J0 = 0, gs_J0s = 0, gs_At = 0, gs_Ct = 0;
  int gs_ok = 0;
  int max_iter;
  double viol_tol;
  int rcap = 0;              // kPackedR: columns the packed R can hold
  int * ovf_count = nullptr; // kPackedR: number of problems whose active set outgrew rcap ...
  int * ovf_list = nullptr;  // ... and their ids; the full-R kernel solves them again
  const int * list = nullptr;       // solve problems list[0 .. *list_count) instead of 0 .. B-1 (the fallback pass)
  const int * list_count = nullptr;
  double * out_x;
  int * out_iters;
  int * out_status;
  int * out_n_active;
  int * out_active; // [B][n]
};

CCC_DEV double givens_hypot(double a, double b)
{
  const double a1 = dabs(a), b1 = dabs(b);
  if(a1 > b1)
  {
    const double t = a1 > 0 ? b1 / a1 : 0.0;
    return a1 * dsqrt(dfma(t, t, 1.0));
  }
  if(b1 > a1)
  {
    const double t = a1 / b1;
    return b1 * dsqrt(dfma(t, t, 1.0));
  }
  return a1 * dsqrt(2.0);
}

/** Shared-memory layout of one QP (offsets in doubles). */
template<int NT, bool kGlobal, bool kPackedR = false>
struct QpSm
{
  int n, ld, rcap;
  CCC_DEV QpSm(int n_, int ld_, int rcap_ = 0) : n(n_), ld(ld_), rcap(rcap_) {}
  CCC_HD static int packed(int rcap) { return (rcap * (rcap + 1) / 2 + 1) & ~1; } // doubles of a packed R, even
  CCC_DEV int mats() const { return kGlobal ? 0 : n * ld + (kPackedR ? packed(rcap) : n * ld); }
  CCC_DEV int J() const { return 0; }
  CCC_DEV int R() const { return n * ld; }
  CCC_DEV int vec(int k) const { return mats() + k * NT; } // x, z, d, np, r, u(+1 in next), tmp, gs, gx, rinv
  CCC_DEV int red() const { return vec(11); }               // 2 NT doubles: two trees / argmin values / scan ping-pong
  CCC_DEV int ints() const { return vec(11) + 2 * NT; }     // int area (2 NT doubles): A[NT+4], red_i[NT], is_active[4 NT bytes], ctrl
  static size_t bytes(int n, int ld, int rcap = 0)
  {
    return (size_t)((kGlobal ? 0 : n * ld + (kPackedR ? packed(rcap) : n * ld)) + 11 * NT + 2 * NT + 2 * NT) * sizeof(double);
  }
};

struct QpCtrl
{
  double t, s_ip;
  int action; // 0: done/stop, 1: dual step only, 2: full step, 3: partial step
  int l, ip, status;
};

template<int NT, bool kGlobal, bool kPackedR = false>
struct QpCta
{
  const QpParams & P;
  double * sm;
  int b, tid, n, me, mi, ld;
  const double *gJ0, *gJ0s, *gAt, *gCt; // this problem's group of batch-shared matrices
  const int * gok;
  bool box = false; // C = [-I; I] (setup flag 2, read in solve())
  double *J, *R, *x, *z, *d, *np, *r, *u, *tmp, *gs, *gx, *rinv, *red;
  int *A, *red_i;
  unsigned char * is_active;
  QpCtrl * ctrl;
  int q;
  double R_norm;
  unsigned long long * mbar = nullptr; // mbarrier of the TMA load of J (kernel-owned, initialised once per CTA)
  unsigned mbar_phase = 0;

  /** `gmat`: this CTA's slab of 2 n ld doubles in global memory (kGlobal only). */
  CCC_DEV QpCta(const QpParams & p, double * smem, int prob, double * gmat = nullptr)
  : P(p), sm(smem), b(prob), tid(thread_id()), n(p.n), me(p.me), mi(p.mi), ld(p.ld), q(0), R_norm(1.0)
  {
    static_assert(!(kGlobal && kPackedR), "the packed R exists to fit two CTAs' J into shared memory");
    const size_t g = p.grp ? static_cast<size_t>(ldg(p.grp + prob)) : 0;
    gJ0 = p.J0 + g * p.gs_J0;
    gJ0s = p.J0s ? p.J0s + g * p.gs_J0s : nullptr;
    gAt = p.At + g * p.gs_At;
    gCt = p.Ct + g * p.gs_Ct;
    gok = p.setup_ok + g * p.gs_ok;
    QpSm<NT, kGlobal, kPackedR> L(n, ld, p.rcap);
    J = (kGlobal ? gmat : sm) + L.J();
    R = (kGlobal ? gmat : sm) + L.R();
    x = sm + L.vec(0);
    z = sm + L.vec(1);
    d = sm + L.vec(2);
    np = sm + L.vec(3);
    r = sm + L.vec(4);
    u = sm + L.vec(5); // NT + 1 entries: spills one double into vec(6)'s first slot
    tmp = sm + L.vec(7);
    gs = sm + L.vec(8);
    gx = sm + L.vec(9);
    rinv = sm + L.vec(10); // 1 / R[i][i], kept up to date by add_constraint / delete_constraint
    red = sm + L.red();
    int * ib = reinterpret_cast<int *>(sm + L.ints());
    A = ib;                                                          // NT + 1 ints
    red_i = ib + NT + 4;                                             // NT ints
    is_active = reinterpret_cast<unsigned char *>(ib + 2 * NT + 4); // 4 NT bytes
    ctrl = reinterpret_cast<QpCtrl *>(ib + 3 * NT + 4);
  }

  /** R(i, j), i <= j < q (plus the sub-diagonal R(j + 1, j) in the full layout while a constraint is being deleted). */
  CCC_DEV double & Rat(int i, int j) const { return kPackedR ? R[j * (j + 1) / 2 + i] : R[i * ld + j]; }

  CCC_DEV double normal(int id, int j) const
  {
    return id < me ? ldg(gAt + (size_t)j * me + id) : -ldg(gCt + (size_t)j * mi + (id - me));
  }

  /** n_id . x + offset (>= 0 when satisfied) */
  CCC_DEV double slack(int id) const
  {
    const double acc = slack_dot(id);
    return id < me ? acc - ldg(P.b + (size_t)b * me + id) : acc + ldg(P.d + (size_t)b * mi + (id - me));
  }

  /** n_id . x */
  CCC_DEV double slack_dot(int id) const
  {
    // C = [-I; I]: the chain over an identity row is x_i or -x_i (its zero terms leave the sum unchanged, `+ 0.0` is
    // the chain's -0 -> +0)
    if(box && id >= me) return id - me < n ? x[id - me] + 0.0 : (-x[id - me - n]) + 0.0;
    // the fma chain is sequential (oracle order); the loads (L2-resident constraint matrix) are not: the next
    // block of 16 is in flight while the chain consumes the current one
    constexpr int kBlk = 16;
    double acc = 0.0;
    double v[kBlk], w[kBlk];
    int j = 0;
    if(n >= kBlk)
    {
      CCC_UNROLL
      for(int e = 0; e < kBlk; e++) v[e] = normal(id, e);
      for(; j + 2 * kBlk <= n; j += kBlk)
      {
        CCC_UNROLL
        for(int e = 0; e < kBlk; e++) w[e] = normal(id, j + kBlk + e);
        CCC_UNROLL
        for(int e = 0; e < kBlk; e++) acc = dfma(v[e], x[j + e], acc);
        CCC_UNROLL
        for(int e = 0; e < kBlk; e++) v[e] = w[e];
      }
      CCC_UNROLL
      for(int e = 0; e < kBlk; e++) acc = dfma(v[e], x[j + e], acc);
      j += kBlk;
    }
    for(; j < n; j++) acc = dfma(normal(id, j), x[j], acc);
    return acc;
  }

  /** two NT-leaf pairwise trees at once over (z.z, z.np); results in red[0], red[NT].  Same tree as
   *  oracle/qp.hpp tree_sum_leaves (strides NT/2 .. 1); the strides that cross warps go through shared memory,
   *  the last five run as shuffles on warp 0 (lane i < off adds the value of lane i + off: the same pair). */
  CCC_DEV void dot_trees()
  {
    red[tid] = tid < n ? z[tid] * z[tid] : 0.0;
    red[NT + tid] = tid < n ? z[tid] * np[tid] : 0.0;
    cta_sync();
    for(int off = NT / 2; off >= 32; off >>= 1)
    {
      if(tid < off)
      {
        red[tid] = red[tid] + red[tid + off];
        red[NT + tid] = red[NT + tid] + red[NT + tid + off];
      }
      cta_sync();
    }
    if(tid < 32)
    {
      double a = red[tid], c = red[NT + tid];
      CCC_UNROLL
      for(int off = 16; off >= 1; off >>= 1)
      {
        const double ao = warp_shfl(a, (tid + off) & 31), co = warp_shfl(c, (tid + off) & 31);
        a = a + ao;
        c = c + co;
      }
      if(tid == 0)
      {
        red[0] = a;
        red[NT] = c;
      }
    }
    cta_sync();
  }

  /** CTA-wide argmin of (v, idx) with the lowest index winning ties; idx < 0 = no candidate.  Comparisons only,
   *  so the order of the reduction does not matter: shuffles inside the warps, the NT/32 partial results through
   *  shared memory.  Every thread returns with the result. */
  CCC_DEV void cta_argmin(double & v, int & idx)
  {
    const int lane = tid & 31;
    CCC_UNROLL
    for(int off = 16; off >= 1; off >>= 1)
    {
      const double vo = warp_shfl_xor(v, off);
      const int io = warp_shfl_i(idx, lane ^ off);
      const bool take = io >= 0 && (idx < 0 || vo < v || (vo == v && io < idx));
      v = take ? vo : v;
      idx = take ? io : idx;
    }
    if(lane == 0)
    {
      red[tid >> 5] = v;
      red_i[tid >> 5] = idx;
    }
    cta_sync();
    v = red[0];
    idx = red_i[0];
    CCC_UNROLL
    for(int w = 1; w < NT / 32; w++)
    {
      const double vo = red[w];
      const int io = red_i[w];
      const bool take = io >= 0 && (idx < 0 || vo < v || (vo == v && io < idx));
      v = take ? vo : v;
      idx = take ? io : idx;
    }
    cta_sync(); // red / red_i are free again
  }

  /** d = J' np, z = J[:, q:] d[q:], r = R^-1 d[:q] */
  CCC_DEV void compute_dzr()
  {
    // both products as four interleaved fma chains (oracle num.hpp dot4: slot = position % 4 inside the summed
    // range, combined as (s0 + s1) + (s2 + s3))
    if(tid < n)
    {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int i = 0;
      for(; i + 4 <= n; i += 4)
      {
        s0 = dfma(J[i * ld + tid], np[i], s0);
        s1 = dfma(J[(i + 1) * ld + tid], np[i + 1], s1);
        s2 = dfma(J[(i + 2) * ld + tid], np[i + 2], s2);
        s3 = dfma(J[(i + 3) * ld + tid], np[i + 3], s3);
      }
      if(i < n) s0 = dfma(J[i * ld + tid], np[i], s0);
      if(i + 1 < n) s1 = dfma(J[(i + 1) * ld + tid], np[i + 1], s1);
      if(i + 2 < n) s2 = dfma(J[(i + 2) * ld + tid], np[i + 2], s2);
      d[tid] = (s0 + s1) + (s2 + s3);
    }
    cta_sync();
    if(tid < n)
    {
      const double * row = J + tid * ld;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int j = q;
      for(; j + 4 <= n; j += 4)
      {
        s0 = dfma(row[j], d[j], s0);
        s1 = dfma(row[j + 1], d[j + 1], s1);
        s2 = dfma(row[j + 2], d[j + 2], s2);
        s3 = dfma(row[j + 3], d[j + 3], s3);
      }
      if(j < n) s0 = dfma(row[j], d[j], s0);
      if(j + 1 < n) s1 = dfma(row[j + 1], d[j + 1], s1);
      if(j + 2 < n) s2 = dfma(row[j + 2], d[j + 2], s2);
      z[tid] = (s0 + s1) + (s2 + s3);
    }
    if(tid < 32)
    {
      // column sweep of the back substitution on warp 0: lane owns rows lane, lane + 32, ...
      constexpr int kSlots = NT / 32;
      double a[kSlots];
      CCC_UNROLL
      for(int s = 0; s < kSlots; s++) a[s] = tid + 32 * s < q ? d[tid + 32 * s] : 0.0;
      for(int i = q - 1; i >= 0; i--)
      {
        if((i & 31) == tid)
        {
          const int slot = i >> 5;
          double ai = a[0];
          CCC_UNROLL
          for(int s = 1; s < kSlots; s++) ai = slot == s ? a[s] : ai;
          r[i] = ai * rinv[i]; // no division on the sweep's dependency chain
        }
        warp_sync();
        const double ri = r[i];
        CCC_UNROLL
        for(int s = 0; s < kSlots; s++)
          if(tid + 32 * s < i) a[s] = dfma(-Rat(tid + 32 * s, i), ri, a[s]);
      }
    }
    cta_sync();
  }

  /** Givens sweep that folds d[q..n-1] into d[q] and rotates J; appends the column to R.  The rotation
   *  parameters of all n-q-1 steps come from the suffix sums of squares of d (Kogge-Stone scan, then thread j
   *  evaluates step j: two square roots and three divisions in parallel instead of a chain of n-q-1 dependent
   *  hypot / division steps), after which every thread rotates its own row of J with two fma per step.
   *  Operation order: oracle/qp.hpp add_constraint.
   *  Returns false if the new constraint is linearly dependent on the active ones. */
  CCC_DEV bool add_constraint()
  {
    const double eps = 2.220446049250313e-16;
    double * Ta = red;
    double * Tb = red + NT;
    Ta[tid] = (tid >= q && tid < n) ? d[tid] * d[tid] : 0.0;
    cta_sync();
    for(int off = 1; off < n - q; off <<= 1)
    {
      Tb[tid] = Ta[tid] + (tid + off < n ? Ta[tid + off] : 0.0);
      cta_sync();
      double * t = Ta;
      Ta = Tb;
      Tb = t;
    }
    double * gc = tmp;
    if(tid >= q + 1 && tid <= n - 1)
    {
      const int j = tid;
      const double aa = d[j - 1];
      const double bmag = j == n - 1 ? dabs(d[n - 1]) : dsqrt(Ta[j]);
      const bool bneg = d[j] < 0.0;
      const double h = dsqrt(Ta[j - 1]);
      double cc = -1.0, ss = 0.0, xny = 0.0; // cc = -1: no rotation at this step
      if(!(h < eps))
      {
        cc = dabs(aa) / h;
        ss = bmag / h;
        if((aa < 0.0) != bneg) ss = -ss;
        xny = ss / (1.0 + cc);
      }
      gc[j] = cc;
      gs[j] = ss;
      gx[j] = xny;
    }
    const double dq = n - 1 == q ? d[q] : (d[q] < 0.0 ? -dsqrt(Ta[q]) : dsqrt(Ta[q]));
    cta_sync();
    if(tid < n)
    {
      // the chain runs through `a` only (two fma per step); parameters and the row entry of the next step are
      // loaded one step ahead so that no shared-memory latency sits on it
      double * row = J + tid * ld;
      double t2 = row[n - 1];
      int j = n - 1;
      double cc = 0.0, ss = 0.0, xny = 0.0, t1 = 0.0;
      if(j >= q + 1)
      {
        cc = gc[j];
        ss = gs[j];
        xny = gx[j];
        t1 = row[j - 1];
      }
      for(; j >= q + 1; j--)
      {
        double ccn = 0.0, ssn = 0.0, xnyn = 0.0, t1n = 0.0;
        if(j - 1 >= q + 1)
        {
          ccn = gc[j - 1];
          ssn = gs[j - 1];
          xnyn = gx[j - 1];
          t1n = row[j - 2];
        }
        if(cc < 0.0)
        {
          t2 = t1; // no rotation at this step
        }
        else
        {
          const double a = dfma(t2, ss, t1 * cc);
          row[j] = dfma(xny, t1 + a, -t2);
          row[j - 1] = a;
          t2 = a;
        }
        cc = ccn;
        ss = ssn;
        xny = xnyn;
        t1 = t1n;
      }
    }
    cta_sync();
    if(tid < q) Rat(tid, q) = d[tid];
    if(tid == 0)
    {
      Rat(q, q) = dq;
      rinv[q] = 1.0 / dq;
    }
    q++;
    const bool ok = !(dabs(dq) <= eps * R_norm);
    if(ok) R_norm = R_norm < dabs(dq) ? dabs(dq) : R_norm;
    cta_sync();
    return ok;
  }

  CCC_DEV void delete_constraint(int l)
  {
    const double eps = 2.220446049250313e-16;
    int qq = -1;
    for(int i = me; i < q; i++)
      if(A[i] == l)
      {
        qq = i;
        break;
      }
    cta_sync();
    if(qq < 0) return;
    // drop column qq: the columns to its right move one place left, which leaves one sub-diagonal entry per moved
    // column (the old diagonal).  Full layout: it lands in R(i + 1, i); packed layout: in sub[i] (the gs vector, free
    // outside add_constraint).
    double * sub = gs;
    if(kPackedR)
    {
      if(tid < q)
        for(int i = (qq > tid - 1 ? qq : tid - 1); i < q - 1; i++)
        {
          const double v = Rat(tid, i + 1);
          if(tid <= i)
            Rat(tid, i) = v;
          else
            sub[i] = v;
        }
    }
    else if(tid < n)
    {
      for(int i = qq; i < q - 1; i++) R[tid * ld + i] = R[tid * ld + i + 1];
    }
    if(tid == 0)
    {
      for(int i = qq; i < q - 1; i++)
      {
        A[i] = A[i + 1];
        u[i] = u[i + 1];
      }
      A[q - 1] = A[q];
      u[q - 1] = u[q];
      A[q] = -1;
      u[q] = 0.0;
    }
    cta_sync();
    if(!kPackedR && tid < q) R[tid * ld + q - 1] = 0.0;
    q--;
    cta_sync();
    if(q == 0) return;
    for(int j = qq; j < q; j++)
    {
      double cc = Rat(j, j), ss = kPackedR ? sub[j] : R[(j + 1) * ld + j];
      cta_sync(); // everyone has read the pivot pair before it is rewritten
      const double h = givens_hypot(cc, ss);
      if(dabs(h) < eps) continue;
      cc = cc / h;
      ss = ss / h;
      double rjj;
      if(cc < 0.0)
      {
        rjj = -h;
        cc = -cc;
        ss = -ss;
      }
      else
        rjj = h;
      const double xny = ss / (1.0 + cc);
      if(tid == 0)
      {
        if(kPackedR)
          sub[j] = 0.0;
        else
          R[(j + 1) * ld + j] = 0.0;
        Rat(j, j) = rjj;
      }
      if(tid > j && tid < q)
      {
        const double t1 = Rat(j, tid), t2 = Rat(j + 1, tid);
        const double a = dfma(t2, ss, t1 * cc);
        Rat(j, tid) = a;
        Rat(j + 1, tid) = dfma(xny, t1 + a, -t2);
      }
      if(tid < n)
      {
        const double t1 = J[tid * ld + j], t2 = J[tid * ld + j + 1];
        const double a = dfma(t2, ss, t1 * cc);
        J[tid * ld + j] = a;
        J[tid * ld + j + 1] = dfma(xny, t1 + a, -t2);
      }
      cta_sync();
    }
    if(tid >= qq && tid < q) rinv[tid] = 1.0 / Rat(tid, tid);
    cta_sync();
  }

  CCC_DEV void solve()
  {
    const double inf = 1.0 / 0.0;
    const double eps = 2.220446049250313e-16;
    int status = 0, iter = 0;
    if(ldg(gok) == 0)
    {
      if(tid < n && P.out_x) P.out_x[(size_t)b * n + tid] = 0.0;
      if(tid == 0)
      {
        if(P.out_iters) P.out_iters[b] = 0;
        if(P.out_status) P.out_status[b] = 3;
        if(P.out_n_active) P.out_n_active[b] = 0;
      }
      if(tid < n && P.out_active) P.out_active[(size_t)b * n + tid] = -1;
      return;
    }
    // load the shared factor, reset the bookkeeping
    bool by_tma = false;
#if CCC_HAS_TMA
    // one TMA bulk copy brings the pre-strided factor (n x ld doubles, 80.8 KB at n = 100) into shared memory
    // while the threads reset the bookkeeping; the previous problem's generic accesses to J are ordered before
    // the async-proxy write by the fence (every thread passed the barrier that ends solve())
    by_tma = !kGlobal && mbar != nullptr && gJ0s != nullptr && ((n * ld) & 1) == 0;
    if(by_tma && tid == 0)
    {
      fence_proxy_async_smem();
      const unsigned bytes = static_cast<unsigned>(n * ld * sizeof(double));
      mbar_arrive_expect_tx(mbar, bytes);
      tma_bulk_g2s(J, gJ0s, bytes, mbar);
    }
#endif
    if(!by_tma)
      for(int e = tid; e < n * n; e += NT) J[(e / n) * ld + (e % n)] = ldg(gJ0 + e);
    for(int e = tid; e < me + mi; e += NT) is_active[e] = 0;
    for(int e = tid; e <= n; e += NT)
    {
      A[e] = -1;
      u[e] = 0.0;
    }
#if CCC_HAS_TMA
    if(by_tma)
    {
      mbar_wait(mbar, mbar_phase);
      mbar_phase ^= 1u;
    }
#endif
    cta_sync();
    // unconstrained minimiser x = -J (J' c)
    if(tid < n)
    {
      double acc = 0.0;
      if(P.c)
        for(int i = 0; i < n; i++) acc = dfma(J[i * ld + tid], ldg(P.c + (size_t)b * n + i), acc);
      tmp[tid] = acc;
    }
    cta_sync();
    if(tid < n)
    {
      double acc = 0.0;
      for(int j = 0; j < n; j++) acc = dfma(J[tid * ld + j], tmp[j], acc);
      x[tid] = -acc;
    }
    cta_sync();

    // equality constraints: always active, one full step each
    for(int e = 0; e < me && status == 0; e++)
    {
      if(tid < n) np[tid] = normal(e, tid);
      cta_sync();
      compute_dzr();
      dot_trees();
      if(tid == 0)
      {
        double t2 = 0.0;
        if(dabs(red[0]) > eps) t2 = (-slack(e)) / red[NT];
        ctrl->t = t2;
      }
      cta_sync();
      const double t2 = ctrl->t;
      if(tid < n) x[tid] = dfma(t2, z[tid], x[tid]);
      if(tid < q) u[tid] = dfma(-t2, r[tid], u[tid]);
      if(tid == 0)
      {
        u[q] = t2;
        A[q] = e;
        is_active[e] = 1;
      }
      cta_sync();
      if(!add_constraint()) status = 1;
    }

    const bool paired = ldg(gok + 1) != 0;
    box = ldg(gok + 2) != 0;
    bool need_pick = true;
    int ip = -1;
    double s_ip = 0.0;
    while(status == 0)
    {
      if(need_pick)
      {
        // most violated inactive inequality, lowest index on ties
        double best = -P.viol_tol;
        int best_i = -1;
        if(paired)
        {
          // rows i and i + h are exact negatives: fma(-a, b, -c) = -fma(a, b, c), so one chain serves both
          const int h = mi / 2;
          for(int i = tid; i < h; i += NT)
          {
            const bool a0 = is_active[me + i], a1 = is_active[me + i + h];
            if(a0 && a1) continue;
            const double acc = slack_dot(me + i);
            if(!a0)
            {
              const double s = acc + ldg(P.d + (size_t)b * mi + i);
              if(s < best || (s == best && best_i >= 0 && i < best_i))
              {
                best = s;
                best_i = i;
              }
            }
            if(!a1)
            {
              const double s = (-acc) + ldg(P.d + (size_t)b * mi + i + h);
              if(s < best || (s == best && best_i >= 0 && i + h < best_i))
              {
                best = s;
                best_i = i + h;
              }
            }
          }
        }
        else
          for(int i = tid; i < mi; i += NT)
          {
            if(is_active[me + i]) continue;
            const double s = slack(me + i);
            if(s < best)
            {
              best = s;
              best_i = i;
            }
          }
        cta_argmin(best, best_i);
        const int pick = best_i;
        const double worst = best;
        if(pick < 0) break; // optimal
        // (q == n with a violated constraint left: z is the empty sum, the step below is the dual-only step that drops a
        // blocking constraint; A and u hold n + 1 entries for the pending one — oracle/qp.hpp)
        if(kPackedR && q >= P.rcap)
        {
          status = 5; // the packed R is full: this problem goes to the full-R kernel (never leaves the library)
          break;
        }
        iter++;
        if(iter > P.max_iter)
        {
          status = 2;
          break;
        }
        ip = me + pick;
        s_ip = worst;
        if(tid < n) np[tid] = normal(ip, tid);
        if(tid == 0)
        {
          u[q] = 0.0;
          A[q] = ip;
        }
        cta_sync();
      }
      compute_dzr();
      dot_trees();
      // step length: t1 = min over the active inequalities with r_k > 0 of u_k / r_k (first minimum in active-set
      // order, one division per thread instead of a serial loop on one thread), t2 = -s_ip / (z . np)
      const double zz = red[0], znp = red[NT];
      cta_sync(); // everyone has read the two sums before cta_argmin reuses the buffer
      double t1 = inf;
      int kmin = -1;
      if(tid >= me && tid < q && r[tid] > 0.0)
      {
        t1 = u[tid] / r[tid];
        kmin = tid;
      }
      cta_argmin(t1, kmin);
      const int l = kmin >= 0 ? A[kmin] : -1;
      if(kmin < 0) t1 = inf;
      double t2 = inf;
      if(dabs(zz) > eps) t2 = (-s_ip) / znp;
      const double t = t1 < t2 ? t1 : t2;
      int action; // 0: infeasible, 1: dual step only, 2: full step, 3: partial step
      if(!(t < inf))
        action = 0;
      else if(!(t2 < inf))
        action = 1;
      else if(t == t2)
        action = 2;
      else
        action = 3;
      if(action == 0)
      {
        status = 1;
        break;
      }
      if(action == 1)
      {
        if(tid < q) u[tid] = dfma(-t, r[tid], u[tid]);
        if(tid == 0)
        {
          u[q] = u[q] + t;
          is_active[l] = 0;
        }
        cta_sync();
        delete_constraint(l);
        need_pick = false;
        continue;
      }
      if(tid < n) x[tid] = dfma(t, z[tid], x[tid]);
      if(tid < q) u[tid] = dfma(-t, r[tid], u[tid]);
      if(tid == 0) u[q] = u[q] + t;
      cta_sync();
      if(action == 2)
      {
        if(!add_constraint())
        {
          status = 1;
          break;
        }
        if(tid == 0) is_active[ip] = 1;
        cta_sync();
        need_pick = true;
      }
      else
      {
        if(tid == 0) is_active[l] = 0;
        cta_sync();
        delete_constraint(l);
        if(tid == 0) ctrl->s_ip = slack(ip);
        cta_sync();
        s_ip = ctrl->s_ip;
        need_pick = false;
      }
    }
    cta_sync();
    if(tid < n)
    {
      if(P.out_x) P.out_x[(size_t)b * n + tid] = x[tid];
      if(P.out_active) P.out_active[(size_t)b * n + tid] = tid < q ? A[tid] : -1;
    }
    if(tid == 0)
    {
      if(P.out_iters) P.out_iters[b] = iter;
      if(P.out_status) P.out_status[b] = status;
      if(P.out_n_active) P.out_n_active[b] = q;
#ifndef CCC_WARP_EMU
      if(kPackedR && status == 5) P.ovf_list[atomicAdd(P.ovf_count, 1)] = b;
#else
      if(kPackedR && status == 5) P.ovf_list[(*P.ovf_count)++] = b;
#endif
    }
    cta_sync();
  }
};
} // namespace ccc

namespace ccc
{
/** Batch-invariant setup, one CTA: L = chol(Q) (scratch Lg), J0 = L^-T, transposed A and C.
 *  Evaluation order: oracle/qp.hpp DenseQpShared::setup — every entry is the oracle's sequential fma chain over
 *  j ascending.  The oracle writes the chains entry by entry (left looking); here they advance together, one term per
 *  step j (right looking): once column j of L (of L^-1) is final, every later entry takes its j-th term.  The terms
 *  of one entry still arrive in ascending j, so the bits are the same, but a step is n^2 / 2 independent fma spread
 *  over the CTA with coalesced accesses (thread = column) instead of n-long dependent chains per thread
 *  (n = 240: 7.0 ms -> well under 1 ms, profiles/r02_summary.md).
 *  Lg holds the trailing matrix on and below the diagonal and L' above it (column j of L contiguous). */
CCC_DEV void qp_setup_cta(int n, int me, int mi, const double * Q, const double * A, const double * C, double * Lg,
                          double * invd, double * J0, double * At, double * Ct, int * ok_flag, double * J0s, bool write_ct,
                          double * bcast /* n doubles of shared memory: the column every thread reads in a step */)
{
  const int tid = thread_id();
  if(tid == 0) *ok_flag = 1;
  // trailing matrix = lower triangle of Q; J0 = identity (the right-hand sides e_i, row i)
  for(int i = 0; i < n; i++)
    for(int k = tid; k < n; k += kQpThreads)
    {
      if(k <= i) Lg[(size_t)i * n + k] = Q[(size_t)i * n + k];
      J0[(size_t)i * n + k] = i == k ? 1.0 : 0.0;
    }
  cta_sync();
  constexpr int kU = 32; // entries in flight per thread: the loads of a block are issued before its fma / stores
  for(int j = 0; j < n; j++)
  {
    // column j is final: pivot (every thread evaluates it for itself), L(i, j) = a(i, j) / d, kept as row j of L'
    // (above the diagonal of Lg, for the second phase) and in shared memory (for this step)
    double acc = Lg[(size_t)j * n + j];
    const bool bad = !(acc > 0.0);
    if(bad) acc = 1.0;
    const double dd = dsqrt(acc);
    const double inv = 1.0 / dd;
    if(tid == 0)
    {
      if(bad) *ok_flag = 0;
      invd[j] = inv;
    }
    for(int i = j + 1 + tid; i < n; i += kQpThreads)
    {
      const double l = Lg[(size_t)i * n + j] * inv;
      Lg[(size_t)j * n + i] = l;
      bcast[i] = l;
    }
    cta_sync();
    // j-th term of every later entry a(i, k), j < k <= i: thread k owns column k
    for(int k = j + 1 + tid; k < n; k += kQpThreads)
    {
      const double lk = bcast[k];
      int i = k;
      for(; i + kU <= n; i += kU)
      {
        double v[kU];
        CCC_UNROLL
        for(int e = 0; e < kU; e++) v[e] = Lg[(size_t)(i + e) * n + k];
        CCC_UNROLL
        for(int e = 0; e < kU; e++) v[e] = dfma(-bcast[i + e], lk, v[e]);
        CCC_UNROLL
        for(int e = 0; e < kU; e++) Lg[(size_t)(i + e) * n + k] = v[e];
      }
      for(; i < n; i++) Lg[(size_t)i * n + k] = dfma(-bcast[i], lk, Lg[(size_t)i * n + k]);
    }
    cta_sync();
  }
  // J0 row i = L^-1 e_i: z(i, r) = (delta_ir - sum_{j < r} L(r, j) z(i, j)) / d_r; entries r < i are exact zeros and
  // their terms leave the later chains unchanged, so only rows i <= j take the j-th term
  for(int j = 0; j < n; j++)
  {
    const double inv = invd[j];
    for(int i = tid; i <= j; i += kQpThreads)
    {
      const double z = J0[(size_t)i * n + j] * inv;
      J0[(size_t)i * n + j] = z;
      bcast[i] = z;
    }
    cta_sync();
    const double * Lt = Lg + (size_t)j * n;
    for(int r = j + 1 + tid; r < n; r += kQpThreads)
    {
      const double lrj = Lt[r];
      int i = 0;
      for(; i + kU <= j + 1; i += kU)
      {
        double v[kU];
        CCC_UNROLL
        for(int e = 0; e < kU; e++) v[e] = J0[(size_t)(i + e) * n + r];
        CCC_UNROLL
        for(int e = 0; e < kU; e++) v[e] = dfma(-lrj, bcast[i + e], v[e]);
        CCC_UNROLL
        for(int e = 0; e < kU; e++) J0[(size_t)(i + e) * n + r] = v[e];
      }
      for(; i <= j; i++) J0[(size_t)i * n + r] = dfma(-lrj, bcast[i], J0[(size_t)i * n + r]);
    }
    cta_sync();
  }
  if(J0s)
  {
    const int ld = n | 1;
    for(int e = tid; e < n * ld; e += kQpThreads) J0s[e] = (e % ld) < n ? J0[(e / ld) * n + (e % ld)] : 0.0;
  }
  for(int e = tid; e < n * me; e += kQpThreads) At[e] = A[(e % me) * n + e / me];
  if(write_ct)
    for(int e = tid; e < n * mi; e += kQpThreads) Ct[e] = C[(size_t)(e % mi) * n + e / mi];
  // two-sided constraints lo <= G x <= hi arrive as C = [-G; G] (reference src/LinearMpcZmp.cpp:25,
  // src/IntrinsicallyStableMpc.cpp:42; the box rows of LinearMpcXY): if row i + mi/2 is the exact negative of
  // row i for every i, the solve kernel evaluates one product per pair (the other one is its exact negative)
  if(tid == 0) ok_flag[1] = (mi % 2 == 0) ? 1 : 0;
  cta_sync();
  if(mi % 2 == 0)
  {
    const int h = mi / 2;
    bool same = true;
    for(int e = tid; e < n * h; e += kQpThreads)
    {
      const int i = e % h, j = e / h;
      same = same && (C[(size_t)(i + h) * n + j] == -C[(size_t)i * n + j]);
    }
    if(!same) ok_flag[1] = 0;
  }
  cta_sync();
  // C == [-I; I] exactly (the bounds of a QpCoeff as inequality rows): the solve kernel reads x_i instead of
  // streaming the identity
  if(tid == 0) ok_flag[2] = (mi == 2 * n) ? 1 : 0;
  cta_sync();
  if(mi == 2 * n)
  {
    bool ident = true;
    for(int e = tid; e < n * n; e += kQpThreads)
    {
      const int i = e / n, j = e % n;
      ident = ident && (C[(size_t)i * n + j] == (i == j ? -1.0 : 0.0)) && (C[(size_t)(n + i) * n + j] == (i == j ? 1.0 : 0.0));
    }
    if(!ident) ok_flag[2] = 0;
  }
  cta_sync();
}
} // namespace ccc
