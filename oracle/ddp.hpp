// oracle/ddp.hpp — CPU restatement of nmpc_ddp::DDPSolver<StateDim, InputDim>.
//
// TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED beyond the reference tests' closed-loop
// tolerances: nmpc_ddp (isri-aist/NMPC; un-versioned find_package at reference
// CMakeLists.txt:25, CI builds its default branch, .github/workflows/ci-colcon.yaml:80-85)
// is not under /root/reference.  This file restates the published algorithm it ports
// (Tassa et al., iLQG.m / "Control-limited DDP", ICRA 2014; regularisation schedule of
// Tassa, Erez, Todorov IROS 2012) and anchors on the reference's call sites:
//   solve(t0, x0, u_list)            src/DdpCentroidal.cpp:229,233  src/DdpZmp.cpp:160-171
//   config() fields                  src/DdpCentroidal.cpp:197-201  tests/src/TestDdpCentroidal.cpp:116
//   setInputLimitsFunc               src/DdpCentroidal.cpp:202-210
//   controlData().u_list             src/DdpCentroidal.cpp:236
//   traceDataList().back().iter      tests/src/TestDdpCentroidal.cpp:129
// Only first-order dynamics derivatives are used (the second-order overload throws,
// include/CCC/DdpCentroidal.h:217-228).
//
// Stage data are addressed by stage index k (t_k = t0 + k dt): callbacks are sampled once
// by the caller, exactly as the batched engine does.
#pragma once
#include "boxqp.hpp"

#include <algorithm>
#include <cstring>

namespace oracle
{
struct DdpConfig
{
  bool with_input_constraint = false;
  int max_iter = 500;
  int reg_type = 1;
  double initial_lambda = 1e-4;
  double initial_dlambda = 1.0;
  double lambda_factor = 1.6;
  double lambda_min = 1e-6;
  double lambda_max = 1e10;
  double k_rel_norm_thre = 1e-4;
  double lambda_thre = 1e-5;
  std::vector<double> alpha_list; // 10^linspace(0,-3,11)
  double cost_update_ratio_thre = 0;
  double cost_update_thre = 1e-7;
  BoxQpConfig boxqp;

  DdpConfig()
  {
    for(int i = 0; i < 11; i++) alpha_list.push_back(std::pow(10.0, -3.0 * i / 10.0));
  }
};

/** Problem interface: the nmpc_ddp::DDPProblem virtuals, indexed by stage. */
struct DdpProblem
{
  int nx = 0;
  int N = 0;
  virtual ~DdpProblem() {}
  virtual int inputDim(int k) const = 0;
  virtual void stateEq(int k, const double * x, const double * u, double * xn) const = 0;
  virtual double runningCost(int k, const double * x, const double * u) const = 0;
  virtual double terminalCost(const double * x) const = 0;
  /** Fx: nx*nx row-major; Fu: nx*m row-major. */
  virtual void stateEqDeriv(int k, const double * x, const double * u, double * Fx, double * Fu) const = 0;
  /** Lx(nx) Lu(m) Lxx(nx*nx) Luu(m*m) Lxu(nx*m). */
  virtual void runningCostDeriv(int k,
                                const double * x,
                                const double * u,
                                double * Lx,
                                double * Lu,
                                double * Lxx,
                                double * Luu,
                                double * Lxu) const = 0;
  virtual void terminalCostDeriv(const double * x, double * Vx, double * Vxx) const = 0;
  virtual void inputLimits(int k, double * lo, double * hi) const
  {
    (void)k;
    (void)lo;
    (void)hi;
  }
};

struct DdpTrace
{
  int iter = 0;
  double cost = 0, lambda = 0, dlambda = 0, alpha = 0;
  int alpha_idx = -4; // see include/ccc_b200.h ccc_ddp_result_t::alpha_idx
  double k_rel_norm = 0, cost_update_actual = 0, cost_update_expected = 0;
  int backward_retries = 0;
};

struct DdpSolver
{
  DdpConfig cfg;
  const DdpProblem & p;
  int nx, N;
  std::vector<std::vector<double>> x, u, xc, uc, k_list, K_list;
  std::vector<double> cost, costc;
  std::vector<uint32_t> clamped_mask; // per stage, last completed backward pass
  std::vector<DdpTrace> trace;
  double lambda = 0, dlambda = 0, dV[2] = {0, 0};
  int retval = 0;
  // statistics only: backward passes, forward passes, BoxQP calls / iterations / factorisations / Armijo steps
  long long stats[6] = {0, 0, 0, 0, 0, 0};

  explicit DdpSolver(const DdpProblem & prob) : p(prob), nx(prob.nx), N(prob.N) {}

  static double sum_seq(const std::vector<double> & v)
  {
    double s = 0.0;
    for(double c : v) s = s + c;
    return s;
  }

  bool solve(const double * x0, const std::vector<std::vector<double>> & u_init)
  {
    lambda = cfg.initial_lambda;
    dlambda = cfg.initial_dlambda;
    x.assign(N + 1, std::vector<double>(nx, 0.0));
    xc = x;
    u = u_init;
    uc = u_init;
    cost.assign(N + 1, 0.0);
    costc = cost;
    // gains start from zero at every solve (the BoxQP warm start of the last stage reads
    // k_list[N-1] of the previous backward pass, zeros before the first one)
    k_list.assign(N, {});
    K_list.assign(N, {});
    clamped_mask.assign(N, 0);
    for(int k = 0; k < N; k++)
    {
      int m = p.inputDim(k);
      k_list[k].assign(m, 0.0);
      K_list[k].assign(static_cast<size_t>(m) * nx, 0.0);
    }
    // initial rollout
    std::memcpy(x[0].data(), x0, sizeof(double) * nx);
    for(int k = 0; k < N; k++)
    {
      p.stateEq(k, x[k].data(), u[k].data(), x[k + 1].data());
      cost[k] = p.runningCost(k, x[k].data(), u[k].data());
    }
    cost[N] = p.terminalCost(x[N].data());

    trace.clear();
    DdpTrace t0;
    t0.iter = 0;
    t0.cost = sum_seq(cost);
    t0.lambda = lambda;
    t0.dlambda = dlambda;
    trace.push_back(t0);

    retval = 0;
    for(int iter = 1; iter <= cfg.max_iter; iter++)
    {
      retval = procOnce(iter);
      if(retval != 0) break;
    }
    return retval == 1;
  }

  void increaseLambda()
  {
    dlambda = std::max(dlambda * cfg.lambda_factor, cfg.lambda_factor);
    lambda = std::max(lambda * dlambda, cfg.lambda_min);
  }
  void decreaseLambda()
  {
    dlambda = std::min(dlambda / cfg.lambda_factor, 1.0 / cfg.lambda_factor);
    lambda = (lambda * dlambda) * (lambda > cfg.lambda_min ? 1.0 : 0.0);
  }

  int procOnce(int iter)
  {
    trace.push_back(DdpTrace());
    DdpTrace & tr = trace.back();
    tr.iter = iter;
    auto finish = [&](int rv) {
      tr.cost = sum_seq(cost);
      tr.lambda = lambda;
      tr.dlambda = dlambda;
      return rv;
    };

    // STEP 1 (derivatives) is evaluated inside backwardPass stage by stage; the
    // derivatives depend only on (x,u), which do not change between retries.

    // STEP 2: backward pass
    while(!backwardPass())
    {
      tr.backward_retries++;
      increaseLambda();
      if(lambda > cfg.lambda_max)
      {
        tr.alpha_idx = -3;
        return finish(-1);
      }
    }

    // termination on small gradient
    // nmpc_ddp: max over the stages of k_list[i].norm() / (u_list[i].norm() + 1); the squared norms are
    // summed with the fixed 32-leaf tree like every other reduction over the input index
    double k_rel_norm = 0;
    for(int k = 0; k < N; k++)
    {
      const size_t m = k_list[k].size();
      if(choice(kChoiceKRelElementwise))
      {
        for(size_t j = 0; j < m; j++)
          k_rel_norm = std::max(k_rel_norm, std::fabs(k_list[k][j]) / (std::fabs(u[k][j]) + 1.0));
        continue;
      }
      double kk2[32], uu2[32];
      for(size_t j = 0; j < m; j++)
      {
        kk2[j] = k_list[k][j] * k_list[k][j];
        uu2[j] = u[k][j] * u[k][j];
      }
      const double kn = std::sqrt(tree_sum32(kk2, static_cast<int>(m)));
      const double un = std::sqrt(tree_sum32(uu2, static_cast<int>(m)));
      k_rel_norm = std::max(k_rel_norm, kn / (un + 1.0));
    }
    tr.k_rel_norm = k_rel_norm;
    if(k_rel_norm < cfg.k_rel_norm_thre && lambda < cfg.lambda_thre)
    {
      decreaseLambda();
      tr.alpha_idx = -2;
      return finish(1);
    }

    // STEP 3: line search
    bool success = false;
    double actual = 0, expected = 0;
    double J = sum_seq(cost);
    int n_alpha = static_cast<int>(cfg.alpha_list.size());
    for(int a = 0; a < n_alpha; a++)
    {
      double alpha = cfg.alpha_list[a];
      forwardPass(alpha);
      actual = J - sum_seq(costc);
      expected = -(alpha * fmad(alpha, dV[1], dV[0]));
      double ratio;
      if(expected > 0)
        ratio = actual / expected;
      else
        ratio = static_cast<double>((0 < actual) - (actual < 0));
      if(ratio > cfg.cost_update_ratio_thre)
      {
        success = true;
        tr.alpha = alpha;
        tr.alpha_idx = a;
        break;
      }
    }
    tr.cost_update_actual = actual;
    tr.cost_update_expected = expected;

    // STEP 4: accept or reject
    int rv = 0;
    if(success)
    {
      decreaseLambda();
      x = xc;
      u = uc;
      cost = costc;
      if(actual < cfg.cost_update_thre) rv = 1;
    }
    else
    {
      tr.alpha_idx = -1;
      increaseLambda();
      if(lambda > cfg.lambda_max) rv = -1;
    }
    return finish(rv);
  }

  bool backwardPass()
  {
    std::vector<double> Vx(nx), Vxx(nx * nx), Qx(nx), Qxx(nx * nx), T(nx * nx), Fx(nx * nx), Lx(nx), Lxx(nx * nx);
    std::vector<double> Vn(nx * nx);
    p.terminalCostDeriv(x[N].data(), Vx.data(), Vxx.data());
    dV[0] = dV[1] = 0.0;
    stats[0]++;

    for(int k = N - 1; k >= 0; k--)
    {
      const int m = p.inputDim(k);
      std::vector<double> Fu(nx * m), Lu(m), Luu(m * m), Lxu(nx * m), W(nx * m), Qu(m), Quu(m * m), QuuF(m * m),
          Qux(m * nx);
      p.stateEqDeriv(k, x[k].data(), u[k].data(), Fx.data(), Fu.data());
      p.runningCostDeriv(k, x[k].data(), u[k].data(), Lx.data(), Lu.data(), Lxx.data(), Luu.data(), Lxu.data());

      // Q-function
      for(int i = 0; i < nx; i++) Qx[i] = Lx[i] + dot_seq(Fx.data() + i, nx, Vx.data(), 1, nx);
      for(int j = 0; j < m; j++) Qu[j] = Lu[j] + dot_seq(Fu.data() + j, m, Vx.data(), 1, nx);
      for(int i = 0; i < nx; i++) // T = Vxx Fx
        for(int j = 0; j < nx; j++) T[i * nx + j] = dot_seq(Vxx.data() + i * nx, 1, Fx.data() + j, nx, nx);
      for(int i = 0; i < nx; i++)
        for(int j = 0; j < nx; j++)
          Qxx[i * nx + j] = Lxx[i * nx + j] + dot_seq(Fx.data() + i, nx, T.data() + j, nx, nx);
      for(int i = 0; i < nx; i++) // W = Vxx Fu
        for(int j = 0; j < m; j++) W[i * m + j] = dot_seq(Vxx.data() + i * nx, 1, Fu.data() + j, m, nx);
      // Quu is evaluated on its lower triangle and mirrored, so that it is symmetric bit for bit
      // (Eigen::LLT reads the lower triangle only; the engine stores one triangle)
      for(int i = 0; i < m; i++)
        for(int j = 0; j <= i; j++)
        {
          Quu[i * m + j] = Luu[i * m + j] + dot_seq(Fu.data() + i, m, W.data() + j, m, nx);
          Quu[j * m + i] = Quu[i * m + j];
        }
      for(int j = 0; j < m; j++) // Qux = Lxu' + Fu' (Vxx Fx)
        for(int i = 0; i < nx; i++) Qux[j * nx + i] = Lxu[i * m + j] + dot_seq(Fu.data() + j, m, T.data() + i, nx, nx);

      // regularisation, reg_type 1: Quu_F = Quu + lambda I, Qux_reg = Qux
      QuuF = Quu;
      for(int j = 0; j < m; j++) QuuF[j * m + j] = Quu[j * m + j] + lambda;

      std::vector<double> & kk = k_list[k];
      std::vector<double> & KK = K_list[k];
      uint32_t mask = 0;
      if(m > 0)
      {
        if(cfg.with_input_constraint)
        {
          std::vector<double> lo(m), hi(m), k0(m, 0.0);
          p.inputLimits(k, lo.data(), hi.data());
          for(int j = 0; j < m; j++)
          {
            lo[j] = lo[j] - u[k][j];
            hi[j] = hi[j] - u[k][j];
          }
          // warm start from the gain of the next stage (iLQG.m: k(:,min(i+1,N-1)))
          const std::vector<double> & kn =
              choice(kChoiceBoxQpWarmSameStage) ? k_list[k] : k_list[std::min(k + 1, N - 1)];
          if(static_cast<int>(kn.size()) == m && !choice(kChoiceBoxQpColdStart)) k0 = kn;
          BoxQp qp;
          qp.cfg = cfg.boxqp;
          int r = qp.solve(QuuF.data(), Qu.data(), lo.data(), hi.data(), k0.data(), m);
          stats[2]++;
          stats[3] += qp.iters;
          stats[4] += qp.nfactor;
          stats[5] += qp.ls_steps;
          if(r < 1) return false;
          kk = qp.x;
          KK.assign(static_cast<size_t>(m) * nx, 0.0);
          bool any_free = false;
          for(int j = 0; j < m; j++)
          {
            if(qp.clamped[j])
              mask |= (1u << j);
            else
              any_free = true;
          }
          if(any_free)
          {
            const FreeLlt & llt = qp.llt;
            std::vector<double> rhs(llt.nf);
            for(int c = 0; c < nx; c++)
            {
              for(int a = 0; a < llt.nf; a++) rhs[a] = Qux[llt.idx[a] * nx + c];
              llt.solve(rhs.data(), 1);
              for(int a = 0; a < llt.nf; a++) KK[llt.idx[a] * nx + c] = -rhs[a];
            }
          }
        }
        else
        {
          FreeLlt llt;
          std::vector<int> all(m);
          for(int j = 0; j < m; j++) all[j] = j;
          if(!llt.compute(QuuF.data(), m, all)) return false;
          std::vector<double> rhs(m);
          for(int j = 0; j < m; j++) rhs[j] = Qu[j];
          llt.solve(rhs.data(), 1);
          kk.assign(m, 0.0);
          for(int j = 0; j < m; j++) kk[j] = -rhs[j];
          KK.assign(static_cast<size_t>(m) * nx, 0.0);
          for(int c = 0; c < nx; c++)
          {
            for(int j = 0; j < m; j++) rhs[j] = Qux[j * nx + c];
            llt.solve(rhs.data(), 1);
            for(int j = 0; j < m; j++) KK[j * nx + c] = -rhs[j];
          }
        }
      }
      clamped_mask[k] = mask;

      // cost-to-go update (unregularised Quu)
      if(choice(kChoiceVxxRegularised)) Quu = QuuF;
      std::vector<double> Quuk(m), QuuK(m * nx), t1(m), t2(m);
      for(int i = 0; i < m; i++)
      {
        Quuk[i] = dot4(Quu.data() + i * m, 1, kk.data(), 1, m);
        for(int c = 0; c < nx; c++) QuuK[i * nx + c] = dot_seq(Quu.data() + i * m, 1, KK.data() + c, nx, m);
        t1[i] = kk[i] * Qu[i];
        t2[i] = kk[i] * Quuk[i];
      }
      dV[0] = dV[0] + tree_sum32(t1.data(), m);
      dV[1] = fmad(0.5, tree_sum32(t2.data(), m), dV[1]);

      for(int c = 0; c < nx; c++)
      {
        double a1 = dot_seq(KK.data() + c, nx, Quuk.data(), 1, m);
        double a2 = dot_seq(KK.data() + c, nx, Qu.data(), 1, m);
        double a3 = dot_seq(Qux.data() + c, nx, kk.data(), 1, m);
        Vx[c] = ((Qx[c] + a1) + a2) + a3;
      }
      for(int a = 0; a < nx; a++)
        for(int b = 0; b < nx; b++)
        {
          double s1 = dot_seq(KK.data() + a, nx, QuuK.data() + b, nx, m);
          double s2 = dot_seq(KK.data() + a, nx, Qux.data() + b, nx, m);
          double s3 = dot_seq(Qux.data() + a, nx, KK.data() + b, nx, m);
          Vn[a * nx + b] = ((Qxx[a * nx + b] + s1) + s2) + s3;
        }
      for(int a = 0; a < nx; a++)
        for(int b = 0; b < nx; b++) Vxx[a * nx + b] = 0.5 * (Vn[a * nx + b] + Vn[b * nx + a]);
    }
    return true;
  }

  void forwardPass(double alpha)
  {
    xc[0] = x[0];
    stats[1]++;
    std::vector<double> dx(nx), lo, hi;
    for(int k = 0; k < N; k++)
    {
      const int m = p.inputDim(k);
      for(int c = 0; c < nx; c++) dx[c] = xc[k][c] - x[k][c];
      uc[k].assign(m, 0.0);
      if(cfg.with_input_constraint)
      {
        lo.assign(m, 0.0);
        hi.assign(m, 0.0);
        p.inputLimits(k, lo.data(), hi.data());
      }
      for(int j = 0; j < m; j++)
      {
        double fb = dot_seq(K_list[k].data() + j * nx, 1, dx.data(), 1, nx);
        double v = fmad(alpha, k_list[k][j], u[k][j]) + fb;
        if(cfg.with_input_constraint) v = clampd(v, lo[j], hi[j]);
        uc[k][j] = v;
      }
      p.stateEq(k, xc[k].data(), uc[k].data(), xc[k + 1].data());
      costc[k] = p.runningCost(k, xc[k].data(), uc[k].data());
    }
    costc[N] = p.terminalCost(xc[N].data());
  }
};
} // namespace oracle
