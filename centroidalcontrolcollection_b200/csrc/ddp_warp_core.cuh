// ddp_warp_core.cuh — one warp solves one DDP problem with ridge-force inputs (<= 32 per stage,
// box limits) start to finish; generic over the dynamics model M (CentroidalModel: 9 states,
// SrbModel: 12 states).
//
// Replaces (reference file:line):
//   nmpc_ddp::DDPSolver<NX,Dynamic>::solve   call sites src/DdpCentroidal.cpp:229,233,
//                                            src/DdpSingleRigidBody.cpp:299,303
//   the DdpProblem callbacks it makes        src/DdpCentroidal.cpp:21-177,
//                                            src/DdpSingleRigidBody.cpp:41-245  (model_*.cuh)
//   input limits                             src/DdpCentroidal.cpp:202-210, src/DdpSingleRigidBody.cpp:272-280
// Algorithm and evaluation order: oracle/ddp.hpp (bit-exact where the model uses no libm).
//
// Work split inside the warp: lane j owns input (ridge) j — its vertex/ridge pair, its
// column of Fu (6 non-zero rows starting at M::R0), row j of Quu (32 FP64 registers), row j of
// Qux / K / Quu*K.  State-sized objects (Vx, Vxx, Fx, Vxx*Fx, Qxx; NX x NX) live in the warp's
// shared-memory slice and their entries are dealt round-robin to the lanes.  Derivatives Fx/Fu
// are rebuilt from (x,u) and the stage tables in registers, never stored.  The gain lists k, K
// spill to HBM in a lane-contiguous layout ([stage][1+NX][32]): every access is a 256-byte row.
//
// A solve can be suspended at an iteration boundary and resumed later by any warp (DdpResume):
// the kernel uses this to time-slice problems round-robin, so that the rare problems that need
// hundreds of DDP iterations start early instead of forming a serial tail.
#pragma once
#include "boxqp_warp.cuh"

namespace ccc
{
// plain-data mirror of ccc_ddp_config_t's solver part (kept separate so that this header
// compiles without the public C header)
struct DdpCfg
{
  int with_input_constraint, max_iter, n_alpha;
  double initial_lambda, initial_dlambda, lambda_factor, lambda_min, lambda_max;
  double k_rel_norm_thre, lambda_thre, cost_update_ratio_thre, cost_update_thre;
  double alpha[16];
  BoxQpCfg boxqp;
};

/** What survives between two visits of a suspended solve (everything else lives in the
 *  trajectory / gain buffers in HBM or is rebuilt). */
struct DdpResume
{
  double lambda, dlambda, J;
  int cur, iter;
};

// The once-per-stage state-sized loops of the backward pass (T = Vxx Fx, Qxx, Qux rows) are unrolled M::STAGE_UNROLL
// iterations at a time (same operations in the same order whatever the factor).  Fully unrolled they are straight-line
// code that every stage fetches cold: with 12 states the per-stage straight-line part of the kernel is 44 KB and a quarter
// of all stall samples are instruction fetches (profiles/r02_summary.md §5); rolled three at a time DdpSingleRigidBody
// runs 5.5 % faster, while the 9-state DdpCentroidal kernel is 1.5 % faster fully unrolled (r02g_ab_stage_unroll.txt).
#define CCC_STAGE_UNROLL CCC_UNROLL_N(M::STAGE_UNROLL)
#ifndef CCC_TILE_UNROLL
#  define CCC_TILE_UNROLL 2 // unroll factor of the rolled (compact-code) Quu assembly / Quu K loops
#endif

template<class M>
struct DdpParams
{
  int N, B, S;
  // Stage k of a problem reads entry tab_off + k * tab_stride of its schedule's tables.  A plain solve has one
  // entry per stage (tab_len = N, ref_len = N + 1, offset 0, stride 1); the closed loop keeps the tables on the
  // plant's time grid and slides the horizon over them (offset = control cycle, stride = horizon_dt / sim_dt).
  int tab_len = 0, ref_len = 0, tab_off = 0, tab_stride = 1;
  const int * sched_id;   // [B]
  const int * m;          // [S][tab_len]
  const double * tab;     // [S][tab_len][M::TAB_ROWS][32] packed stage tables (lane-contiguous rows)
  const double * ref;     // [S][ref_len][M::NREF] reference of the first NREF states
  double w_run[M::NX + 1]; // diagonal running weights of the states, then the force weight
  double w_term[M::NX];    // diagonal terminal weights
  typename M::Params mp;   // model constants (dt, mass, ...)
  double u_lo, u_hi;
  const double * x0;     // [B][NX]
  const double * u_init; // [B][N][32] or null
  DdpCfg cfg;
  int chunk_iters; // > 0: suspend a solve after this many iterations per visit
  int abort_ok;    // 1: all cost weights >= 0 and cost_update_ratio_thre >= 0, so a line-search rollout whose partial
                   // cost already exceeds the nominal cost is certain to be rejected and may stop early
  // workspace (device)
  double * xbuf;  // [2][B][N+1][NX]  nominal / candidate state trajectories (ping-pong)
  double * ubuf;  // [2][B][N][32]
  double * gains; // [B][N][1+NX][32]  k (row 0) and the NX columns of K
  DdpResume * resume; // [B]
  // outputs (device, nullable)
  double * out_x;
  double * out_u;
  double * out_cost;
  int * out_iters;
  int * out_status;
  int trace_len;
  signed char * out_alpha_idx;
  double * out_lambda;
  unsigned * out_clamped;
};

/** Feature bits of the solver core (template parameter FEAT), kept selectable so that every one of them has a measured
 *  A/B on the same build (profiles/r02_summary.md):
 *   kFeatAbort — a line-search rollout stops as soon as its partial cost exceeds the nominal cost (costs are sums of
 *     non-negative terms, so the candidate is rejected whatever follows; results are unchanged);
 *   kFeatTma — the gain lists are laid out one row per input ([stage][32][NXP]: K(NX), k) and move through the TMA
 *     engine: the backward pass writes the rows of a stage from its K staging buffer with ONE bulk store
 *     (cp.async.bulk.global.shared::cta) of m rows, the line-search rollouts stream them back through a 2-3 slot
 *     shared-memory ring filled by bulk loads completing on mbarriers (cp.async.bulk.shared::cluster.global), issued
 *     two stages ahead by lane 0.  Without the bit: [stage][1+NX][32] rows, 10 STG / 11 LDG per lane and stage. */
constexpr int kFeatAbort = 1;
constexpr int kFeatTma = 2;
constexpr int kFeatAll = kFeatAbort | kFeatTma;
constexpr int kFeatDefault = kFeatAbort; // kFeatTma measured 7 % slower on config 3 (profiles/r02_summary.md): built, tested, off

// per-warp shared-memory slice, offsets in doubles
template<int NX, int NXP>
struct SmLayout
{
  static constexpr int NN = NX * NX + ((NX * NX) & 1);         // NX x NX matrix, even-padded
  static constexpr int NV = NX + (NX & 1);                      // NX vector, even-padded
  static constexpr int A = 0;                                   // 32 x kLda tile: strict lower = compact L (L D L' of the free block)
  static constexpr int VB = A + 32 * kLda;                      // BoxQP column buffers (2 * kCbStride) + publish vector (32)
  static constexpr int IDX = VB + 2 * kCbStride + 32;           // 32 ints: free list of the compact factor
  static constexpr int VXX = IDX + 16;
  static constexpr int VX = VXX + NN;
  static constexpr int FX = VX + NV;                            // dense Fx
  static constexpr int T = FX + NN;                             // Vxx * Fx
  static constexpr int QXX = T + NN;
  static constexpr int S2 = QXX + NN;                           // scratch NX x NX
  static constexpr int QX = S2 + NN;
  static constexpr int QUXR = QX + NV;                          // [32][NXP] Qux rows (parked here across BoxQP: registers)
  static constexpr int SYM = QUXR + 32 * NXP;                   // 32 x kLda tile: Quu_F in full (both triangles), rows 16-byte aligned
  static constexpr int BAR = SYM + 32 * kLda;                   // 4 x 8 bytes: mbarriers of the gain ring (kFeatTma)
  static constexpr int TOTAL = BAR + 4;
  static_assert(SYM % 2 == 0 && TOTAL % 2 == 0, "16-byte alignment of the tile rows and of the next warp's slice");
  // gain ring of the line-search rollouts: slots of 32 rows x NXP inside the factor tile A (dead during rollouts)
  static constexpr int SLOT = 32 * NXP;
  static constexpr int RING = 3 * SLOT <= 32 * kLda ? 3 : 2;
  static_assert(RING * SLOT <= 32 * kLda, "gain ring must fit in the factor tile");
  // aliases inside A (+ the start of VB), valid while no factor is alive
  static constexpr int WT = A;                                  // [32][6]  the 6 live rows of Vxx*Fu, transposed
  static constexpr int KB = A;                                  // [32][NXP] K rows
  static constexpr int ZB = A + 32 * NXP;                       // [32][NXP] (Quu K) rows
  static constexpr int VB0 = A + 64 * NXP;                      // three 32-vectors: k, Quu k, Qu
  static constexpr int VB1 = VB0 + 32;
  static constexpr int VB2 = VB0 + 64;
  static_assert(VB2 + 32 <= IDX, "cost-to-go scratch must fit in the tile + BoxQP buffers");
  static_assert(NXP % 2 == 0 && NXP >= NX + 1, "row stride of the K/Z/Q buffers: NX gains + k, even");
};

template<class M, bool kConstrained, int FEAT = kFeatDefault>
struct DdpWarp
{
  static constexpr int NX = M::NX, NXP = M::NXP, R0 = M::R0, NREF = M::NREF;
  static constexpr bool kTma = (FEAT & kFeatTma) != 0, kAbort = (FEAT & kFeatAbort) != 0;
#if defined(CCC_NO_ROLLED_TILE_LOOPS)
  static constexpr bool kRolled = false;
#elif defined(CCC_FORCE_ROLLED_TILE_LOOPS)
  static constexpr bool kRolled = true;
#else
  static constexpr bool kRolled = M::STAGE_UNROLL < NX; // compact-code forms of the Quu assembly and Quu K loops too
#endif
  static constexpr int GBLK = 32 * NXP; // doubles per stage of the gain lists (either layout fits: NXP >= NX + 1)
  using sm = SmLayout<NX, NXP>;
  const DdpParams<M> & P;
  double * s; // this warp's shared-memory slice
  int b, sched, lane;
  int cur; // index of the nominal trajectory buffer (0/1)
  double lambda, dlambda, dV0, dV1, J;
  // Per-stage accumulations of the backward pass, one quantity per lane group (warp_sum4_scattered): lanes 0-7 dV[0],
  // lanes 8-15 dV[1], lanes 16-23 max_k ||k_k|| / (||u_k|| + 1) (nmpc_ddp's small-gradient measure), lanes 24-31 unused.
  double acc4;
  double krel;
  unsigned * ring_parity; // phase bits of the ring's mbarriers: they live as long as the kernel, not the solve

  CCC_DEV DdpWarp(const DdpParams<M> & p, double * smem, int prob, unsigned * parity)
  : P(p), s(smem), b(prob), sched(p.sched_id[prob]), lane(lane_id()), cur(0), lambda(0), dlambda(0), dV0(0), dV1(0),
    J(0), acc4(0), krel(0), ring_parity(parity)
  {
  }

  /** Once per warp and kernel: the mbarriers of the gain ring (one arrival each: lane 0's expect_tx). */
  CCC_DEV static void init_warp(double * smem)
  {
    if(kTma && lane_id() == 0)
    {
      unsigned long long * bars = reinterpret_cast<unsigned long long *>(smem + sm::BAR);
      for(int i = 0; i < sm::RING; i++) mbar_init(bars + i, 1);
    }
    warp_sync();
  }

  CCC_DEV double * xtraj(int which) const { return P.xbuf + ((size_t)which * P.B + b) * (size_t)(P.N + 1) * NX; }
  CCC_DEV double * utraj(int which) const { return P.ubuf + ((size_t)which * P.B + b) * (size_t)P.N * 32; }
  CCC_DEV double * gain(int k) const { return P.gains + ((size_t)b * P.N + k) * GBLK; }
  CCC_DEV unsigned long long * ring_bar(int slot) const { return reinterpret_cast<unsigned long long *>(s + sm::BAR) + slot; }
  CCC_DEV double * ring_slot(int slot) const { return s + sm::A + slot * sm::SLOT; }
  /** Lane 0 starts the bulk load of stage k's gain rows into ring slot k % RING (nothing to load for m = 0). */
  CCC_DEV void ring_issue(int k) const
  {
    const int m = stage_m(k);
    if(lane == 0 && m > 0)
    {
      const int slot = k % sm::RING;
      const unsigned bytes = (unsigned)m * (NXP * 8);
      fence_proxy_async_smem(); // the slot's previous contents were read through the generic proxy
      mbar_arrive_expect_tx(ring_bar(slot), bytes);
      tma_bulk_g2s(ring_slot(slot), gain(k), bytes, ring_bar(slot));
    }
  }
  /** All lanes wait for stage k's rows (issued earlier by ring_issue(k)); `parity` holds one phase bit per slot. */
  CCC_DEV void ring_wait(int k, int m, unsigned & parity) const
  {
    if(m > 0)
    {
      const int slot = k % sm::RING;
      mbar_wait(ring_bar(slot), (parity >> slot) & 1u);
      parity ^= 1u << slot;
    }
  }
  CCC_DEV int entry(int k) const { return P.tab_off + k * P.tab_stride; }
  CCC_DEV int stage_m(int k) const { return ldg(P.m + (size_t)sched * P.tab_len + entry(k)); }
  CCC_DEV const double * stage_tab(int k) const { return P.tab + ((size_t)sched * P.tab_len + entry(k)) * (32 * M::TAB_ROWS); }
  CCC_DEV const double * ref(int k) const { return P.ref + ((size_t)sched * P.ref_len + entry(k)) * NREF; }

  /** Lane 0 stores the (warp-uniform) state vector: NX independent stores instead of an NX-deep select chain. */
  CCC_DEV void storeX(double * dst, const double (&x)[NX]) const
  {
    if(lane == 0)
    {
      CCC_UNROLL
      for(int i = 0; i < NX; i++) dst[i] = x[i];
    }
  }

  /** sum_a w[a] (x[a] - r[a])^2 over the referenced states, then sum_a w[a] x[a]^2 over the rest
   *  (the diagonal quadratic costs of src/DdpCentroidal.cpp:66-83, src/DdpSingleRigidBody.cpp:93-113). */
  CCC_DEV static double quad(const double * w, const double (&x)[NX], const double * r)
  {
    double c = 0.0;
    CCC_UNROLL
    for(int a = 0; a < NREF; a++)
    {
      double d = x[a] - r[a];
      c = dfma(w[a], d * d, c);
    }
    CCC_UNROLL
    for(int a = NREF; a < NX; a++) c = dfma(w[a], x[a] * x[a], c);
    return c;
  }

  /** x <- stateEq(x, u); returns runningCost(x, u) of the stage (all lanes hold x; lane j holds u_j). */
  CCC_DEV double step_and_cost(int k, int m, double (&x)[NX], double u)
  {
    double rr[NREF];
    CCC_UNROLL
    for(int a = 0; a < NREF; a++) rr[a] = ldg(ref(k) + a);
    const double q = quad(P.w_run, x, rr);
    const double usq = M::step(*this, k, m, x, u);
    return dfma(0.5 * P.w_run[NX], usq, 0.5 * q);
  }

  CCC_DEV double terminal_cost(const double (&x)[NX])
  {
    double rr[NREF];
    CCC_UNROLL
    for(int a = 0; a < NREF; a++) rr[a] = ldg(ref(P.N) + a);
    return 0.5 * quad(P.w_term, x, rr);
  }

  /** Initial rollout or line-search forward pass into trajectory buffer `dst`.
   *  Returns the summed cost of the produced trajectory. */
  CCC_DEV double rollout(int dst, double alpha, bool initial)
  {
    const int N = P.N;
    double * xd = xtraj(dst);
    double * ud = utraj(dst);
    const double * xn = xtraj(cur);
    const double * un = utraj(cur);
    double x[NX];
    CCC_UNROLL
    for(int i = 0; i < NX; i++) x[i] = initial ? ldg(P.x0 + (size_t)b * NX + i) : xn[i];
    double Jc = 0.0;
    // line-search passes: the gains and the nominal (x, u) of a stage do not depend on the stages before
    // it, so they are fetched one stage ahead into registers (gq/uq/xq) and the HBM / L2 latency hides
    // behind the previous stage's feedback -> reduce -> step chain; lanes >= m fetch values they never use
    double gq[1 + NX], xq[NX], uq = 0.0;
    CCC_UNROLL
    for(int c = 0; c <= NX; c++) gq[c] = 0.0;
    CCC_UNROLL
    for(int c = 0; c < NX; c++) xq[c] = 0.0;
    unsigned parity = *ring_parity;
    int issued = 0; // kTma: stages [0, issued) have had their bulk load started
    if(!initial)
    {
      if(kTma)
      {
        warp_sync(); // the factor tile (ring slots) is dead: every lane is past the backward pass
        CCC_NOUNROLL
        for(; issued < sm::RING - 1 && issued < N; issued++) ring_issue(issued);
      }
      else
      {
        const double * g = gain(0);
        CCC_UNROLL
        for(int c = 0; c <= NX; c++) gq[c] = g[c * 32 + lane];
      }
      uq = un[lane];
      CCC_UNROLL
      for(int c = 0; c < NX; c++) xq[c] = xn[c];
    }
    CCC_NOUNROLL
    for(int k = 0; k < N; k++)
    {
      const int m = stage_m(k);
      const bool active = lane < m;
      double u;
      if(initial)
      {
        u = (active && P.u_init) ? ldg(P.u_init + ((size_t)b * N + k) * 32 + lane) : 0.0;
      }
      else
      {
        double gk[1 + NX], dx[NX];
        const double uj = uq;
        if(kTma)
        {
          // the slot two stages ahead was last read at stage k - 1 (every lane is past it: the stage's reductions
          // synchronise the warp); start its refill, then take this stage's row from its slot
          if(issued < N)
          {
            warp_sync();
            ring_issue(issued);
            issued++;
          }
          ring_wait(k, m, parity);
          const double * row = ring_slot(k % sm::RING) + lane * NXP;
          CCC_UNROLL
          for(int c = 0; c < NXP; c += 2)
          {
            const d2 v = active ? ld2(row + c) : d2{0.0, 0.0};
            if(c < NX) gk[1 + c] = v.x;
            if(c == NX) gk[0] = v.x;
            if(c + 1 < NX) gk[2 + c] = v.y;
            if(c + 1 == NX) gk[0] = v.y;
          }
        }
        else
        {
          CCC_UNROLL
          for(int c = 0; c <= NX; c++) gk[c] = gq[c];
        }
        CCC_UNROLL
        for(int c = 0; c < NX; c++) dx[c] = x[c] - xq[c];
        if(k + 1 < N)
        {
          if(!kTma)
          {
            const double * g = gain(k + 1);
            CCC_UNROLL
            for(int c = 0; c <= NX; c++) gq[c] = g[c * 32 + lane];
          }
          uq = un[(size_t)(k + 1) * 32 + lane];
          CCC_UNROLL
          for(int c = 0; c < NX; c++) xq[c] = xn[(size_t)(k + 1) * NX + c];
          prefetch_span(stage_tab(k + 1), 32 * M::TAB_ROWS * 8);
        }
        double fb = 0.0;
        CCC_UNROLL
        for(int c = 0; c < NX; c++) fb = dfma(gk[1 + c], dx[c], fb);
        u = dfma(alpha, gk[0], uj) + fb;
        if(kConstrained) u = clampd(u, P.u_lo, P.u_hi);
        if(!active) u = 0.0;
      }
      ud[(size_t)k * 32 + lane] = u;
      storeX(xd + (size_t)k * NX, x);
      const double c = step_and_cost(k, m, x, u);
      Jc = Jc + c;
      if(kAbort && !initial && P.abort_ok && Jc > J)
      {
        // every later term is >= 0: this candidate costs more than the nominal trajectory and will be rejected.
        // Drain the bulk loads that are still in flight (the ring lives in the factor tile of the next backward pass).
        if(kTma)
        {
          CCC_NOUNROLL
          for(int kk = k + 1; kk < issued; kk++) ring_wait(kk, stage_m(kk), parity);
          *ring_parity = parity;
        }
        return Jc;
      }
    }
    if(kTma) *ring_parity = parity;
    storeX(xd + (size_t)N * NX, x);
    Jc = Jc + terminal_cost(x);
    return Jc;
  }

  // ---- backward pass ----------------------------------------------------------------------
  /** One stage of the backward recursion.  Returns false if BoxQP / LLT failed. */
  CCC_DEV bool backward_stage(int k, double & k_next, int & m_next)
  {
    const int N = P.N;
    const int m = stage_m(k);
    const bool active = lane < m;
    const double * xn = xtraj(cur) + (size_t)k * NX;
    const double u = active ? utraj(cur)[(size_t)k * 32 + lane] : 0.0;

    // model: this lane's column of Fu (rows R0..R0+5) and the stage-dependent entries of Fx
    double Fu[6];
    M::lane_derivs(*this, k, m, xn, u, Fu);

    const double * Fx = s + sm::FX;
    const double * Vxx = s + sm::VXX;
    const double * Vx = s + sm::VX;
    double * T = s + sm::T;
    double * Qxx = s + sm::QXX;
    double * Qx = s + sm::QX;
    // T = Vxx Fx: the NQE entries of a lane advance together (NQE independent fma chains, not one after the other)
    constexpr int NQE = (NX * NX + 31) / 32;
    int ei[NQE], ej[NQE];
    CCC_UNROLL
    for(int q = 0; q < NQE; q++)
    {
      const int e = lane + 32 * q;
      const int ee = e < NX * NX ? e : 0;
      ei[q] = ee / NX;
      ej[q] = ee - NX * ei[q];
    }
    {
      double acc[NQE];
      CCC_UNROLL
      for(int q = 0; q < NQE; q++) acc[q] = 0.0;
      CCC_STAGE_UNROLL
      for(int c = 0; c < NX; c++)
      {
        CCC_UNROLL
        for(int q = 0; q < NQE; q++) acc[q] = dfma(Vxx[ei[q] * NX + c], Fx[c * NX + ej[q]], acc[q]);
      }
      CCC_UNROLL
      for(int q = 0; q < NQE; q++)
        if(lane + 32 * q < NX * NX) T[lane + 32 * q] = acc[q];
    }
    if(lane < NX)
    {
      double rr = lane < NREF ? ldg(ref(k) + lane) : 0.0;
      double xl = xn[lane];
      double lx = lane < NREF ? P.w_run[lane] * (xl - rr) : P.w_run[lane] * xl;
      double acc = 0.0;
      CCC_UNROLL
      for(int c = 0; c < NX; c++) acc = dfma(Fx[c * NX + lane], Vx[c], acc);
      Qx[lane] = lx + acc;
    }
    warp_sync();
    // Qxx = Lxx + Fx' T
    {
      double acc[NQE];
      CCC_UNROLL
      for(int q = 0; q < NQE; q++) acc[q] = 0.0;
      CCC_STAGE_UNROLL
      for(int c = 0; c < NX; c++)
      {
        CCC_UNROLL
        for(int q = 0; q < NQE; q++) acc[q] = dfma(Fx[c * NX + ei[q]], T[c * NX + ej[q]], acc[q]);
      }
      CCC_UNROLL
      for(int q = 0; q < NQE; q++)
        if(lane + 32 * q < NX * NX) Qxx[lane + 32 * q] = (ei[q] == ej[q] ? P.w_run[ei[q]] : 0.0) + acc[q];
    }

    if(m == 0)
    {
      // no input: Vx = Qx, Vxx = sym(Qxx)
      warp_sync();
      double * Vxxw = s + sm::VXX;
      double * Vxw = s + sm::VX;
      CCC_NOUNROLL
      for(int e = lane; e < NX * NX; e += 32)
      {
        const int i = e / NX, j = e - NX * i;
        Vxxw[e] = 0.5 * (Qxx[e] + Qxx[j * NX + i]);
      }
      if(lane < NX) Vxw[lane] = Qx[lane];
      if(P.out_clamped && lane == 0) P.out_clamped[(size_t)b * N + k] = 0u;
      warp_sync();
      k_next = 0.0;
      m_next = 0;
      return true;
    }

    // Qu
    double Qu;
    {
      double acc = 0.0;
      CCC_UNROLL
      for(int c = 0; c < 6; c++) acc = dfma(Fu[c], Vx[R0 + c], acc);
      Qu = M::lu(*this, k, u) + acc;
    }
    // W = Vxx Fu, rows R0..R0+5, published transposed: WT[lane][0..5]
    if(kTma)
    {
      // WT aliases the K staging rows that the previous stage's bulk store may still be reading
      if(lane == 0) bulk_wait_read_all();
      warp_sync();
    }
    double * WT = s + sm::WT;
    double quu_diag; // Quu(lane, lane) = luu + Fu' (Vxx Fu) of this lane: from the registers, no 32-way select
    {
      double w[6];
      CCC_UNROLL
      for(int r = 0; r < 6; r++)
      {
        double acc = 0.0;
        CCC_UNROLL
        for(int c = 0; c < 6; c++) acc = dfma(Vxx[(R0 + r) * NX + R0 + c], Fu[c], acc);
        w[r] = active ? acc : 0.0;
      }
      CCC_UNROLL
      for(int r = 0; r < 6; r++) WT[lane * 6 + r] = w[r];
      double acc = 0.0;
      CCC_UNROLL
      for(int r = 0; r < 6; r++) acc = dfma(Fu[r], w[r], acc);
      quu_diag = M::luu(*this) + acc;
    }
    // Qux row of this lane: Fu' (Vxx Fx), parked in shared memory until the gain solve and the
    // cost-to-go update need it (rows of inactive lanes are +0.0)
    double * QUXR = s + sm::QUXR;
    CCC_STAGE_UNROLL
    for(int c = 0; c < NX; c++)
    {
      double acc = 0.0;
      CCC_UNROLL
      for(int r = 0; r < 6; r++) acc = dfma(Fu[r], T[(R0 + r) * NX + c], acc);
      QUXR[lane * NXP + c] = active ? 0.0 + acc : 0.0;
    }
    warp_sync();
    // Quu row (lower triangle is the definition; mirrored through A's upper triangle)
    double H[32];
    CCC_UNROLL
    for(int j = 0; j < 32; j++) H[j] = 0.0;
    double * A = s + sm::A;
    double * S = s + sm::SYM;
    if(kRolled)
    {
      // compact-code form: a rolled loop over the columns that writes each entry straight into the tile (both
      // triangles); the row comes back into registers with load_sym_row below, as after every factorisation
      CCC_UNROLL_N(CCC_TILE_UNROLL)
      for(int j = 0; j < m; j++)
      {
        double acc = 0.0;
        CCC_UNROLL
        for(int r = 0; r < 3; r++)
        {
          d2 w = ld2(WT + j * 6 + 2 * r);
          acc = dfma(Fu[2 * r], w.x, acc);
          acc = dfma(Fu[2 * r + 1], w.y, acc);
        }
        const double h = 0.0 + acc;
        if(active && j < lane)
        {
          S[lane * kLda + j] = h;
          S[j * kLda + lane] = h;
        }
      }
    }
    else
    {
      CCC_UNROLL
      for(int j = 0; j < 32; j++)
      {
        if(j == 16 && m <= 16) break;
        double acc = 0.0;
        CCC_UNROLL
        for(int r = 0; r < 3; r++)
        {
          d2 w = ld2(WT + j * 6 + 2 * r);
          acc = dfma(Fu[2 * r], w.x, acc);
          acc = dfma(Fu[2 * r + 1], w.y, acc);
        }
        H[j] = 0.0 + acc; // off-diagonal entries (the diagonal is quu_diag above)
      }
      warp_sync(); // everyone is done reading WT (aliases A)
      // the lower triangle is the definition (oracle); it is written in row form and mirrored in column form, so that
      // every later reload of a row (after each factorisation, which borrows the registers) is 16 aligned LDS.128 and
      // the compact gather a plain indexed read of one row
      CCC_UNROLL
      for(int j = 0; j < 31; j++)
      {
        if(j == 16 && m <= 16) break;
        if(active && j < lane)
        {
          S[lane * kLda + j] = H[j];
          S[j * kLda + lane] = H[j];
        }
      }
    }
    if(active) S[lane * kLda + lane] = quu_diag + lambda;
    warp_sync();
    load_sym_row(H, S, m);

    // gains
    double kk = 0.0;
    double K[NX];
    unsigned clamped = 0;
    int * idxbuf = reinterpret_cast<int *>(s + sm::IDX);
    if(kConstrained)
    {
      const double lo = P.u_lo - u, hi = P.u_hi - u;
      // warm start: gain of the next stage if the dimensions agree (iLQG.m: k(:,min(i+1,N-1)))
      double x0;
      if(k == N - 1)
        x0 = active ? (kTma ? ldcg(gain(k) + lane * NXP + NX) : gain(k)[lane]) : 0.0; // (written by a bulk store: not through L1)
      else
        x0 = (m_next == m) ? k_next : 0.0;
      BoxQpOut r = boxqp_warp(H, S, A, s + sm::VB, idxbuf, Qu, lo, hi, x0, m, P.cfg.boxqp);
      if(r.retval < 1) return false;
      kk = active ? x0 : 0.0;
      clamped = r.clamped;
      const unsigned active_mask = m >= 32 ? 0xffffffffu : ((1u << m) - 1u);
      if(clamped != active_mask)
      {
        // K[free,:] = -(Quu_F[free,free])^-1 Qux[free,:] with BoxQP's factor (compact numbering)
        CCC_UNROLL
        for(int c = 0; c < NX; c++) K[c] = QUXR[r.fs.idx * NXP + c];
        llt_solve_compactN<NX>(K, A, s + sm::VB, r.fs.nf, r.invd_c); // the BoxQP column buffers are free now
      }
      // Back to the original numbering through the K staging rows of the cost-to-go update (KB aliases the factor
      // tile, which is dead from here on): compact lane r writes -K into row idx(r), a lane that is not free zeroes
      // its own row, then every lane reads its row back — instead of one 64-bit shuffle per column of K.
      {
        double * KBw = s + sm::KB;
        const bool free_i = active && !((clamped >> lane) & 1u);
        const bool owner = clamped != active_mask && lane < r.fs.nf;
        warp_sync();
        if(!free_i)
        {
          CCC_UNROLL
          for(int c = 0; c < NXP; c += 2) st2(KBw + lane * NXP + c, 0.0, 0.0);
        }
        if(owner)
        {
          double * row = KBw + r.fs.idx * NXP;
          CCC_UNROLL
          for(int c = 0; c < NXP; c += 2) st2(row + c, c < NX ? -K[c] : 0.0, c + 1 < NX ? -K[c + 1] : 0.0);
        }
        warp_sync();
        CCC_UNROLL
        for(int c = 0; c < NX; c += 2)
        {
          const d2 v = ld2(KBw + lane * NXP + c);
          K[c] = v.x;
          if(c + 1 < NX) K[c + 1] = v.y;
        }
      }
    }
    else
    {
      FreeSet fs = make_free_set(0u, m, idxbuf);
      double invd_c = 1.0;
      load_compact_row(H, S, idxbuf, fs);
      double dummy = 0.0;
      const bool ok = llt_factor_compact(H, A, s + sm::VB, fs.nf, invd_c, dummy);
      load_sym_row(H, S, m);
      if(!ok) return false;
      double r1[NX + 1];
      r1[0] = Qu;
      CCC_UNROLL
      for(int c = 0; c < NX; c++) r1[1 + c] = QUXR[lane * NXP + c];
      llt_solve_compactN<NX + 1>(r1, A, s + sm::VB, fs.nf, invd_c);
      kk = active ? -r1[0] : 0.0;
      CCC_UNROLL
      for(int c = 0; c < NX; c++) K[c] = active ? -r1[1 + c] : 0.0;
      double * KBw = s + sm::KB; // the factor tile is dead: stage K for the cost-to-go update
      CCC_UNROLL
      for(int c = 0; c < NX; c++) KBw[lane * NXP + c] = K[c];
      CCC_UNROLL
      for(int c = NX; c < NXP; c++) KBw[lane * NXP + c] = 0.0;
    }
    if(P.out_clamped && lane == 0) P.out_clamped[(size_t)b * N + k] = clamped;

    // store gains
    if(kTma)
    {
      // the K staging rows [32][NXP] (original numbering, rows >= m zero) are the HBM image of the stage: k goes into
      // slot NX of the lane's own row, then lane 0 hands rows 0..m-1 to the TMA engine as one bulk store.  The rows are
      // read asynchronously; backward_stage waits for that (bulk_wait_read_all) before the tile is written again.
      s[sm::KB + lane * NXP + NX] = kk;
      warp_sync();
      if(lane == 0)
      {
        fence_proxy_async_smem(); // the rows were written through the generic proxy
        tma_bulk_s2g(gain(k), s + sm::KB, (unsigned)m * (NXP * 8));
        bulk_commit();
      }
    }
    else
    {
      double * g = gain(k);
      g[lane] = kk;
      CCC_UNROLL
      for(int c = 0; c < NX; c++) g[(1 + c) * 32 + lane] = K[c];
    }

    // ---- cost-to-go update with the unregularised Quu ------------------------------------
    // the unregularised diagonal goes into the tile and the row is reloaded (17 instructions instead of a 32-way select)
    warp_sync();
    if(active) S[lane * kLda + lane] = quu_diag;
    warp_sync();
    load_sym_row(H, S, m);
    double * VB0 = s + sm::VB0;
    double * VB1 = s + sm::VB1;
    double * VB2 = s + sm::VB2;
    warp_sync(); // the factor in A and the BoxQP buffers are dead from here on: KB/ZB/VB0-2 alias them
    double * KB = s + sm::KB;
    double * ZB = s + sm::ZB;
    const double * QB = QUXR;
    VB0[lane] = kk; // (KB was filled when the gains were formed)
    warp_sync();
    const double Quuk = matvec32(H, VB0, m);
    double Z[NX];
    CCC_UNROLL
    for(int c = 0; c < NX; c++) Z[c] = 0.0;
    if(kRolled)
    {
      // compact-code form: the row of Quu comes from the tile (it was just reloaded from there: same values)
      const double * hrow = S + lane * kLda;
      CCC_UNROLL_N(CCC_TILE_UNROLL)
      for(int j = 0; j < m; j++)
      {
        const double h = hrow[j];
        CCC_UNROLL
        for(int c = 0; c + 1 < NX; c += 2)
        {
          const d2 kc = ld2(KB + j * NXP + c);
          Z[c] = dfma(h, kc.x, Z[c]);
          Z[c + 1] = dfma(h, kc.y, Z[c + 1]);
        }
        if(NX & 1) Z[NX - 1] = dfma(h, KB[j * NXP + NX - 1], Z[NX - 1]);
      }
    }
    else
    {
      CCC_UNROLL
      for(int j = 0; j < 32; j++)
      {
        if(j == 16 && m <= 16) break;
        const double h = H[j];
        CCC_UNROLL
        for(int c = 0; c + 1 < NX; c += 2)
        {
          const d2 kc = ld2(KB + j * NXP + c);
          Z[c] = dfma(h, kc.x, Z[c]);
          Z[c + 1] = dfma(h, kc.y, Z[c + 1]);
        }
        if(NX & 1) Z[NX - 1] = dfma(h, KB[j * NXP + NX - 1], Z[NX - 1]);
      }
    }
    {
      // k'Qu, k'Quu k (dV) and ||k||^2, ||u||^2 (small-gradient measure) through one scattered 4-way tree sum
      const double v4[4] = {active ? kk * Qu : 0.0, active ? kk * Quuk : 0.0, active ? kk * kk : 0.0, active ? u * u : 0.0};
      const double tot = warp_sum4_scattered(v4);
      const double other = warp_shfl_xor(tot, 8); // lanes 16-23: ||u||^2
      const int grp = lane >> 3;
      if(grp == 0)
        acc4 = acc4 + tot;
      else if(grp == 1)
        acc4 = dfma(0.5, tot, acc4);
      else if(grp == 2)
      {
        const double r = ddiv(dsqrt(tot), dsqrt(other) + 1.0);
        acc4 = acc4 < r ? r : acc4;
      }
    }
    CCC_UNROLL
    for(int c = 0; c < NX; c++) ZB[lane * NXP + c] = active ? Z[c] : 0.0;
    VB1[lane] = active ? Quuk : 0.0;
    VB2[lane] = active ? Qu : 0.0;
    warp_sync();
    double * Vxw = s + sm::VX;
    double * Vxxw = s + sm::VXX;
    double * S2 = s + sm::S2;
    // One pass over the inputs j feeds all the sums of this lane at once: K'(Quu k), K'Qu, Qux'k for Vx (lanes < NX)
    // and the (K'QuuK, K'Qux) pair of each of its NQ entries of Vxx — 3 + 2 NQ independent fma chains instead of
    // four loops with two or three chains each (same sums, same order within each chain).
    constexpr int NQ = (NX * NX + 31) / 32;
    double s1v[NQ], s2v[NQ];
    int ka[NQ], kc[NQ];
    CCC_UNROLL
    for(int q = 0; q < NQ; q++)
    {
      const int e = lane + 32 * q;
      const int ee = e < NX * NX ? e : 0; // lanes past the last entry repeat entry 0 and drop the result
      ka[q] = ee / NX;
      kc[q] = ee - NX * ka[q];
      s1v[q] = 0.0;
      s2v[q] = 0.0;
    }
    {
      const int lv = lane < NX ? lane : 0;
      double a1 = 0.0, a2 = 0.0, a3 = 0.0;
      CCC_NOUNROLL
      for(int j = 0; j < m; j++)
      {
        const double * kr = KB + j * NXP;
        const double * zr = ZB + j * NXP;
        const double * qr = QB + j * NXP;
        const double kj = kr[lv];
        a1 = dfma(kj, VB1[j], a1);
        a2 = dfma(kj, VB2[j], a2);
        a3 = dfma(qr[lv], VB0[j], a3);
        CCC_UNROLL
        for(int q = 0; q < NQ; q++)
        {
          const double kq = kr[ka[q]];
          s1v[q] = dfma(kq, zr[kc[q]], s1v[q]);
          s2v[q] = dfma(kq, qr[kc[q]], s2v[q]);
        }
      }
      if(lane < NX) Vxw[lane] = ((Qx[lane] + a1) + a2) + a3;
    }
    CCC_UNROLL
    for(int q = 0; q < NQ; q++)
      if(lane + 32 * q < NX * NX) S2[lane + 32 * q] = s2v[q];
    warp_sync();
    // Vn = ((Qxx + K'QuuK) + K'Qux) + Qux'K   (Qux'K = (K'Qux)' bit for bit), staged in T
    CCC_UNROLL
    for(int q = 0; q < NQ; q++)
    {
      const int e = lane + 32 * q;
      if(e < NX * NX)
      {
        const int a = e / NX, c = e - NX * a;
        T[e] = ((Qxx[e] + s1v[q]) + s2v[q]) + S2[c * NX + a];
      }
    }
    warp_sync();
    CCC_UNROLL
    for(int q = 0; q < NQ; q++)
    {
      const int e = lane + 32 * q;
      if(e < NX * NX)
      {
        const int a = e / NX, c = e - NX * a;
        Vxxw[e] = 0.5 * (T[e] + T[c * NX + a]);
      }
    }
    warp_sync();
    k_next = kk;
    m_next = m;
    return true;
  }

  CCC_DEV bool backward_pass()
  {
    const int N = P.N;
    // terminal cost derivatives (src/DdpCentroidal.cpp:156-177, src/DdpSingleRigidBody.cpp:224-245)
    const double * xN = xtraj(cur) + (size_t)N * NX;
    warp_sync();
    CCC_NOUNROLL
    for(int e = lane; e < NX * NX; e += 32)
    {
      const int i = e / NX, j = e - NX * i;
      s[sm::VXX + e] = (i == j) ? P.w_term[i] : 0.0;
    }
    if(lane < NX)
    {
      double rr = lane < NREF ? ldg(ref(N) + lane) : 0.0;
      double xv = xN[lane];
      s[sm::VX + lane] = lane < NREF ? P.w_term[lane] * (xv - rr) : P.w_term[lane] * xv;
    }
    warp_sync();
    acc4 = 0.0;
    double k_next = 0.0;
    int m_next = -1;
    bool ok = true;
    CCC_NOUNROLL
    for(int k = N - 1; k >= 0 && ok; k--) ok = backward_stage(k, k_next, m_next);
    if(kTma)
    {
      // the gain rows must be in global memory before the rollouts' bulk loads (and the next pass's warm start) read them
      if(lane == 0)
      {
        bulk_wait_all();
        fence_proxy_async_all();
      }
      warp_sync();
    }
    dV0 = warp_shfl(acc4, 0);
    dV1 = warp_shfl(acc4, 8);
    krel = warp_shfl(acc4, 16);
    return ok;
  }

  CCC_DEV void increase_lambda()
  {
    const double f = P.cfg.lambda_factor;
    double t = dlambda * f;
    dlambda = t < f ? f : t; // std::max(dlambda * f, f)
    double l = lambda * dlambda;
    lambda = l < P.cfg.lambda_min ? P.cfg.lambda_min : l;
  }
  CCC_DEV void decrease_lambda()
  {
    const double f = P.cfg.lambda_factor;
    double t = ddiv(dlambda, f), u = ddiv(1.0, f);
    dlambda = u < t ? u : t; // std::min(dlambda / f, 1 / f)
    lambda = (lambda * dlambda) * (lambda > P.cfg.lambda_min ? 1.0 : 0.0);
  }

  CCC_DEV void trace(int iter, int alpha_idx)
  {
    if(lane == 0 && iter - 1 < P.trace_len)
    {
      size_t o = (size_t)b * P.trace_len + (iter - 1);
      if(P.out_alpha_idx) P.out_alpha_idx[o] = (signed char)alpha_idx;
      if(P.out_lambda) P.out_lambda[o] = lambda;
    }
  }

  /** nmpc_ddp DDPSolver::solve + procOnce as one state machine, so that the forward pass
   *  (initial rollout and every line-search trial) and the backward pass each have a single
   *  call site in the instruction stream.  a = -1: initial rollout; a >= 0: line-search index.
   *  resume = false: start the solve; true: continue a suspended one (state in P.resume[b]).
   *  Returns true when the solve terminated (outputs written), false when it was suspended
   *  after P.chunk_iters iterations of this visit. */
  CCC_DEV bool solve(bool resume)
  {
    const int N = P.N;
    // per-visit shared-memory state: the stage-independent part of Fx and a finite (zero) tile
    warp_sync();
    M::init_Fx(*this);
    CCC_NOUNROLL
    for(int e = lane; e < sm::IDX - sm::A; e += 32) s[sm::A + e] = 0.0;
    CCC_NOUNROLL
    for(int e = lane; e < 32 * kLda; e += 32) s[sm::SYM + e] = 0.0;
    int rv = 0, iter = 0, a = -1;
    bool need_rollout = true;
    if(!resume)
    {
      lambda = P.cfg.initial_lambda;
      dlambda = P.cfg.initial_dlambda;
      cur = 0;
      // the BoxQP warm start of the last stage reads its own previous gain: zero it
      if(kTma)
      {
        gain(N - 1)[lane * NXP + NX] = 0.0;
        fence_proxy_async_all(); // ... through the generic proxy, and a bulk store will overwrite it
      }
      else
        gain(N - 1)[lane] = 0.0;
      if(P.out_clamped)
      {
        CCC_NOUNROLL
        for(int k = lane; k < N; k += 32) P.out_clamped[(size_t)b * N + k] = 0u;
      }
      CCC_NOUNROLL
      for(int i = lane; i < P.trace_len; i += 32)
      {
        size_t o = (size_t)b * P.trace_len + i;
        if(P.out_alpha_idx) P.out_alpha_idx[o] = (signed char)-4;
        if(P.out_lambda) P.out_lambda[o] = 0.0;
      }
    }
    else
    {
      const DdpResume st = P.resume[b];
      lambda = st.lambda;
      dlambda = st.dlambda;
      J = st.J;
      cur = st.cur;
      iter = st.iter;
      need_rollout = false;
    }
    const int iter_entry = iter;
    warp_sync();
    CCC_NOUNROLL
    for(;;)
    {
      if(need_rollout)
      {
        const double alpha = a >= 0 ? P.cfg.alpha[a] : 0.0;
        const double Jc = rollout(a >= 0 ? 1 - cur : 0, alpha, a < 0);
        if(a < 0)
        {
          J = Jc;
        }
        else
        {
          const double actual = J - Jc;
          const double expected = -(alpha * dfma(alpha, dV1, dV0));
          double ratio;
          if(expected > 0)
            ratio = ddiv(actual, expected);
          else
            ratio = (double)((0 < actual) - (actual < 0));
          if(ratio > P.cfg.cost_update_ratio_thre)
          {
            // accept the step
            decrease_lambda();
            cur = 1 - cur;
            J = Jc;
            if(actual < P.cfg.cost_update_thre) rv = 1;
            trace(iter, a);
          }
          else if(a + 1 < P.cfg.n_alpha)
          {
            a++;
            continue;
          }
          else
          {
            // line search failed for every alpha
            increase_lambda();
            if(lambda > P.cfg.lambda_max) rv = -1;
            trace(iter, -1);
          }
        }
      }
      need_rollout = true;
      // ---- next iteration: backward pass (with regularisation retries) ----
      if(rv != 0 || iter >= P.cfg.max_iter) break;
      if(P.chunk_iters > 0 && iter - iter_entry >= P.chunk_iters)
      {
        if(lane == 0)
        {
          DdpResume st;
          st.lambda = lambda;
          st.dlambda = dlambda;
          st.J = J;
          st.cur = cur;
          st.iter = iter;
          P.resume[b] = st;
        }
        warp_sync();
        return false;
      }
      iter++;
      bool gave_up = false;
      while(!backward_pass())
      {
        increase_lambda();
        if(lambda > P.cfg.lambda_max)
        {
          gave_up = true;
          break;
        }
      }
      if(gave_up)
      {
        trace(iter, -3);
        rv = -1;
        break;
      }
      if(krel < P.cfg.k_rel_norm_thre && lambda < P.cfg.lambda_thre)
      {
        decrease_lambda();
        trace(iter, -2);
        rv = 1;
        break;
      }
      a = 0;
    }
    // outputs
    warp_sync();
    const double * xs = xtraj(cur);
    const double * us = utraj(cur);
    if(P.out_x)
    {
      CCC_NOUNROLL
      for(int i = lane; i < (N + 1) * NX; i += 32) P.out_x[(size_t)b * (N + 1) * NX + i] = xs[i];
    }
    if(P.out_u)
    {
      CCC_NOUNROLL
      for(int k = 0; k < N; k++) P.out_u[((size_t)b * N + k) * 32 + lane] = us[(size_t)k * 32 + lane];
    }
    if(lane == 0)
    {
      if(P.out_cost) P.out_cost[b] = J;
      if(P.out_iters) P.out_iters[b] = iter;
      if(P.out_status) P.out_status[b] = rv;
    }
    return true;
  }
};
} // namespace ccc
