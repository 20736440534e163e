// ddp_team.cuh — small-batch (latency) mapping of the DDP solve: one CTA of TEAM warps per problem.
//
// The throughput kernel (ddp_host.cuh ddp_solve_kernel) gives every problem one warp and fills the SM with eight of
// them.  A caller with a handful of problems — the reference's own use: one planOnce per control tick
// (tests/src/TestDdpCentroidal.cpp:94-150) — leaves the chip empty and waits for one warp's serial recursion.  Here a
// problem gets a whole SM: warp 0 (the leader) runs what is serial by nature — the initial rollout, the backward
// recursion with its BoxQP per stage, the lambda schedule — and the line search, nmpc_ddp's loop over the step sizes
// alpha[0], alpha[1], ... until one is accepted, runs as ONE round of concurrent rollouts: warp w rolls out candidate
// alpha[base + w] into its own trajectory buffer, the acceptance test of every candidate is evaluated exactly as the
// serial loop would, and the candidate with the smallest index that passes is taken.  The serial loop stops at the first
// accepted index and never looks at the later ones, so taking the smallest accepted index of the round is the same
// decision, and each rollout is the same instruction sequence on the same inputs: results are bit-identical to the
// warp-per-problem kernel and to oracle/ddp.hpp (tests/test_emu_team.py on the CPU, tests/test_gpu_ddp_team.py on B200).
//
// Trajectory buffers: TEAM + 1 per problem (P.xbuf / P.ubuf indexed by `which` in 0..TEAM); the nominal trajectory is
// buffer `cur`, candidate w goes to buffer (cur + 1 + w) mod (TEAM + 1), and accepting a candidate renames its buffer.
#pragma once
#include "ddp_warp_core.cuh"

namespace ccc
{
/** Control block of a team in shared memory (behind the warps' slices). */
template<int TEAM>
struct TeamCtl
{
  double J, dV0, dV1;
  double Jc[TEAM];
  int ok[TEAM];
  int cur, phase; // phase 1: line search follows, 0: the solve has terminated
};

template<int TEAM>
constexpr int team_ctl_doubles()
{
  return (int)((sizeof(TeamCtl<TEAM>) + 7) / 8);
}

/** nmpc_ddp DDPSolver::solve for problem b by the TEAM warps of this CTA (every thread of the CTA calls it).
 *  Same termination logic and trace as DdpWarp::solve; no suspension (a team runs its problem to the end). */
template<class M, bool kConstrained, int FEAT, int TEAM>
CCC_DEV void team_solve(const DdpParams<M> & P, double * smem, int b)
{
  static_assert((FEAT & kFeatTma) == 0, "the team mapping uses the register-prefetch rollouts");
  using Warp = DdpWarp<M, kConstrained, FEAT>;
  using sm = typename Warp::sm;
  const int wid = thread_id() >> 5;
  const int lane = lane_id();
  const bool leader = wid == 0;
  double * s = smem + wid * sm::TOTAL;
  TeamCtl<TEAM> * ctl = reinterpret_cast<TeamCtl<TEAM> *>(smem + TEAM * sm::TOTAL);
  unsigned ring_parity = 0;
  Warp w(P, s, b, &ring_parity);
  const int N = P.N;

  // per-warp shared-memory state, as DdpWarp::solve sets it up (the helpers only use the reduction scratch)
  warp_sync();
  M::init_Fx(w);
  CCC_NOUNROLL
  for(int e = lane; e < sm::IDX - sm::A; e += 32) s[sm::A + e] = 0.0;
  CCC_NOUNROLL
  for(int e = lane; e < 32 * kLda; e += 32) s[sm::SYM + e] = 0.0;
  warp_sync();

  int rv = 0, iter = 0;
  if(leader)
  {
    w.lambda = P.cfg.initial_lambda;
    w.dlambda = P.cfg.initial_dlambda;
    w.cur = 0;
    w.gain(N - 1)[lane] = 0.0; // BoxQP warm start of the last stage reads its own previous gain
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
    warp_sync();
    w.J = w.rollout(0, 0.0, true);
  }

  CCC_NOUNROLL
  for(;;)
  {
    // ---- leader: termination tests and the backward pass (with regularisation retries) ----
    if(leader)
    {
      int phase = 1;
      if(rv != 0 || iter >= P.cfg.max_iter)
        phase = 0;
      else
      {
        iter++;
        bool gave_up = false;
        while(!w.backward_pass())
        {
          w.increase_lambda();
          if(w.lambda > P.cfg.lambda_max)
          {
            gave_up = true;
            break;
          }
        }
        if(gave_up)
        {
          w.trace(iter, -3);
          rv = -1;
          phase = 0;
        }
        else if(w.krel < P.cfg.k_rel_norm_thre && w.lambda < P.cfg.lambda_thre)
        {
          w.decrease_lambda();
          w.trace(iter, -2);
          rv = 1;
          phase = 0;
        }
      }
      if(lane == 0)
      {
        ctl->phase = phase;
        ctl->J = w.J;
        ctl->dV0 = w.dV0;
        ctl->dV1 = w.dV1;
        ctl->cur = w.cur;
      }
    }
    cta_sync(); // the control block, the gain lists and the nominal trajectory are visible to the team
    if(ctl->phase == 0) break;

    // ---- line search: rounds of TEAM concurrent candidates ----
    const double J = ctl->J, dV0 = ctl->dV0, dV1 = ctl->dV1;
    const int cur = ctl->cur;
    w.J = J; // (the early abort of a hopeless rollout compares against the nominal cost)
    w.cur = cur;
    int acc = -1, acc_slot = 0;
    CCC_NOUNROLL
    for(int base = 0; base < P.cfg.n_alpha; base += TEAM)
    {
      const int a = base + wid;
      int ok = 0;
      double Jc = 0.0;
      if(a < P.cfg.n_alpha)
      {
        const double alpha = P.cfg.alpha[a];
        Jc = w.rollout((cur + 1 + wid) % (TEAM + 1), alpha, false);
        const double actual = J - Jc;
        const double expected = -(alpha * dfma(alpha, dV1, dV0));
        double ratio;
        if(expected > 0)
          ratio = ddiv(actual, expected);
        else
          ratio = (double)((0 < actual) - (actual < 0));
        ok = ratio > P.cfg.cost_update_ratio_thre ? 1 : 0;
      }
      if(lane == 0)
      {
        ctl->Jc[wid] = Jc;
        ctl->ok[wid] = ok;
      }
      cta_sync(); // verdicts and candidate trajectories of the round are visible
      CCC_NOUNROLL
      for(int i = TEAM - 1; i >= 0; i--)
        if(ctl->ok[i])
        {
          acc = base + i;
          acc_slot = i;
        }
      if(acc >= 0) break;
      cta_sync(); // everyone has read the verdicts before the next round overwrites them
    }
    if(leader)
    {
      if(acc >= 0)
      {
        const double Jc = ctl->Jc[acc_slot];
        const double actual = J - Jc;
        w.decrease_lambda();
        w.cur = (cur + 1 + acc_slot) % (TEAM + 1);
        w.J = Jc;
        if(actual < P.cfg.cost_update_thre) rv = 1;
        w.trace(iter, acc);
      }
      else
      {
        w.increase_lambda();
        if(w.lambda > P.cfg.lambda_max) rv = -1;
        w.trace(iter, -1);
      }
    }
    // no barrier here: the helpers wait at the barrier after the next backward pass, and the leader rewrites J / cur /
    // phase only after it has read Jc above; ok[] / Jc[] are rewritten after that barrier
  }

  // outputs (the leader's registers hold the final state)
  if(leader)
  {
    const double * xs = w.xtraj(w.cur);
    const double * us = w.utraj(w.cur);
    if(P.out_x)
    {
      CCC_NOUNROLL
      for(int i = lane; i < (N + 1) * M::NX; i += 32) P.out_x[(size_t)b * (N + 1) * M::NX + i] = xs[i];
    }
    if(P.out_u)
    {
      CCC_NOUNROLL
      for(int k = 0; k < N; k++) P.out_u[((size_t)b * N + k) * 32 + lane] = us[(size_t)k * 32 + lane];
    }
    if(lane == 0)
    {
      if(P.out_cost) P.out_cost[b] = w.J;
      if(P.out_iters) P.out_iters[b] = iter;
      if(P.out_status) P.out_status[b] = rv;
    }
  }
}
} // namespace ccc
