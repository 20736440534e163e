// tests/emu/warp_emu.cpp — lock-step warp emulator + C entry point that runs the *kernel
// source* (centroidalcontrolcollection_b200/csrc/*_core.cuh, compiled with CCC_WARP_EMU) on the CPU.
//
// TEST INFRASTRUCTURE ONLY.  Purpose: check the CUDA solver core bit for bit against the
// oracle in the CPU-only test tier (no GPU in the build container), and catch warp-sync
// mistakes (a lane that skips a collective dead-locks the emulator).  It is far too slow to be
// a fallback and is never linked into libccc_b200.so.
//
// 32 lanes = 32 ucontext fibres on one OS thread, scheduled round-robin; every warp
// collective (__syncwarp, shuffles, ballot) is a barrier over the live fibres.
#define CCC_WARP_EMU 1
#include "../../include/ccc_b200.h"
#include "../../centroidalcontrolcollection_b200/csrc/ddp_team.cuh"
#include "../../centroidalcontrolcollection_b200/csrc/model_centroidal.cuh"
#include "../../centroidalcontrolcollection_b200/csrc/model_srb.cuh"
#include "../../centroidalcontrolcollection_b200/csrc/model_zmp.cuh"
#include "../../centroidalcontrolcollection_b200/csrc/ddp_thread_zmp.cuh"
#include "../../centroidalcontrolcollection_b200/csrc/qp_cta_core.cuh"
#include "../../centroidalcontrolcollection_b200/csrc/preview_core.cuh"

#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <type_traits>
#include <vector>

namespace ccc_emu
{
namespace
{
constexpr int kMaxThreads = 256;
constexpr int kCtaGroup = kMaxThreads / 32; // barrier groups: 0..7 = the warps, 8 = the whole CTA
constexpr size_t kStack = 1 << 19;
ucontext_t g_main, g_ctx[kMaxThreads];
std::vector<char> g_stacks;
bool g_done[kMaxThreads];
int g_cur = 0, g_nthreads = 32, g_alive = 0;
int g_arrived[kCtaGroup + 1] = {};
unsigned long g_gen[kCtaGroup + 1] = {};
int g_wait_group[kMaxThreads];          // barrier group a fibre is parked at, or -1
unsigned long g_wait_gen[kMaxThreads];  // ... and the generation it waits to see pass
double g_xd[kMaxThreads];
int g_xi[kMaxThreads];
std::function<void()> g_body;

void yield()
{
  int from = g_cur;
  int nxt = from;
  for(int s = 1; s <= g_nthreads; s++)
  {
    int c = (from + s) % g_nthreads;
    // a fibre parked at a barrier that has not moved is not worth a context switch (the seven helper warps of a team
    // wait at the CTA barrier for the whole backward pass of the leader)
    if(!g_done[c] && !(g_wait_group[c] >= 0 && g_gen[g_wait_group[c]] == g_wait_gen[c]))
    {
      nxt = c;
      break;
    }
  }
  if(nxt == from) return;
  g_cur = nxt;
  swapcontext(&g_ctx[from], &g_ctx[nxt]);
}

void barrier(int group, int size)
{
  unsigned long gen = g_gen[group];
  if(++g_arrived[group] == size)
  {
    g_arrived[group] = 0;
    g_gen[group]++;
  }
  else
  {
    g_wait_group[g_cur] = group;
    g_wait_gen[g_cur] = gen;
    while(g_gen[group] == gen) yield();
    g_wait_group[g_cur] = -1;
  }
}

int warp_size_of(int w)
{
  int lo = w * 32, hi = lo + 32 < g_nthreads ? lo + 32 : g_nthreads;
  return hi - lo;
}

void trampoline()
{
  g_body();
  g_done[g_cur] = true;
  g_alive--;
  if(g_arrived[kCtaGroup] > 0 && g_alive > 0)
  {
    // a thread exited while others wait at a CTA barrier: divergence bug in the kernel
    std::fprintf(stderr, "ccc_emu: thread %d exited while %d threads wait at __syncthreads\n", g_cur, g_arrived[kCtaGroup]);
    std::abort();
  }
  for(int c = 0; c < g_nthreads; c++)
    if(!g_done[c])
    {
      int from = g_cur;
      g_cur = c;
      swapcontext(&g_ctx[from], &g_ctx[c]);
    }
  setcontext(&g_main);
}
} // namespace

int tid() { return g_cur; }
int lane() { return g_cur & 31; }
void syncwarp() { barrier(g_cur >> 5, warp_size_of(g_cur >> 5)); }
void syncthreads() { barrier(kCtaGroup, g_nthreads); }
double shfl(double v, int src)
{
  const int w = g_cur >> 5;
  g_xd[g_cur] = v;
  barrier(w, warp_size_of(w));
  double r = g_xd[w * 32 + (src & 31)];
  barrier(w, warp_size_of(w));
  return r;
}
double shfl_xor(double v, int mask) { return shfl(v, (g_cur & 31) ^ mask); }
int shfl_i(int v, int src)
{
  const int w = g_cur >> 5;
  g_xi[g_cur] = v;
  barrier(w, warp_size_of(w));
  int r = g_xi[w * 32 + (src & 31)];
  barrier(w, warp_size_of(w));
  return r;
}
unsigned ballot(bool p)
{
  const int w = g_cur >> 5;
  g_xi[g_cur] = p ? 1 : 0;
  barrier(w, warp_size_of(w));
  unsigned r = 0;
  for(int i = 0; i < warp_size_of(w); i++)
    if(g_xi[w * 32 + i]) r |= (1u << i);
  barrier(w, warp_size_of(w));
  return r;
}

/** Run `body` once on each of `nthreads` (<= 256) threads of one CTA in lock step. */
void run_cta(int nthreads, const std::function<void()> & body)
{
  if(g_stacks.empty()) g_stacks.resize(kStack * kMaxThreads);
  g_body = body;
  g_nthreads = nthreads;
  g_alive = nthreads;
  for(int g = 0; g <= kCtaGroup; g++) g_arrived[g] = 0;
  for(int i = 0; i < nthreads; i++)
  {
    g_done[i] = false;
    g_wait_group[i] = -1;
    getcontext(&g_ctx[i]);
    g_ctx[i].uc_stack.ss_sp = g_stacks.data() + kStack * i;
    g_ctx[i].uc_stack.ss_size = kStack;
    g_ctx[i].uc_link = &g_main;
    makecontext(&g_ctx[i], trampoline, 0);
  }
  g_cur = 0;
  swapcontext(&g_main, &g_ctx[0]);
}

void run_warp(const std::function<void()> & body)
{
  run_cta(32, body);
}
} // namespace ccc_emu

namespace
{
ccc::DdpCfg toCfg(const ccc_ddp_config_t * c)
{
  ccc::DdpCfg d;
  d.with_input_constraint = c->with_input_constraint;
  d.max_iter = c->max_iter;
  d.n_alpha = c->n_alpha;
  d.initial_lambda = c->initial_lambda;
  d.initial_dlambda = c->initial_dlambda;
  d.lambda_factor = c->lambda_factor;
  d.lambda_min = c->lambda_min;
  d.lambda_max = c->lambda_max;
  d.k_rel_norm_thre = c->k_rel_norm_thre;
  d.lambda_thre = c->lambda_thre;
  d.cost_update_ratio_thre = c->cost_update_ratio_thre;
  d.cost_update_thre = c->cost_update_thre;
  for(int i = 0; i < 16; i++) d.alpha[i] = c->alpha[i];
  d.boxqp.max_iter = c->boxqp_max_iter;
  d.boxqp.grad_thre = c->boxqp_grad_thre;
  d.boxqp.rel_improve_thre = c->boxqp_rel_improve_thre;
  d.boxqp.step_factor = c->boxqp_step_factor;
  d.boxqp.min_step = c->boxqp_min_step;
  d.boxqp.armijo = c->boxqp_armijo;
  return d;
}
} // namespace

namespace
{
int g_chunk = 0;
int g_feat = 1;
int g_team = 0;
constexpr int kTeam = 8; // ccc_host::kTeam (ddp_host.cuh)
}

/** 1: emulate the small-batch kernel (ddp_team.cuh: a CTA of 8 warps per problem, concurrent line-search rollouts)
 *  instead of the warp-per-problem kernel. */
extern "C" void ccc_emu_set_team(int32_t team)
{
  g_team = team;
}

/** Which build of the solver core to emulate (ddp_warp_core.cuh kFeat*): 0 = no feature bits, 1 = the product default,
 *  2 = every feature (incl. the TMA-staged gain lists, whose bulk copies the emulator performs synchronously). */
extern "C" void ccc_emu_set_feat(int32_t feat)
{
  g_feat = feat;
}

/** Iterations per visit before a solve is suspended and re-queued (0 = never), as in the kernel. */
extern "C" void ccc_emu_set_chunk(int32_t chunk)
{
  g_chunk = chunk;
}

/** Host-side replay of ccc_host::DdpEngine<M>::solve for one emulated warp. */
template<class M, class FillExtra>
int32_t emuSolve(int N, int B, int S, int mm, const int32_t * sched_id, const int32_t * m, const double * ridge,
                 const double * vertex, const double * ref, const double * x0, const double * u_init_in,
                 const double * w_run, const double * w_term, double u_lo, double u_hi, const typename M::Params & mp,
                 const ccc_ddp_config_t * c, ccc_ddp_result_t * r, FillExtra && fill_extra)
{
  constexpr int NX = M::NX;
  using sm = ccc::SmLayout<M::NX, M::NXP>;
  if(mm > 32) return CCC_ERR_INVALID;
  // pack stage tables [S][N][TAB_ROWS][32] exactly as the engine's pack kernels do
  std::vector<double> tab((size_t)S * N * 32 * M::TAB_ROWS, 0.0);
  for(int s = 0; s < S; s++)
    for(int k = 0; k < N; k++)
      for(int j = 0; j < mm; j++)
        for(int a = 0; a < 3; a++)
        {
          size_t src = (((size_t)s * N + k) * mm + j) * 3 + a;
          tab[((size_t)s * N + k) * 32 * M::TAB_ROWS + a * 32 + j] = ridge[src];
          tab[((size_t)s * N + k) * 32 * M::TAB_ROWS + (3 + a) * 32 + j] = vertex[src];
        }
  fill_extra(tab.data());
  std::vector<double> u_init;
  if(u_init_in)
  {
    u_init.assign((size_t)B * N * 32, 0.0);
    for(size_t bk = 0; bk < (size_t)B * N; bk++)
      for(int j = 0; j < mm; j++) u_init[bk * 32 + j] = u_init_in[bk * mm + j];
  }
  const size_t ntraj = g_team ? kTeam + 1 : 2;
  std::vector<double> xbuf(ntraj * B * (N + 1) * NX, 0.0), ubuf(ntraj * B * N * 32, 0.0),
      gains((size_t)B * N * 32 * M::NXP, 0.0), out_u((size_t)B * N * 32, 0.0);
  ccc::DdpParams<M> P{};
  P.N = N;
  P.B = B;
  P.S = S;
  P.tab_len = N;
  P.ref_len = N + 1;
  P.tab_off = 0;
  P.tab_stride = 1;
  P.sched_id = sched_id;
  P.m = m;
  P.tab = tab.data();
  P.ref = ref;
  for(int i = 0; i <= NX; i++) P.w_run[i] = w_run[i];
  for(int i = 0; i < NX; i++) P.w_term[i] = w_term[i];
  P.mp = mp;
  P.u_lo = u_lo;
  P.u_hi = u_hi;
  P.x0 = x0;
  P.u_init = u_init_in ? u_init.data() : nullptr;
  P.cfg = toCfg(c);
  P.xbuf = xbuf.data();
  P.ubuf = ubuf.data();
  P.gains = gains.data();
  P.out_x = r->x;
  P.out_u = r->u ? out_u.data() : nullptr;
  P.out_cost = r->cost;
  P.out_iters = r->iters;
  P.out_status = r->status;
  P.trace_len = r->trace_len;
  P.out_alpha_idx = reinterpret_cast<signed char *>(r->alpha_idx);
  P.out_lambda = r->lambda_trace;
  P.out_clamped = r->clamped;
  std::vector<ccc::DdpResume> resume(B);
  P.resume = resume.data();
  P.chunk_iters = g_chunk;
  P.abort_ok = c->cost_update_ratio_thre >= 0.0 ? 1 : 0;
  for(int i = 0; i <= NX; i++)
    if(!(w_run[i] >= 0.0)) P.abort_ok = 0;
  for(int i = 0; i < NX; i++)
    if(!(w_term[i] >= 0.0)) P.abort_ok = 0;
  if(g_team)
  {
    // ddp_team_kernel: one CTA of kTeam warps per problem, run to completion
    P.chunk_iters = 0;
    std::vector<double> tsm((size_t)kTeam * sm::TOTAL + ccc::team_ctl_doubles<kTeam>() + 2, 0.0);
    for(int b = 0; b < B; b++)
      ccc_emu::run_cta(kTeam * 32, [&]() {
        if(P.cfg.with_input_constraint)
          ccc::team_solve<M, true, ccc::kFeatAbort, kTeam>(P, tsm.data(), b);
        else
          ccc::team_solve<M, false, ccc::kFeatAbort, kTeam>(P, tsm.data(), b);
      });
    if(r->u)
      for(size_t bk = 0; bk < (size_t)B * N; bk++)
        for(int j = 0; j < mm; j++) r->u[bk * mm + j] = out_u[bk * 32 + j];
    return CCC_OK;
  }
  std::vector<double> smem(sm::TOTAL + 2, 0.0);
  unsigned ring_parity = 0;
  // the kernel's round-robin queue, replayed by one emulated warp: (problem, resumed?) entries
  std::deque<std::pair<int, bool>> queue;
  for(int b = 0; b < B; b++) queue.emplace_back(b, false);
  while(!queue.empty())
  {
    const int b = queue.front().first;
    const bool resumed = queue.front().second;
    queue.pop_front();
    bool finished = true;
    ccc_emu::run_warp([&]() {
      bool f;
      auto run = [&](auto tag_c, auto tag_f) {
        ccc::DdpWarp<M, decltype(tag_c)::value, decltype(tag_f)::value> w(P, smem.data(), b, &ring_parity);
        return w.solve(resumed);
      };
      using T = std::true_type;
      using F = std::false_type;
      using FeatOn = std::integral_constant<int, ccc::kFeatDefault>;
      using FeatAll = std::integral_constant<int, ccc::kFeatAll>;
      using FeatOff = std::integral_constant<int, 0>;
      if(P.cfg.with_input_constraint)
        f = g_feat == 2 ? run(T{}, FeatAll{}) : g_feat ? run(T{}, FeatOn{}) : run(T{}, FeatOff{});
      else
        f = g_feat == 2 ? run(F{}, FeatAll{}) : g_feat ? run(F{}, FeatOn{}) : run(F{}, FeatOff{});
      if(ccc_emu::lane() == 0) finished = f;
    });
    if(!finished) queue.emplace_back(b, true);
  }
  if(r->u)
    for(size_t bk = 0; bk < (size_t)B * N; bk++)
      for(int j = 0; j < mm; j++) r->u[bk * mm + j] = out_u[bk * 32 + j];
  return CCC_OK;
}

extern "C" int32_t ccc_emu_ddp_centroidal_solve(const ccc_ddp_centroidal_batch_t * bt,
                                                const ccc_ddp_config_t * c,
                                                ccc_ddp_result_t * r)
{
  ccc::CentroidalModel::Params mp;
  mp.dt = bt->dt;
  mp.mass = bt->mass;
  return emuSolve<ccc::CentroidalModel>(bt->horizon_steps, bt->batch, bt->n_sched, bt->m_max, bt->sched_id, bt->m,
                                        bt->ridge, bt->vertex, bt->ref_pos, bt->x0, bt->u_init, bt->w_run, bt->w_term,
                                        bt->u_lo, bt->u_hi, mp, c, r, [](double *) {});
}

extern "C" int32_t ccc_emu_ddp_srb_solve(const ccc_ddp_srb_batch_t * bt, const ccc_ddp_config_t * c, ccc_ddp_result_t * r)
{
  ccc::SrbModel::Params mp;
  mp.dt = bt->dt;
  mp.mass = bt->mass;
  const int stages = bt->n_sched * bt->horizon_steps;
  return emuSolve<ccc::SrbModel>(bt->horizon_steps, bt->batch, bt->n_sched, bt->m_max, bt->sched_id, bt->m, bt->ridge,
                                 bt->vertex, bt->ref, bt->x0, bt->u_init, bt->w_run, bt->w_term, bt->u_lo, bt->u_hi, mp, c, r,
                                 [&](double * tab) {
                                   // srb_pack_inertia_kernel, on the host
                                   for(int st = 0; st < stages; st++)
                                   {
                                     ccc::Inertia3 in;
                                     for(int i = 0; i < 9; i++) in.I[i] = bt->inertia[(size_t)st * 9 + i];
                                     in.factor();
                                     double * row = tab + ((size_t)st * ccc::SrbModel::TAB_ROWS + 6) * 32;
                                     for(int i = 0; i < 9; i++) row[i] = in.I[i];
                                     row[9] = in.l10;
                                     row[10] = in.l20;
                                     row[11] = in.l21;
                                     row[12] = in.i0;
                                     row[13] = in.i1;
                                     row[14] = in.i2;
                                   }
                                 });
}

namespace
{
int g_qp_rcap = 0, g_qp_last_overflows = -1;
}

/** Columns of the packed R of the first pass (0: full R only, the round-1 kernel). */
extern "C" void ccc_emu_qp_set_rcap(int32_t rcap)
{
  g_qp_rcap = rcap;
}

/** Problems of the last ccc_emu_qp_solve that outgrew the packed R and went through the full-R pass (-1: no packed pass). */
extern "C" int32_t ccc_emu_qp_last_overflows(void)
{
  return g_qp_last_overflows;
}

extern "C" int32_t ccc_emu_qp_solve(const ccc_qp_batch_t * bt, ccc_qp_result_t * r)
{
  const int n = bt->n, me = bt->n_eq, mi = bt->n_ineq, B = bt->batch, ld = n | 1;
  if(n > 256 || me + mi > (n > 128 ? 1024 : 512)) return CCC_ERR_INVALID;
  std::vector<double> Lg((size_t)n * n, 0.0), invd(n), J0((size_t)n * n), At((size_t)n * (me ? me : 1)), Ct((size_t)n * mi);
  int ok_flag_buf[4] = {0, 0, 0, 0};
  std::vector<double> bcast(256, 0.0);
  ccc_emu::run_cta(ccc::kQpThreads, [&]() {
    ccc::qp_setup_cta(n, me, mi, bt->Q, bt->A, bt->C, Lg.data(), invd.data(), J0.data(), At.data(), Ct.data(), ok_flag_buf, nullptr, true,
                      bcast.data());
  });
  ccc::QpParams P{};
  P.n = n;
  P.me = me;
  P.mi = mi;
  P.B = B;
  P.ld = ld;
  P.J0 = J0.data();
  P.At = At.data();
  P.Ct = Ct.data();
  P.c = bt->c;
  P.b = bt->b;
  P.d = bt->d;
  P.setup_ok = ok_flag_buf;
  P.max_iter = 1000;
  P.viol_tol = 1e-10;
  P.out_x = r->x;
  P.out_iters = r->iters;
  P.out_status = r->status;
  P.out_n_active = r->n_active;
  P.out_active = r->active;
  // same shapes as qp.cu: 128 threads with J/R in "shared" memory up to n = 116, 128 threads with the
  // global slab up to n = 128, 256 threads with the global slab beyond
  std::vector<double> gmat((size_t)2 * n * ld, 0.0);
  if(n > 128)
  {
    std::vector<double> smem(ccc::QpSm<256, true>::bytes(n, ld) / sizeof(double) + 2, 0.0);
    for(int b = 0; b < B; b++)
      ccc_emu::run_cta(256, [&]() {
        ccc::QpCta<256, true> cta(P, smem.data(), b, gmat.data());
        cta.solve();
      });
  }
  else if(ccc::QpSm<128, false>::bytes(n, ld) > 227 * 1024)
  {
    std::vector<double> smem(ccc::QpSm<128, true>::bytes(n, ld) / sizeof(double) + 2, 0.0);
    for(int b = 0; b < B; b++)
      ccc_emu::run_cta(128, [&]() {
        ccc::QpCta<128, true> cta(P, smem.data(), b, gmat.data());
        cta.solve();
      });
  }
  else
  {
    // as qp.cu: first pass with the packed R (g_qp_rcap columns; the engine sizes it so that two CTAs fit on an SM),
    // then the problems that outgrew it once more with the full R
    std::vector<int> ovf_list(B, -1);
    int ovf_count = 0;
    const bool packed = g_qp_rcap > me;
    if(packed)
    {
      ccc::QpParams P1 = P;
      P1.rcap = g_qp_rcap < n ? g_qp_rcap : n;
      P1.ovf_count = &ovf_count;
      P1.ovf_list = ovf_list.data();
      std::vector<double> smem(ccc::QpSm<128, false, true>::bytes(n, ld, P1.rcap) / sizeof(double) + 2, 0.0);
      for(int b = 0; b < B; b++)
        ccc_emu::run_cta(128, [&]() {
          ccc::QpCta<128, false, true> cta(P1, smem.data(), b);
          cta.solve();
        });
    }
    std::vector<double> smem(ccc::QpSm<128, false>::bytes(n, ld) / sizeof(double) + 2, 0.0);
    for(int t = 0; t < (packed ? ovf_count : B); t++)
    {
      const int b = packed ? ovf_list[t] : t;
      ccc_emu::run_cta(128, [&]() {
        ccc::QpCta<128, false> cta(P, smem.data(), b);
        cta.solve();
      });
    }
    g_qp_last_overflows = packed ? ovf_count : -1;
  }
  return CCC_OK;
}

extern "C" int32_t ccc_emu_preview_input(int32_t B, int32_t N, const double * K, const double * F, const double * x,
                                         const double * ref_seq, double * u)
{
  for(int b = 0; b < B; b++)
    ccc_emu::run_warp([&]() {
      const double v = ccc::preview_row(N, K, F, x + (size_t)b * 3, ref_seq + (size_t)b * N);
      if(ccc_emu::lane() == 0) u[b] = v;
    });
  return CCC_OK;
}

extern "C" int32_t ccc_emu_ddp_zmp_solve(const ccc_ddp_zmp_batch_t * bt, const ccc_ddp_config_t * c, ccc_ddp_result_t * r)
{
  const int N = bt->horizon_steps, S = bt->n_sched;
  ccc::ZmpModel::Params mp;
  mp.dt = bt->dt;
  mp.mass = bt->mass;
  mp.w_u[0] = bt->w[1];
  mp.w_u[1] = bt->w[1];
  mp.w_u[2] = bt->w[2];
  const double w_run[7] = {0, 0, 0, 0, bt->w[0], 0, 1.0};
  const double w_term[6] = {bt->w[3], bt->w[5], bt->w[3], bt->w[5], bt->w[4], bt->w[5]};
  std::vector<int32_t> m((size_t)S * N, 3);
  std::vector<double> zero((size_t)S * N * 9, 0.0), ref((size_t)S * (N + 1) * 6, 0.0);
  for(int s = 0; s < S; s++)
    for(int k = 0; k <= N; k++)
    {
      const size_t idx = (size_t)s * (N + 1) + k;
      if(k == N)
      {
        ref[idx * 6 + 0] = bt->ref_zmp[idx * 3];
        ref[idx * 6 + 2] = bt->ref_zmp[idx * 3 + 1];
      }
      ref[idx * 6 + 4] = bt->com_z[idx];
    }
  return emuSolve<ccc::ZmpModel>(N, bt->batch, S, 3, bt->sched_id, m.data(), zero.data(), zero.data(), ref.data(), bt->x0,
                                 bt->u_init, w_run, w_term, 0.0, 0.0, mp, c, r, [&](double * tab) {
                                   for(int st = 0; st < S * N; st++)
                                   {
                                     const int s = st / N, k = st - s * N;
                                     const double * z = bt->ref_zmp + ((size_t)s * (N + 1) + k) * 3;
                                     double * row = tab + ((size_t)st * ccc::ZmpModel::TAB_ROWS + 6) * 32;
                                     row[0] = z[0];
                                     row[1] = z[1];
                                     row[2] = bt->mass * 9.80665;
                                     row[3] = z[2];
                                   }
                                 });
}

/** The thread-per-problem DdpZmp solver (csrc/ddp_thread_zmp.cuh) compiled for the host: the same source the
 *  zmp_thread_kernel runs, with the engine's strided per-problem storage (stride = batch). */
extern "C" int32_t ccc_emu_ddp_zmp_thread_solve(const ccc_ddp_zmp_batch_t * bt, const ccc_ddp_config_t * c, ccc_ddp_result_t * r)
{
  const int N = bt->horizon_steps, B = bt->batch;
  std::vector<double> slab(ccc_thread::zmp_work_doubles(N) * (size_t)B, 0.0);
  for(int b = 0; b < B; b++) ccc_thread::zmp_thread_run(*bt, *c, *r, ccc_thread::zmp_work_at(slab.data(), N, (size_t)B, (size_t)b), b);
  return CCC_OK;
}
