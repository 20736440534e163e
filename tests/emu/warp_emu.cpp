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
#include "../../centroidalcontrolcollection_b200/csrc/model_centroidal.cuh"
#include "../../centroidalcontrolcollection_b200/csrc/model_srb.cuh"

#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <vector>

namespace ccc_emu
{
namespace
{
constexpr int kLanes = 32;
constexpr size_t kStack = 1 << 20;
ucontext_t g_main, g_ctx[kLanes];
std::vector<char> g_stacks;
bool g_done[kLanes];
int g_cur = 0, g_alive = 0, g_arrived = 0;
unsigned long g_gen = 0;
double g_xd[kLanes];
int g_xi[kLanes];
std::function<void()> g_body;

void yield()
{
  int from = g_cur;
  int nxt = from;
  for(int s = 1; s <= kLanes; s++)
  {
    int c = (from + s) % kLanes;
    if(!g_done[c])
    {
      nxt = c;
      break;
    }
  }
  if(nxt == from) return;
  g_cur = nxt;
  swapcontext(&g_ctx[from], &g_ctx[nxt]);
}

void barrier()
{
  unsigned long gen = g_gen;
  if(++g_arrived == g_alive)
  {
    g_arrived = 0;
    g_gen++;
  }
  else
  {
    while(g_gen == gen) yield();
  }
}

void trampoline()
{
  g_body();
  g_done[g_cur] = true;
  g_alive--;
  if(g_arrived == g_alive && g_alive > 0 && g_arrived > 0)
  {
    // a lane exited while the others wait at a collective: divergence bug in the kernel
    std::fprintf(stderr, "ccc_emu: lane %d exited while %d lanes wait at a warp collective\n", g_cur, g_arrived);
    std::abort();
  }
  // switch to any live fibre, or back to main
  for(int c = 0; c < kLanes; c++)
    if(!g_done[c])
    {
      int from = g_cur;
      g_cur = c;
      swapcontext(&g_ctx[from], &g_ctx[c]);
    }
  setcontext(&g_main);
}
} // namespace

int lane() { return g_cur; }
void syncwarp() { barrier(); }
double shfl(double v, int src)
{
  g_xd[g_cur] = v;
  barrier();
  double r = g_xd[src & 31];
  barrier();
  return r;
}
double shfl_xor(double v, int mask) { return shfl(v, g_cur ^ mask); }
int shfl_i(int v, int src)
{
  g_xi[g_cur] = v;
  barrier();
  int r = g_xi[src & 31];
  barrier();
  return r;
}
unsigned ballot(bool p)
{
  g_xi[g_cur] = p ? 1 : 0;
  barrier();
  unsigned r = 0;
  for(int i = 0; i < kLanes; i++)
    if(g_xi[i]) r |= (1u << i);
  barrier();
  return r;
}

/** Run `body` once on each of the 32 lanes in lock step. */
void run_warp(const std::function<void()> & body)
{
  if(g_stacks.empty()) g_stacks.resize(kStack * kLanes);
  g_body = body;
  g_alive = kLanes;
  g_arrived = 0;
  for(int i = 0; i < kLanes; i++)
  {
    g_done[i] = false;
    getcontext(&g_ctx[i]);
    g_ctx[i].uc_stack.ss_sp = g_stacks.data() + kStack * i;
    g_ctx[i].uc_stack.ss_size = kStack;
    g_ctx[i].uc_link = &g_main;
    makecontext(&g_ctx[i], trampoline, 0);
  }
  g_cur = 0;
  swapcontext(&g_main, &g_ctx[0]);
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
  std::vector<double> xbuf((size_t)2 * B * (N + 1) * NX, 0.0), ubuf((size_t)2 * B * N * 32, 0.0),
      gains((size_t)B * N * 32 * (1 + NX), 0.0), out_u((size_t)B * N * 32, 0.0);
  ccc::DdpParams<M> P;
  std::memset(&P, 0, sizeof(P));
  P.N = N;
  P.B = B;
  P.S = S;
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
  std::vector<double> smem(sm::TOTAL + 2, 0.0);
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
      if(P.cfg.with_input_constraint)
      {
        ccc::DdpWarp<M, true> w(P, smem.data(), b);
        f = w.solve(resumed);
      }
      else
      {
        ccc::DdpWarp<M, false> w(P, smem.data(), b);
        f = w.solve(resumed);
      }
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
