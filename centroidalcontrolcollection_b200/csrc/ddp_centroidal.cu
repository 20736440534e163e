// ddp_centroidal.cu — C-ABI entry points ccc_ddp_centroidal_* (include/ccc_b200.h) on top of the
// generic DDP engine (ddp_host.cuh) with the centroidal model policy (model_centroidal.cuh).
#include "ddp_host.cuh"
#include "model_centroidal.cuh"

/** Device buffers of ccc_ddp_centroidal_closed_loop, grown on demand and owned by the workspace. */
struct LoopBuffers
{
  double *tab = nullptr, *ref = nullptr, *ridge = nullptr, *vertex = nullptr, *plant = nullptr, *plant0 = nullptr;
  double *plant_log = nullptr, *u0_log = nullptr;
  int *m = nullptr, *iters_log = nullptr;
  size_t cap_grid = 0, cap_log = 0; // S * grid_len entries, B * (ticks + 1) entries
  void release()
  {
    void * ptrs[] = {tab, ref, ridge, vertex, plant, plant0, plant_log, u0_log, m, iters_log};
    for(void * p : ptrs)
      if(p) cudaFree(p);
    *this = LoopBuffers();
  }
};

struct ccc_ddp_centroidal_ws
{
  ccc_host::DdpEngine<ccc::CentroidalModel> eng;
  LoopBuffers loop;
};

namespace
{
/** plant <- plant0, x0 = (pos, mass * vel, L), log entry 0. */
__global__ void loop_init_kernel(int B, int ticks, double mass, const double * __restrict__ plant0, double * __restrict__ plant,
                                 double * __restrict__ x0, double * __restrict__ plant_log)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= B) return;
  for(int i = 0; i < 9; i++)
  {
    const double v = plant0[(size_t)b * 9 + i];
    plant[(size_t)b * 9 + i] = v;
    plant_log[(size_t)b * (ticks + 1) * 9 + i] = v;
    x0[(size_t)b * 9 + i] = (i >= 3 && i < 6) ? mass * v : v;
  }
}

/** Warm start of cycle `tick` (reference tests/src/TestDdpCentroidal.cpp:102-114): the previous plan, stage by
 *  stage (not shifted), zeroed where the stage's input dimension differs from the previous cycle's. */
__global__ void loop_warm_start_kernel(int B, int N, int tick, int stride, int grid_len, const int * __restrict__ sched_id,
                                       const int * __restrict__ m, const double * __restrict__ u_prev, double * __restrict__ u_init)
{
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(idx >= (size_t)B * N * 32) return;
  const int j = (int)(idx & 31);
  const size_t bk = idx >> 5;
  const int k = (int)(bk % N), b = (int)(bk / N);
  const size_t e = (size_t)sched_id[b] * grid_len + tick + (size_t)k * stride;
  const int m_now = m[e], m_prev = m[e - 1];
  u_init[idx] = (m_now == m_prev && j < m_now) ? u_prev[idx] : 0.0;
}

/** One control cycle of the plant (CentroidalSim, reference tests/src/SimModels.h:233-332: the plant's A is nilpotent,
 *  so its zero-order hold is the polynomial below): total wrench of the first stage's force scales about the CoM
 *  (ForceColl::calcTotalWrench), then pos += vel dt + acc dt^2 / 2, vel += acc dt, L += n dt.  One thread per plant;
 *  operation order = oracle/capi.cpp ccc_oracle_ddp_centroidal_closed_loop. */
__global__ void loop_plant_kernel(int B, int N, int tick, int ticks, int grid_len, int m_max, double mass, double sim_dt, int disturb,
                                  double dvx, double dvy, double dvz, const int * __restrict__ sched_id, const int * __restrict__ m,
                                  const double * __restrict__ tab, const double * __restrict__ u, const int * __restrict__ iters,
                                  double * __restrict__ plant, double * __restrict__ x0, double * __restrict__ plant_log,
                                  double * __restrict__ u0_log, int * __restrict__ iters_log)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= B) return;
  const size_t e = (size_t)sched_id[b] * grid_len + tick;
  const int mk = m[e];
  const double * tb = tab + e * (32 * ccc::CentroidalModel::TAB_ROWS);
  double * st = plant + (size_t)b * 9;
  double pos[3] = {st[0], st[1], st[2]}, vel[3] = {st[3], st[4], st[5]}, L[3] = {st[6], st[7], st[8]};
  const double * u0 = u + (size_t)b * N * 32;
  double f[3] = {0.0, 0.0, 0.0}, n[3] = {0.0, 0.0, 0.0};
  for(int j = 0; j < mk; j++)
  {
    const double uj = u0[j];
    double rho[3], d[3], cr[3];
    for(int a = 0; a < 3; a++)
    {
      rho[a] = tb[a * 32 + j];
      d[a] = tb[(3 + a) * 32 + j] - pos[a];
    }
    ccc::cross3(d, rho, cr);
    for(int a = 0; a < 3; a++)
    {
      f[a] = ccc::dfma(uj, rho[a], f[a]);
      n[a] = ccc::dfma(uj, cr[a], n[a]);
    }
  }
  const double half_dt2 = 0.5 * (sim_dt * sim_dt);
  const double dv[3] = {dvx, dvy, dvz};
  for(int a = 0; a < 3; a++)
  {
    const double acc = f[a] / mass + (a == 2 ? -9.80665 : 0.0);
    pos[a] = ccc::dfma(half_dt2, acc, ccc::dfma(sim_dt, vel[a], pos[a]));
    vel[a] = ccc::dfma(sim_dt, acc, vel[a]);
    L[a] = ccc::dfma(sim_dt, n[a], L[a]);
    if(disturb) vel[a] = vel[a] + dv[a];
  }
  double * lg = plant_log + ((size_t)b * (ticks + 1) + tick + 1) * 9;
  for(int a = 0; a < 3; a++)
  {
    st[a] = lg[a] = pos[a];
    st[3 + a] = lg[3 + a] = vel[a];
    st[6 + a] = lg[6 + a] = L[a];
    x0[(size_t)b * 9 + a] = pos[a];
    x0[(size_t)b * 9 + 3 + a] = mass * vel[a];
    x0[(size_t)b * 9 + 6 + a] = L[a];
  }
  if(u0_log)
    for(int j = 0; j < m_max; j++) u0_log[((size_t)b * ticks + tick) * m_max + j] = j < mk ? u0[j] : 0.0;
  if(iters_log) iters_log[(size_t)b * ticks + tick] = iters[b];
}
} // namespace

extern "C" {

ccc_ddp_centroidal_ws_t * ccc_ddp_centroidal_create(int32_t horizon_steps, int32_t max_batch, int32_t max_sched)
{
  if(horizon_steps <= 0 || max_batch <= 0 || max_sched <= 0)
  {
    ccc_host::set_error("ccc_ddp_centroidal_create: non-positive size");
    return nullptr;
  }
  auto * ws = new ccc_ddp_centroidal_ws();
  if(!ws->eng.create(horizon_steps, max_batch, max_sched))
  {
    ws->eng.destroy();
    delete ws;
    return nullptr;
  }
  return ws;
}

void ccc_ddp_centroidal_destroy(ccc_ddp_centroidal_ws_t * ws)
{
  if(!ws) return;
  ws->loop.release();
  ws->eng.destroy();
  delete ws;
}

int32_t ccc_ddp_centroidal_solve(ccc_ddp_centroidal_ws_t * ws,
                                 const ccc_ddp_centroidal_batch_t * bt,
                                 const ccc_ddp_config_t * cfg,
                                 ccc_ddp_result_t * res,
                                 int32_t mem,
                                 void * stream)
{
  if(!ws || !bt || !cfg || !res) return ccc_host::fail(CCC_ERR_INVALID, "null argument");
  if(bt->horizon_steps != ws->eng.N) return ccc_host::fail(CCC_ERR_INVALID, "horizon_steps differs from the workspace's");
  ccc_host::DdpInputs<ccc::CentroidalModel> in;
  in.B = bt->batch;
  in.S = bt->n_sched;
  in.m_max = bt->m_max;
  in.sched_id = bt->sched_id;
  in.m = bt->m;
  in.ridge = bt->ridge;
  in.vertex = bt->vertex;
  in.ref = bt->ref_pos;
  in.x0 = bt->x0;
  in.u_init = bt->u_init;
  for(int i = 0; i < 10; i++) in.w_run[i] = bt->w_run[i];
  for(int i = 0; i < 9; i++) in.w_term[i] = bt->w_term[i];
  in.u_lo = bt->u_lo;
  in.u_hi = bt->u_hi;
  in.mp.dt = bt->dt;
  in.mp.mass = bt->mass;
  return ws->eng.solve(in, cfg, res, mem, stream, [](cudaStream_t, double *) { return 0; });
}

int32_t ccc_ddp_centroidal_closed_loop(ccc_ddp_centroidal_ws_t * ws,
                                       const ccc_ddp_centroidal_loop_t * lp,
                                       const ccc_ddp_config_t * cfg,
                                       ccc_ddp_centroidal_loop_result_t * res,
                                       int32_t mem,
                                       void * stream_v)
{
  using ccc_host::check;
  using ccc_host::dev_alloc;
  using ccc_host::fail;
  if(!ws || !lp || !cfg || !res) return fail(CCC_ERR_INVALID, "null argument");
  auto & eng = ws->eng;
  auto & lb = ws->loop;
  const int N = lp->horizon_steps, B = lp->batch, S = lp->n_sched, mm = lp->m_max, T = lp->ticks, G = lp->grid_len;
  if(N != eng.N) return fail(CCC_ERR_INVALID, "horizon_steps differs from the workspace's");
  if(B <= 0 || S <= 0 || B > eng.max_batch || S > eng.max_sched) return fail(CCC_ERR_ALLOC, "batch or n_sched exceeds workspace");
  if(mm <= 0 || mm > CCC_DDP_M_MAX || T <= 0 || lp->stride <= 0) return fail(CCC_ERR_INVALID, "bad m_max / ticks / stride");
  if((long long)G < (long long)T - 1 + (long long)N * lp->stride + 1) return fail(CCC_ERR_INVALID, "time grid shorter than the last horizon");
  if(!lp->sched_id || !lp->m || !lp->ridge || !lp->vertex || !lp->ref_pos || !lp->plant0 || !res->plant)
    return fail(CCC_ERR_INVALID, "null table");
  if(cfg->reg_type != 1 || cfg->n_alpha < 1 || cfg->n_alpha > CCC_DDP_MAX_ALPHA) return fail(CCC_ERR_INVALID, "bad solver configuration");
  cudaStream_t st = mem == CCC_MEM_HOST ? eng.own_stream : reinterpret_cast<cudaStream_t>(stream_v);
  eng.launches = 0;

  // buffers
  const size_t grid_entries = (size_t)S * G, log_entries = (size_t)B * (T + 1);
  if(grid_entries > lb.cap_grid || log_entries > lb.cap_log)
  {
    lb.release();
    bool ok = dev_alloc(lb.tab, grid_entries * 32 * ccc::CentroidalModel::TAB_ROWS) && dev_alloc(lb.ref, grid_entries * 3)
              && dev_alloc(lb.ridge, grid_entries * 32 * 3) && dev_alloc(lb.vertex, grid_entries * 32 * 3) && dev_alloc(lb.m, grid_entries)
              && dev_alloc(lb.plant, (size_t)eng.max_batch * 9) && dev_alloc(lb.plant0, (size_t)eng.max_batch * 9)
              && dev_alloc(lb.plant_log, log_entries * 9) && dev_alloc(lb.u0_log, log_entries * 32) && dev_alloc(lb.iters_log, log_entries);
    if(!ok)
    {
      lb.release();
      return CCC_ERR_CUDA;
    }
    lb.cap_grid = grid_entries;
    lb.cap_log = log_entries;
  }
  const int * d_sched = lp->sched_id;
  const int * d_m = lp->m;
  const double *d_ridge = lp->ridge, *d_vertex = lp->vertex, *d_ref = lp->ref_pos, *d_plant0 = lp->plant0;
  double *o_plant = res->plant, *o_u0 = res->u0;
  int * o_iters = res->iters;
  if(mem == CCC_MEM_HOST)
  {
    for(size_t i = 0; i < grid_entries; i++)
      if(lp->m[i] < 0 || lp->m[i] > mm) return fail(CCC_ERR_INVALID, "stage input dimension outside [0, m_max]");
    for(int i = 0; i < B; i++)
      if(lp->sched_id[i] < 0 || lp->sched_id[i] >= S) return fail(CCC_ERR_INVALID, "sched_id out of range");
#define CCC_H2D(dst, src, n) \
  if(!check(cudaMemcpyAsync(dst, src, (n), cudaMemcpyHostToDevice, st), "H2D")) return CCC_ERR_CUDA
    CCC_H2D(eng.d_sched_id, lp->sched_id, sizeof(int) * B);
    CCC_H2D(lb.m, lp->m, sizeof(int) * grid_entries);
    CCC_H2D(lb.ridge, lp->ridge, sizeof(double) * grid_entries * mm * 3);
    CCC_H2D(lb.vertex, lp->vertex, sizeof(double) * grid_entries * mm * 3);
    CCC_H2D(lb.ref, lp->ref_pos, sizeof(double) * grid_entries * 3);
    CCC_H2D(lb.plant0, lp->plant0, sizeof(double) * B * 9);
#undef CCC_H2D
    d_sched = eng.d_sched_id;
    d_m = lb.m;
    d_ridge = lb.ridge;
    d_vertex = lb.vertex;
    d_ref = lb.ref;
    d_plant0 = lb.plant0;
    o_plant = lb.plant_log;
    o_u0 = res->u0 ? lb.u0_log : nullptr;
    o_iters = res->iters ? lb.iters_log : nullptr;
  }
  {
    const size_t total = grid_entries * 192;
    ccc_host::pack_tables_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_ridge, d_vertex, lb.tab, (size_t)grid_entries, mm,
                                                                                   ccc::CentroidalModel::TAB_ROWS);
    eng.launches++;
  }
  loop_init_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, T, lp->mass, d_plant0, lb.plant, eng.d_x0, o_plant);
  eng.launches++;

  ccc_host::DdpInputs<ccc::CentroidalModel> in;
  in.B = B;
  in.S = S;
  in.m_max = 32;
  in.sched_id = d_sched;
  in.m = d_m;
  in.ref = d_ref;
  in.x0 = eng.d_x0;
  for(int i = 0; i < 10; i++) in.w_run[i] = lp->w_run[i];
  for(int i = 0; i < 9; i++) in.w_term[i] = lp->w_term[i];
  in.u_lo = lp->u_lo;
  in.u_hi = lp->u_hi;
  in.mp.dt = lp->dt;
  in.mp.mass = lp->mass;
  ccc_ddp_config_t cfg_tick = *cfg;
  for(int tick = 0; tick < T; tick++)
  {
    if(tick == 1) cfg_tick.max_iter = lp->max_iter_later;
    ccc::DdpParams<ccc::CentroidalModel> P = eng.make_params(in, &cfg_tick);
    P.tab = lb.tab;
    P.tab_len = G;
    P.ref_len = G;
    P.tab_off = tick;
    P.tab_stride = lp->stride;
    P.out_u = eng.uo32;
    P.out_iters = eng.d_iters;
    if(tick > 0)
    {
      const size_t total = (size_t)B * N * 32;
      loop_warm_start_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(B, N, tick, lp->stride, G, d_sched, d_m, eng.uo32, eng.u32);
      eng.launches++;
      P.u_init = eng.u32;
    }
    const int rc = eng.launch_solve(P, &cfg_tick, st);
    if(rc != CCC_OK) return rc;
    loop_plant_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, N, tick, T, G, mm, lp->mass, lp->sim_dt, tick == lp->disturb_tick ? 1 : 0,
                                                       lp->disturb_vel[0], lp->disturb_vel[1], lp->disturb_vel[2], d_sched, d_m, lb.tab,
                                                       eng.uo32, eng.d_iters, lb.plant, eng.d_x0, o_plant, o_u0, o_iters);
    eng.launches++;
  }
  if(!check(cudaGetLastError(), "closed loop launches")) return CCC_ERR_CUDA;
  if(mem == CCC_MEM_HOST)
  {
    if(!check(cudaMemcpyAsync(res->plant, lb.plant_log, sizeof(double) * log_entries * 9, cudaMemcpyDeviceToHost, st), "D2H")) return CCC_ERR_CUDA;
    if(res->u0 && !check(cudaMemcpyAsync(res->u0, lb.u0_log, sizeof(double) * (size_t)B * T * mm, cudaMemcpyDeviceToHost, st), "D2H"))
      return CCC_ERR_CUDA;
    if(res->iters && !check(cudaMemcpyAsync(res->iters, lb.iters_log, sizeof(int) * (size_t)B * T, cudaMemcpyDeviceToHost, st), "D2H"))
      return CCC_ERR_CUDA;
    if(!check(cudaStreamSynchronize(st), "cudaStreamSynchronize")) return CCC_ERR_CUDA;
  }
  return CCC_OK;
}

int32_t ccc_ddp_centroidal_last_launches(const ccc_ddp_centroidal_ws_t * ws)
{
  return ws ? ws->eng.launches : 0;
}

/* Tuning hooks shared by the DDP engines (not part of the stable ABI). */
int32_t ccc_ddp_centroidal_set_variant(int32_t v)
{
  int n = 0;
  ccc_host::Variants<ccc::CentroidalModel>::table(n);
  if(v >= 0 && v < n) ccc_host::g_variant() = v;
  return n;
}

void ccc_ddp_centroidal_set_chunk(int32_t chunk)
{
  ccc_host::g_chunk() = chunk < 0 ? 0 : chunk;
}

/** Small-batch policy of every DDP engine (A/B measurements and tests): team = 1 runs batches of at most one problem per
 *  SM on the team kernel (ddp_team.cuh: one CTA per problem, concurrent line-search rollouts); spread = 1 spreads batches
 *  smaller than the resident warps over all SMs.  Negative values leave a setting unchanged.  Both default to 1; results
 *  do not depend on either. */
void ccc_ddp_set_small_batch_policy(int32_t team, int32_t spread)
{
  if(team >= 0) ccc_host::g_team() = team ? 1 : 0;
  if(spread >= 0) ccc_host::g_spread() = spread ? 1 : 0;
}

/** Tuning hook: 0 = host-buffer calls always use one copy per array (A/B of the packed staging block of small calls). */
void ccc_ddp_set_packed_io(int32_t on)
{
  ccc_host::g_packed_io() = on ? 1 : 0;
}

/** 1 if the workspace's last solve ran on the team kernel. */
int32_t ccc_ddp_centroidal_last_team(const ccc_ddp_centroidal_ws_t * ws)
{
  return ws ? ws->eng.last_team : 0;
}

} // extern "C"
