// ddp_zmp.cu — C-ABI entry points ccc_ddp_zmp_* (include/ccc_b200.h): the generic DDP engine (ddp_host.cuh)
// with the CoM-ZMP model policy (model_zmp.cuh), unconstrained path.
#include "ddp_host.cuh"
#include "model_zmp.cuh"

#include <vector>

namespace
{
/** Derived per-stage tables on the device: m = 3, zero ridge/vertex, state references, constants row. */
__global__ void zmp_build_tables_kernel(const double * __restrict__ ref_zmp, const double * __restrict__ com_z, int S, int N,
                                        int * __restrict__ m, double * __restrict__ ridge, double * __restrict__ vertex,
                                        double * __restrict__ ref)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if(idx >= S * (N + 1)) return;
  const int s = idx / (N + 1), k = idx - s * (N + 1);
  double * r = ref + (size_t)idx * 6;
  const bool term = k == N;
  r[0] = term ? ref_zmp[(size_t)idx * 3] : 0.0;
  r[1] = 0.0;
  r[2] = term ? ref_zmp[(size_t)idx * 3 + 1] : 0.0;
  r[3] = 0.0;
  r[4] = com_z[idx];
  r[5] = 0.0;
  if(!term)
  {
    const size_t st = (size_t)s * N + k;
    m[st] = 3;
    for(int i = 0; i < 9; i++)
    {
      ridge[st * 9 + i] = 0.0;
      vertex[st * 9 + i] = 0.0;
    }
  }
}

__global__ void zmp_pack_consts_kernel(const double * __restrict__ ref_zmp, int S, int N, double mass, double * __restrict__ tab)
{
  const int st = blockIdx.x * blockDim.x + threadIdx.x;
  if(st >= S * N) return;
  const int s = st / N, k = st - s * N;
  const double * z = ref_zmp + ((size_t)s * (N + 1) + k) * 3;
  double * row = tab + ((size_t)st * ccc::ZmpModel::TAB_ROWS + 6) * 32;
  row[0] = z[0];
  row[1] = z[1];
  row[2] = mass * 9.80665;
  row[3] = z[2];
  for(int i = 4; i < 32; i++) row[i] = 0.0;
}
} // namespace

struct ccc_ddp_zmp_ws
{
  ccc_host::DdpEngine<ccc::ZmpModel> eng;
  double *d_ref_zmp = nullptr, *d_com_z = nullptr, *d_ridge = nullptr, *d_vertex = nullptr, *d_ref = nullptr;
  int * d_m = nullptr;
};

extern "C" {

ccc_ddp_zmp_ws_t * ccc_ddp_zmp_create(int32_t horizon_steps, int32_t max_batch, int32_t max_sched)
{
  if(horizon_steps <= 0 || max_batch <= 0 || max_sched <= 0)
  {
    ccc_host::set_error("ccc_ddp_zmp_create: non-positive size");
    return nullptr;
  }
  auto * ws = new ccc_ddp_zmp_ws();
  const size_t S = max_sched, N = horizon_steps;
  bool ok = ws->eng.create(horizon_steps, max_batch, max_sched);
  ok = ok && ccc_host::dev_alloc(ws->d_ref_zmp, S * (N + 1) * 3) && ccc_host::dev_alloc(ws->d_com_z, S * (N + 1));
  ok = ok && ccc_host::dev_alloc(ws->d_ridge, S * N * 9) && ccc_host::dev_alloc(ws->d_vertex, S * N * 9);
  ok = ok && ccc_host::dev_alloc(ws->d_ref, S * (N + 1) * 6) && ccc_host::dev_alloc(ws->d_m, S * N);
  if(!ok)
  {
    ccc_ddp_zmp_destroy(ws);
    return nullptr;
  }
  return ws;
}

void ccc_ddp_zmp_destroy(ccc_ddp_zmp_ws_t * ws)
{
  if(!ws) return;
  ws->eng.destroy();
  void * ptrs[] = {ws->d_ref_zmp, ws->d_com_z, ws->d_ridge, ws->d_vertex, ws->d_ref, ws->d_m};
  for(void * p : ptrs)
    if(p) cudaFree(p);
  delete ws;
}

int32_t ccc_ddp_zmp_solve(ccc_ddp_zmp_ws_t * ws,
                          const ccc_ddp_zmp_batch_t * bt,
                          const ccc_ddp_config_t * cfg,
                          ccc_ddp_result_t * res,
                          int32_t mem,
                          void * stream)
{
  if(!ws || !bt || !cfg || !res) return ccc_host::fail(CCC_ERR_INVALID, "null argument");
  const int N = bt->horizon_steps, S = bt->n_sched;
  if(N != ws->eng.N) return ccc_host::fail(CCC_ERR_INVALID, "horizon_steps differs from the workspace's");
  if(S <= 0 || S > ws->eng.max_sched) return ccc_host::fail(CCC_ERR_ALLOC, "n_sched exceeds workspace");
  if(!bt->ref_zmp || !bt->com_z) return ccc_host::fail(CCC_ERR_INVALID, "null reference table");
  if(cfg->with_input_constraint) return ccc_host::fail(CCC_ERR_INVALID, "DdpZmp has no input limits: with_input_constraint must be 0");
  ccc_host::DdpInputs<ccc::ZmpModel> in;
  in.B = bt->batch;
  in.S = S;
  in.m_max = 3;
  in.sched_id = bt->sched_id;
  in.x0 = bt->x0;
  in.u_init = bt->u_init;
  const double w_run[7] = {0, 0, 0, 0, bt->w[0], 0, 1.0};
  const double w_term[6] = {bt->w[3], bt->w[5], bt->w[3], bt->w[5], bt->w[4], bt->w[5]};
  for(int i = 0; i < 7; i++) in.w_run[i] = w_run[i];
  for(int i = 0; i < 6; i++) in.w_term[i] = w_term[i];
  in.mp.dt = bt->dt;
  in.mp.mass = bt->mass;
  in.mp.w_u[0] = bt->w[1];
  in.mp.w_u[1] = bt->w[1];
  in.mp.w_u[2] = bt->w[2];
  const bool host = mem == CCC_MEM_HOST;
  cudaStream_t own = ws->eng.own_stream;
  const double * ref_zmp_dev = bt->ref_zmp;
  // host-mode staging of the derived tables happens on the host so that the engine's own H2D path applies
  std::vector<int> h_m;
  std::vector<double> h_zero, h_ref;
  if(host)
  {
    h_m.assign((size_t)S * N, 3);
    h_zero.assign((size_t)S * N * 9, 0.0);
    h_ref.assign((size_t)S * (N + 1) * 6, 0.0);
    for(int s = 0; s < S; s++)
      for(int k = 0; k <= N; k++)
      {
        const size_t idx = (size_t)s * (N + 1) + k;
        double * r = h_ref.data() + idx * 6;
        if(k == N)
        {
          r[0] = bt->ref_zmp[idx * 3];
          r[2] = bt->ref_zmp[idx * 3 + 1];
        }
        r[4] = bt->com_z[idx];
      }
    in.m = h_m.data();
    in.ridge = h_zero.data();
    in.vertex = h_zero.data();
    in.ref = h_ref.data();
    if(!ccc_host::check(cudaMemcpyAsync(ws->d_ref_zmp, bt->ref_zmp, sizeof(double) * S * (N + 1) * 3, cudaMemcpyHostToDevice, own),
                        "H2D"))
      return CCC_ERR_CUDA;
    ref_zmp_dev = ws->d_ref_zmp;
  }
  else
  {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int total = S * (N + 1);
    zmp_build_tables_kernel<<<(total + 127) / 128, 128, 0, st>>>(bt->ref_zmp, bt->com_z, S, N, ws->d_m, ws->d_ridge, ws->d_vertex,
                                                                  ws->d_ref);
    in.m = ws->d_m;
    in.ridge = ws->d_ridge;
    in.vertex = ws->d_vertex;
    in.ref = ws->d_ref;
  }
  const double mass = bt->mass;
  const int rc = ws->eng.solve(in, cfg, res, mem, stream, [=](cudaStream_t st, double * tab) {
    zmp_pack_consts_kernel<<<(S * N + 127) / 128, 128, 0, st>>>(ref_zmp_dev, S, N, mass, tab);
    return 1;
  });
  if(!host) ws->eng.launches++;
  return rc;
}

int32_t ccc_ddp_zmp_last_launches(const ccc_ddp_zmp_ws_t * ws)
{
  return ws ? ws->eng.launches : 0;
}

} // extern "C"
