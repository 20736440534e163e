// ddp_srb.cu — C-ABI entry points ccc_ddp_srb_* (include/ccc_b200.h): the generic DDP engine
// (ddp_host.cuh) with the single-rigid-body model policy (model_srb.cuh).
#include "ddp_host.cuh"
#include "model_srb.cuh"

namespace
{
/** Row 6 of the stage table: inertia matrix and its LL^T, one thread per schedule stage. */
__global__ void srb_pack_inertia_kernel(const double * __restrict__ inertia, double * __restrict__ tab, int stages)
{
  const int st = blockIdx.x * blockDim.x + threadIdx.x;
  if(st >= stages) return;
  ccc::Inertia3 in;
  for(int i = 0; i < 9; i++) in.I[i] = inertia[(size_t)st * 9 + i];
  in.factor();
  double * row = tab + ((size_t)st * ccc::SrbModel::TAB_ROWS + 6) * 32;
  for(int i = 0; i < 9; i++) row[i] = in.I[i];
  row[9] = in.l10;
  row[10] = in.l20;
  row[11] = in.l21;
  row[12] = in.i0;
  row[13] = in.i1;
  row[14] = in.i2;
  for(int i = 15; i < 32; i++) row[i] = 0.0;
}
} // namespace

struct ccc_ddp_srb_ws
{
  ccc_host::DdpEngine<ccc::SrbModel> eng;
  double * d_inertia = nullptr;
};

extern "C" {

ccc_ddp_srb_ws_t * ccc_ddp_srb_create(int32_t horizon_steps, int32_t max_batch, int32_t max_sched)
{
  if(horizon_steps <= 0 || max_batch <= 0 || max_sched <= 0)
  {
    ccc_host::set_error("ccc_ddp_srb_create: non-positive size");
    return nullptr;
  }
  auto * ws = new ccc_ddp_srb_ws();
  if(!ws->eng.create(horizon_steps, max_batch, max_sched)
     || !ccc_host::dev_alloc(ws->d_inertia, (size_t)max_sched * horizon_steps * 9))
  {
    ccc_ddp_srb_destroy(ws);
    return nullptr;
  }
  return ws;
}

void ccc_ddp_srb_destroy(ccc_ddp_srb_ws_t * ws)
{
  if(!ws) return;
  ws->eng.destroy();
  if(ws->d_inertia) cudaFree(ws->d_inertia);
  delete ws;
}

int32_t ccc_ddp_srb_solve(ccc_ddp_srb_ws_t * ws,
                          const ccc_ddp_srb_batch_t * bt,
                          const ccc_ddp_config_t * cfg,
                          ccc_ddp_result_t * res,
                          int32_t mem,
                          void * stream)
{
  if(!ws || !bt || !cfg || !res) return ccc_host::fail(CCC_ERR_INVALID, "null argument");
  if(bt->horizon_steps != ws->eng.N) return ccc_host::fail(CCC_ERR_INVALID, "horizon_steps differs from the workspace's");
  if(!bt->inertia) return ccc_host::fail(CCC_ERR_INVALID, "null inertia table");
  if(bt->n_sched > ws->eng.max_sched || bt->n_sched <= 0) return ccc_host::fail(CCC_ERR_ALLOC, "n_sched exceeds workspace");
  ccc_host::DdpInputs<ccc::SrbModel> in;
  in.B = bt->batch;
  in.S = bt->n_sched;
  in.m_max = bt->m_max;
  in.sched_id = bt->sched_id;
  in.m = bt->m;
  in.ridge = bt->ridge;
  in.vertex = bt->vertex;
  in.ref = bt->ref;
  in.x0 = bt->x0;
  in.u_init = bt->u_init;
  for(int i = 0; i < 13; i++) in.w_run[i] = bt->w_run[i];
  for(int i = 0; i < 12; i++) in.w_term[i] = bt->w_term[i];
  in.u_lo = bt->u_lo;
  in.u_hi = bt->u_hi;
  in.mp.dt = bt->dt;
  in.mp.mass = bt->mass;
  const int stages = bt->n_sched * bt->horizon_steps;
  const double * inertia = bt->inertia;
  double * d_inertia = ws->d_inertia;
  const bool host = mem == CCC_MEM_HOST;
  cudaStream_t own = ws->eng.own_stream;
  return ws->eng.solve(in, cfg, res, mem, stream, [=](cudaStream_t st, double * tab) {
    const double * src = inertia;
    if(host)
    {
      cudaMemcpyAsync(d_inertia, inertia, sizeof(double) * stages * 9, cudaMemcpyHostToDevice, own);
      src = d_inertia;
    }
    srb_pack_inertia_kernel<<<(stages + 127) / 128, 128, 0, st>>>(src, tab, stages);
    return 1;
  });
}

int32_t ccc_ddp_srb_last_launches(const ccc_ddp_srb_ws_t * ws)
{
  return ws ? ws->eng.launches : 0;
}

} // extern "C"
