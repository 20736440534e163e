// ddp_centroidal.cu — C-ABI entry points ccc_ddp_centroidal_* (include/ccc_b200.h) on top of the
// generic DDP engine (ddp_host.cuh) with the centroidal model policy (model_centroidal.cuh).
#include "ddp_host.cuh"
#include "model_centroidal.cuh"

struct ccc_ddp_centroidal_ws
{
  ccc_host::DdpEngine<ccc::CentroidalModel> eng;
};

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

int32_t ccc_ddp_centroidal_last_launches(const ccc_ddp_centroidal_ws_t * ws)
{
  return ws ? ws->eng.launches : 0;
}

/* Tuning hooks shared by the DDP engines (not part of the stable ABI). */
int32_t ccc_ddp_centroidal_set_variant(int32_t v)
{
  if(v >= 0 && v < 3) ccc_host::g_variant() = v;
  return 3;
}

void ccc_ddp_centroidal_set_chunk(int32_t chunk)
{
  ccc_host::g_chunk() = chunk < 0 ? 0 : chunk;
}

} // extern "C"
