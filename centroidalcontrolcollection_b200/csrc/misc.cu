// misc.cu — ABI version, device count, last error, nmpc_ddp default configuration.
#include "../../include/ccc_b200.h"
#include "common_host.cuh"

#include <cmath>

extern "C" {

int32_t ccc_abi_version(void)
{
  return CCC_B200_ABI_VERSION;
}

int32_t ccc_device_count(void)
{
  int n = 0;
  if(cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char * ccc_last_error(void)
{
  return ccc_host::error_buf();
}

/* nmpc_ddp::DDPSolver<>::Configuration and nmpc_ddp::BoxQP<>::Configuration defaults
 * (external dependency of the reference, values as listed in SURVEY.md App. A). */
void ccc_ddp_config_default(ccc_ddp_config_t * c)
{
  if(!c) return;
  c->with_input_constraint = 0;
  c->max_iter = 500;
  c->reg_type = 1;
  c->n_alpha = 11;
  c->initial_lambda = 1e-4;
  c->initial_dlambda = 1.0;
  c->lambda_factor = 1.6;
  c->lambda_min = 1e-6;
  c->lambda_max = 1e10;
  c->k_rel_norm_thre = 1e-4;
  c->lambda_thre = 1e-5;
  c->cost_update_ratio_thre = 0.0;
  c->cost_update_thre = 1e-7;
  for(int i = 0; i < CCC_DDP_MAX_ALPHA; i++) c->alpha[i] = i < 11 ? std::pow(10.0, -3.0 * i / 10.0) : 0.0;
  c->boxqp_max_iter = 500;
  c->reserved0 = 0;
  c->boxqp_grad_thre = 1e-8;
  c->boxqp_rel_improve_thre = 1e-8;
  c->boxqp_step_factor = 0.6;
  c->boxqp_min_step = 1e-22;
  c->boxqp_armijo = 0.1;
}

} // extern "C"
