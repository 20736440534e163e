/* tests/cpp/oracle_engine_shim.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * Implements the entry points of include/ccc_b200.h on top of the CPU oracle (oracle/libccc_oracle.so) so
 * that the host-side logic of the C++ drop-in classes (callback sampling, coefficient assembly, batching,
 * post-processing) can be exercised by the CPU test tier with the very same test programs that run against
 * libccc_b200.so on the GPU.  It is linked INSTEAD of libccc_b200.so by tests/test_cpp_dropin.py only; the
 * product library has no CPU path and nothing in the package refers to this file.
 */
#include <vector>

#include "../../include/ccc_b200.h"

extern "C" {
void ccc_oracle_ddp_config_default(ccc_ddp_config_t *);
int32_t ccc_oracle_ddp_centroidal_solve(const ccc_ddp_centroidal_batch_t *, const ccc_ddp_config_t *, ccc_ddp_result_t *, int32_t);
int32_t ccc_oracle_ddp_srb_solve(const ccc_ddp_srb_batch_t *, const ccc_ddp_config_t *, ccc_ddp_result_t *, int32_t);
int32_t ccc_oracle_ddp_zmp_solve(const ccc_ddp_zmp_batch_t *, const ccc_ddp_config_t *, ccc_ddp_result_t *, int32_t);
int32_t ccc_oracle_qp_solve(const ccc_qp_batch_t *, ccc_qp_result_t *, int32_t);
int32_t ccc_oracle_linear_mpc_xy_solve(const ccc_linear_mpc_xy_batch_t *, ccc_linear_mpc_xy_result_t *, int32_t);
int32_t ccc_oracle_ddp_centroidal_closed_loop(const ccc_ddp_centroidal_loop_t *, const ccc_ddp_config_t *, ccc_ddp_centroidal_loop_result_t *, int32_t);
int32_t ccc_oracle_preview_input(int32_t, int32_t, const double *, const double *, const double *, const double *, double *);
int32_t ccc_oracle_hardware_threads(void);

struct ccc_ddp_centroidal_ws { int dummy; };
struct ccc_ddp_srb_ws { int dummy; };
struct ccc_ddp_zmp_ws { int dummy; };
struct ccc_qp_ws
{
  std::vector<double> Q, A, C; // matrices of the last call that carried them (Q == NULL reuses them)
};

void ccc_ddp_config_default(ccc_ddp_config_t * c) { ccc_oracle_ddp_config_default(c); }
ccc_ddp_centroidal_ws_t * ccc_ddp_centroidal_create(int32_t, int32_t, int32_t) { return new ccc_ddp_centroidal_ws; }
void ccc_ddp_centroidal_destroy(ccc_ddp_centroidal_ws_t * w) { delete w; }
int32_t ccc_ddp_centroidal_solve(ccc_ddp_centroidal_ws_t *, const ccc_ddp_centroidal_batch_t * b, const ccc_ddp_config_t * c, ccc_ddp_result_t * r, int32_t, void *)
{
  return ccc_oracle_ddp_centroidal_solve(b, c, r, ccc_oracle_hardware_threads());
}
int32_t ccc_ddp_centroidal_closed_loop(ccc_ddp_centroidal_ws_t *, const ccc_ddp_centroidal_loop_t * l, const ccc_ddp_config_t * c, ccc_ddp_centroidal_loop_result_t * r, int32_t, void *)
{
  return ccc_oracle_ddp_centroidal_closed_loop(l, c, r, ccc_oracle_hardware_threads());
}
ccc_ddp_srb_ws_t * ccc_ddp_srb_create(int32_t, int32_t, int32_t) { return new ccc_ddp_srb_ws; }
void ccc_ddp_srb_destroy(ccc_ddp_srb_ws_t * w) { delete w; }
int32_t ccc_ddp_srb_solve(ccc_ddp_srb_ws_t *, const ccc_ddp_srb_batch_t * b, const ccc_ddp_config_t * c, ccc_ddp_result_t * r, int32_t, void *)
{
  return ccc_oracle_ddp_srb_solve(b, c, r, ccc_oracle_hardware_threads());
}
ccc_ddp_zmp_ws_t * ccc_ddp_zmp_create(int32_t, int32_t, int32_t) { return new ccc_ddp_zmp_ws; }
void ccc_ddp_zmp_destroy(ccc_ddp_zmp_ws_t * w) { delete w; }
int32_t ccc_ddp_zmp_solve(ccc_ddp_zmp_ws_t *, const ccc_ddp_zmp_batch_t * b, const ccc_ddp_config_t * c, ccc_ddp_result_t * r, int32_t, void *)
{
  return ccc_oracle_ddp_zmp_solve(b, c, r, ccc_oracle_hardware_threads());
}
struct ccc_linear_mpc_xy_ws { int dummy; };
ccc_linear_mpc_xy_ws_t * ccc_linear_mpc_xy_create(int32_t, int32_t, int32_t, int32_t, int32_t) { return new ccc_linear_mpc_xy_ws; }
void ccc_linear_mpc_xy_destroy(ccc_linear_mpc_xy_ws_t * w) { delete w; }
int32_t ccc_linear_mpc_xy_solve(ccc_linear_mpc_xy_ws_t *, const ccc_linear_mpc_xy_batch_t * b, ccc_linear_mpc_xy_result_t * r, int32_t, void *)
{
  return ccc_oracle_linear_mpc_xy_solve(b, r, ccc_oracle_hardware_threads());
}
ccc_qp_ws_t * ccc_qp_create(int32_t, int32_t, int32_t, int32_t) { return new ccc_qp_ws; }
void ccc_qp_destroy(ccc_qp_ws_t * w) { delete w; }
int32_t ccc_qp_solve(ccc_qp_ws_t * w, const ccc_qp_batch_t * b, ccc_qp_result_t * r, int32_t, void *)
{
  ccc_qp_batch_t bt = *b;
  const size_t n = b->n;
  if(b->Q)
  {
    w->Q.assign(b->Q, b->Q + n * n);
    w->A.assign(b->A ? b->A : b->Q, b->A ? b->A + static_cast<size_t>(b->n_eq) * n : b->Q);
    w->C.assign(b->C, b->C + static_cast<size_t>(b->n_ineq) * n);
  }
  else if(w->Q.empty())
    return CCC_ERR_INVALID;
  bt.Q = w->Q.data();
  bt.A = b->n_eq ? w->A.data() : nullptr;
  bt.C = w->C.data();
  return ccc_oracle_qp_solve(&bt, r, ccc_oracle_hardware_threads());
}
int32_t ccc_preview_input(int32_t batch, int32_t n, const double * K, const double * F, const double * x, const double * ref, double * u, int32_t, void *)
{
  return ccc_oracle_preview_input(batch, n, K, F, x, ref, u);
}
const char * ccc_last_error(void) { return "oracle shim"; }
int32_t ccc_device_count(void) { return 0; }
int32_t ccc_abi_version(void) { return CCC_B200_ABI_VERSION; }
}
