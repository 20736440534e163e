// qp_host.cuh — host-side state of the QP engine shared by qp.cu (ccc_qp_*) and linear_mpc_xy.cu
// (ccc_linear_mpc_xy_*, which feeds the engine from device-resident matrices, one group per schedule).
#pragma once
#include "common_host.cuh"
#include "qp_cta_core.cuh"

struct ccc_qp_ws
{
  int n = 0, me = 0, mi = 0, max_batch = 0, device = 0, launches = 0;
  int max_groups = 1; // matrix groups (Q, A per group; C shared)
  bool have_setup = false; // the matrices of an earlier call are factorised and resident (reused when Q == NULL)
  double *Lg = nullptr, *invd = nullptr, *J0 = nullptr, *J0s = nullptr, *At = nullptr, *Ct = nullptr;
  int *ok_flag = nullptr, *counter = nullptr; // counter[0]: first pass, [1]: fallback pass, [2]: overflow count
  int * ovf_list = nullptr;                      // problems whose active set outgrew the packed R of the first pass
  int rcap = 0;                                  // columns of the packed R (0: one CTA per SM with the full R)
  double * gmat = nullptr; // per-CTA J/R slabs when they do not fit in shared memory
  int n_sm = 148;
  // staging for CCC_MEM_HOST
  double *d_Q = nullptr, *d_A = nullptr, *d_C = nullptr, *d_c = nullptr, *d_b = nullptr, *d_d = nullptr, *d_x = nullptr;
  int *d_iters = nullptr, *d_status = nullptr, *d_nact = nullptr, *d_active = nullptr, *d_grp = nullptr;
  cudaStream_t own_stream = nullptr;
  // chunked host path: copy streams and their events
  static constexpr int kMaxChunks = 64;
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  cudaEvent_t ev_setup = nullptr, ev_h2d[kMaxChunks] = {}, ev_solved[kMaxChunks] = {};
};

namespace ccc_host
{
/** Workspace for `max_groups` matrix groups; `staging` = also the device copies of host-buffer calls. */
ccc_qp_ws * qp_ws_create(int n, int n_eq, int n_ineq, int max_batch, int max_groups, bool staging);
/** Factorise `groups` matrix groups (device pointers: Q [g][n][n], A [g][me][n], C [mi][n] shared) on stream st. */
int qp_setup_launch(ccc_qp_ws * ws, int groups, const double * Q, const double * A, const double * C, cudaStream_t st);
/** Parameter block of a solve on the workspace's resident factorisation (per-problem pointers left to the caller). */
ccc::QpParams qp_params(const ccc_qp_ws * ws, int B, const int * grp);
/** Launch the solve kernels for the problems described by P (device pointers) on stream st. */
int qp_launch(ccc_qp_ws * ws, ccc::QpParams P, cudaStream_t st);
} // namespace ccc_host
