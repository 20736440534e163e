// common_host.cuh — host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

namespace ccc_host
{
inline char * error_buf()
{
  static thread_local char buf[512] = "";
  return buf;
}
inline void set_error(const char * msg)
{
  std::snprintf(error_buf(), 512, "%s", msg);
}
inline bool check(cudaError_t e, const char * what)
{
  if(e == cudaSuccess) return true;
  std::snprintf(error_buf(), 512, "%s: %s", what, cudaGetErrorString(e));
  return false;
}
/** Tuning state of the QP engine (not part of the stable ABI): packed-R two-CTA first pass on / off. */
inline bool & g_qp_packed()
{
  static bool on = true;
  return on;
}
inline int fail(int code, const char * msg)
{
  set_error(msg);
  return code;
}
} // namespace ccc_host
