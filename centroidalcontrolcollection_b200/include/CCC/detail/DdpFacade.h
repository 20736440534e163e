/* CCC/detail/DdpFacade.h — the members of the reference's DDP classes that its callers reach through:
 *
 *   ddp.ddp_solver_->config().max_iter / .horizon_steps     tests/src/TestDdpCentroidal.cpp:104,116
 *   ddp.ddp_solver_->controlData().u_list                   :102   (src/DdpCentroidal.cpp:236)
 *   ddp.ddp_solver_->traceDataList().back().iter            :129
 *   ddp.ddp_problem_->dt(), ddp.ddp_problem_->inputDim(t)   :106-107
 *
 * In the reference ddp_solver_ is an nmpc_ddp::DDPSolver and ddp_problem_ the class's DdpProblem; here they are
 * thin views onto the drop-in class's state (the solver lives behind the C-ABI), refreshed by every planOnce.
 */
#pragma once
#include <functional>
#include <vector>

#include "../../../../include/ccc_b200.h"

namespace CCC
{
namespace detail
{
/** nmpc_ddp::DDPSolver<>::Configuration: the C-ABI's fields plus horizon_steps (the engine takes it per batch). */
struct DdpConfiguration : public ccc_ddp_config_t
{
  int horizon_steps = 0;
};

template<class InputVector>
struct DdpSolverFacade
{
  struct ControlData
  {
    std::vector<InputVector> u_list;
  };
  struct TraceData
  {
    int iter = 0;
  };

  DdpConfiguration & config() { return config_; }
  const DdpConfiguration & config() const { return config_; }
  const ControlData & controlData() const { return control_data_; }
  const std::vector<TraceData> & traceDataList() const { return trace_data_list_; }

  DdpConfiguration config_;
  ControlData control_data_;
  std::vector<TraceData> trace_data_list_;
};

struct DdpProblemFacade
{
  double dt() const { return dt_; }
  /** Input dimension at time t: evaluates the motion-parameter callback of the last planOnce (the reference's
   *  DdpProblem::inputDim does the same, src/DdpCentroidal.cpp:21-30); a fixed dimension when there is none. */
  int inputDim(double t = 0) const { return input_dim_func_ ? input_dim_func_(t) : fixed_input_dim_; }

  double dt_ = 0;
  int fixed_input_dim_ = 0;
  std::function<int(double)> input_dim_func_;
};
} // namespace detail
} // namespace CCC
