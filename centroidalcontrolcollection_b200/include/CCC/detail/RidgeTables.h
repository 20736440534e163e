/* CCC/detail/RidgeTables.h — host-side sampling of contact schedules into the flat stage tables of
 * include/ccc_b200.h, shared by the drop-in classes whose inputs are ridge force scales
 * (CCC::DdpSingleRigidBody, CCC::LinearMpcXY).
 *
 * Order of the inputs inside a stage = reference src/DdpCentroidal.cpp:49-60 /
 * src/DdpSingleRigidBody.cpp:62-68: for every contact, for every vertex, for every ridge.
 */
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <vector>

#include "../../../../include/ccc_b200.h"
#include "../Contact.h"

namespace CCC
{
namespace detail
{
struct RidgeTables
{
  int S = 0, N = 0, M = CCC_DDP_M_MAX;
  std::vector<int32_t> m;            // [S][N]
  std::vector<double> ridge, vertex; // [S][N][M][3]

  void reset(int n_sched, int horizon_steps)
  {
    S = n_sched;
    N = horizon_steps;
    m.assign(static_cast<size_t>(S) * N, 0);
    ridge.assign(static_cast<size_t>(S) * N * M * 3, 0.0);
    vertex.assign(static_cast<size_t>(S) * N * M * 3, 0.0);
  }

  /** Flatten the contact list of stage k of schedule s; returns the stage's input dimension. */
  int setStage(int s, int k, const std::vector<std::shared_ptr<ForceColl::Contact>> & contact_list)
  {
    int j = 0;
    for(const auto & contact : contact_list)
      for(const auto & vr : contact->vertexWithRidgeList_)
        for(const auto & r : vr.ridgeList)
        {
          if(j >= M) throw std::runtime_error("more than CCC_DDP_M_MAX ridge inputs in one stage");
          const size_t o = ((static_cast<size_t>(s) * N + k) * M + j) * 3;
          for(int a = 0; a < 3; a++)
          {
            ridge[o + a] = r[a];
            vertex[o + a] = vr.vertex[a];
          }
          j++;
        }
    m[static_cast<size_t>(s) * N + k] = j;
    return j;
  }

  int inputDim(int s, int k) const { return m[static_cast<size_t>(s) * N + k]; }
};
} // namespace detail
} // namespace CCC
