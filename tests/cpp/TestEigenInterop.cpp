// CCC/EigenInterop.h compiled in both of its states: with <Eigen/Core> found (here: the stand-in of tests/cpp/eigen_standin,
// -I on the command line) the conversions must round-trip the drop-in classes' vector types; without it (-DCCC_B200_NO_EIGEN)
// the header must be empty and harmless.  Uses the nested types of CCC::DdpCentroidal as a caller of the reference would.
#include "../../centroidalcontrolcollection_b200/include/CCC/DdpCentroidal.h"
#include "../../centroidalcontrolcollection_b200/include/CCC/EigenInterop.h"

#include <cstdio>
#include <cstdlib>

#define CHECK(c) \
  do \
  { \
    if(!(c)) \
    { \
      std::printf("CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
      std::exit(1); \
    } \
  } while(0)

int main()
{
#if CCC_B200_HAS_EIGEN
  // what the reference's test loop does with Eigen types (tests/src/TestDdpCentroidal.cpp:96-101), through the conversions
  Eigen::Vector3d sim_pos(0.1, -0.2, 1.0), sim_vel(0.01, 0.02, 0.03);
  CCC::DdpCentroidal::InitialParam initial_param;
  initial_param.pos = CCC::toArray<3>(sim_pos);
  initial_param.vel = CCC::toArray<3>(sim_pos + sim_vel); // an expression, not a plain vector
  CHECK(initial_param.pos[1] == -0.2 && initial_param.vel[2] == 1.03);
  Eigen::Vector3d back = CCC::toEigen(initial_param.pos);
  CHECK(back[0] == 0.1 && back[2] == 1.0 && back.size() == 3);
  // planOnce's return value and u_list
  CCC::VectorXd planned = {1.0, 2.0, 3.0, 4.0};
  Eigen::VectorXd planned_e = CCC::toEigen(planned);
  CHECK(planned_e.size() == 4 && planned_e[3] == 4.0);
  CHECK(CCC::toVector(planned_e) == planned);
  std::vector<Eigen::VectorXd> u_list_e = CCC::toEigenList({planned, {}, {5.0}});
  CHECK(u_list_e.size() == 3 && u_list_e[1].size() == 0 && u_list_e[2][0] == 5.0);
  initial_param.u_list = CCC::toVectorList(u_list_e);
  CHECK(initial_param.u_list.size() == 3 && initial_param.u_list[0] == planned);
  // a dynamic vector of the wrong length is refused
  bool thrown = false;
  try
  {
    (void)CCC::toArray<3>(planned_e);
  }
  catch(const std::invalid_argument &)
  {
    thrown = true;
  }
  CHECK(thrown);
  std::printf("EigenInterop active\n");
#else
  std::printf("EigenInterop inactive (no <Eigen/Core>)\n");
#endif
  std::printf("ALL CHECKS PASSED\n");
  return 0;
}
