/* tests/cpp/TestLinearMpcZ.cpp — the reference's TestLinearMpcZ closed loop (reference
 * tests/src/TestLinearMpcZ.cpp:12-72: jump phases without contact, planned force must be zero in flight)
 * through the drop-in class CCC::LinearMpcZ, plus planBatch == planOnce.
 */
#include "../../centroidalcontrolcollection_b200/include/CCC/LinearMpcZ.h"
#include "TestFixtures.h"

using namespace fixtures;

int main()
{
  const double horizon_duration = 2.0, horizon_dt = 0.05, sim_dt = 0.04, mass = 100.0;
  const int horizon_steps = static_cast<int>(horizon_duration / horizon_dt);
  CCC::LinearMpcZ mpc(mass, horizon_dt, horizon_steps);
  auto contact_func = [](double t) { return !((5.0 < t && t < 5.25) || (6.0 < t && t < 6.5)); };
  auto ref_pos_func = [](double t) { return t < 8.5 ? 1.0 : 0.8; };

  // VerticalSimModel (reference tests/src/SimModels.h:44-73): double integrator with gravity, exact ZOH
  std::array<double, 2> state = {ref_pos_func(0.0), 0.0};
  double t = 0;
  while(t < 10.0)
  {
    const double planned_force = mpc.planOnce(contact_func, ref_pos_func, state, t);
    const double ref_pos = ref_pos_func(t);
    EXPECT_LT(std::fabs(state[0] - ref_pos), 2.0);
    EXPECT_LT(std::fabs(state[1]), 5.0);
    if(!contact_func(t))
      EXPECT_LT(std::fabs(planned_force), 1e-8);
    else
      EXPECT_TRUE(mpc.lastStatus() == 0);
    t += sim_dt;
    const double a = planned_force / mass - kG;
    state = {state[0] + sim_dt * state[1] + 0.5 * sim_dt * sim_dt * a, state[1] + sim_dt * a};
  }
  EXPECT_LT(std::fabs(state[0] - ref_pos_func(t)), 1e-2);
  EXPECT_LT(std::fabs(state[1]), 1e-2);
  std::printf("LinearMpcZ closed loop done: final height %.4f, velocity %.4f\n", state[0], state[1]);

  {
    std::vector<CCC::LinearMpcZ::InitialParam> ips(20);
    for(int i = 0; i < 20; i++) ips[i] = {1.0 + 0.002 * i, -0.05 + 0.005 * i};
    const auto batch = mpc.planBatch(contact_func, ref_pos_func, ips, 4.6);
    double worst = 0;
    for(int i = 0; i < 20; i += 3) worst = std::max(worst, std::fabs(batch[i] - mpc.planOnce(contact_func, ref_pos_func, ips[i], 4.6)));
    EXPECT_LT(worst, 1e-300);
    std::printf("LinearMpcZ planBatch(20) vs planOnce: max diff %g\n", worst);
  }
  return finish("TestLinearMpcZ");
}
