/* tests/cpp/TestDdpZmp.cpp — the reference's TestDdpZmp.PlanOnce closed loop (reference
 * tests/src/TestDdpZmp.cpp:15-137) through the drop-in class CCC::DdpZmp, plus planBatch == planOnce.
 */
#include "../../centroidalcontrolcollection_b200/include/CCC/DdpZmp.h"
#include "TestFixtures.h"

using namespace fixtures;

int main()
{
  const double horizon_dt = 0.02, sim_dt = 0.005, mass = 100.0, com_height = 1.0;
  const int horizon_steps = 100;
  CCC::DdpZmp ddp(mass, horizon_dt, horizon_steps);
  ddp.ddp_solver_->config().max_iter = 3; // reference :28

  FootstepManager fm = walkingPlan();
  ComZmpSim3d sim(mass, sim_dt);
  sim.z[0] = com_height;
  auto ref_data_func = [&](double t) {
    CCC::DdpZmp::RefData rd;
    const Vec2 z = fm.refZmp(t);
    rd.zmp = {z[0], z[1], 0.0};
    rd.com_z = com_height;
    return rd;
  };

  CCC::DdpZmp::PlannedData planned;
  double t = 0;
  while(t < 10.0)
  {
    fm.update(t);
    CCC::DdpZmp::InitialParam ip;
    ip.pos = {sim.x[0], sim.y[0], sim.z[0]};
    ip.vel = {sim.x[1], sim.y[1], sim.z[1]};
    if(!ddp.ddp_solver_->controlData().u_list.empty())
      ip.u_list = ddp.ddp_solver_->controlData().u_list;
    else
      ip.u_list.assign(horizon_steps, CCC::DdpZmp::InputDimVector{sim.x[0], sim.y[0], mass * kG}); // reference :84-93
    planned = ddp.planOnce(ref_data_func, ip, t);
    const Vec2 rz = fm.refZmp(t);
    EXPECT_LT(norm(sub(planned.zmp, rz)), 0.1);
    EXPECT_LT(std::fabs(sim.z[0] - com_height), 0.1);
    t += sim_dt;
    sim.update(planned.zmp, planned.force_z);
    for(double dtm : {4.5, 8.5})
      if(dtm <= t && t < dtm + sim_dt) sim.addDisturb({0.05, 0.05});
  }
  const Vec2 rz = fm.refZmp(t);
  EXPECT_LT(norm(sub(planned.zmp, rz)), 1e-2);
  EXPECT_LT(std::fabs(sim.z[0] - com_height), 1e-2);
  EXPECT_LT(norm(sub(Vec2{sim.x[0], sim.y[0]}, rz)), 1e-2);
  EXPECT_LT(std::sqrt(sim.x[1] * sim.x[1] + sim.y[1] * sim.y[1] + sim.z[1] * sim.z[1]), 1e-2);
  std::printf("DdpZmp closed loop done: final ZMP (%.4f, %.4f), CoM height %.4f\n", planned.zmp[0], planned.zmp[1], sim.z[0]);

  {
    FootstepManager fm2 = walkingPlan();
    fm2.update(1.9);
    auto rf = [&](double tt) {
      CCC::DdpZmp::RefData rd;
      const Vec2 z = fm2.refZmp(tt);
      rd.zmp = {z[0], z[1], 0.0};
      rd.com_z = com_height;
      return rd;
    };
    CCC::DdpZmp a(mass, horizon_dt, 40), b(mass, horizon_dt, 40);
    a.config().max_iter = 5;
    b.config().max_iter = 5;
    std::vector<CCC::DdpZmp::BatchItem> items(24);
    for(int i = 0; i < 24; i++)
    {
      items[i].initial_param.pos = {0.002 * i, -0.001 * i, 1.0 + 0.001 * i};
      items[i].initial_param.vel = {0.01 * (i % 4), 0.0, 0.0};
      items[i].initial_param.u_list.assign(40, CCC::DdpZmp::InputDimVector{0.0, 0.0, mass * kG});
    }
    const auto batch = a.planBatch({rf}, items, 1.9);
    double worst = 0;
    for(int i = 0; i < 24; i += 5)
    {
      const auto one = b.planOnce(rf, items[i].initial_param, 1.9);
      worst = std::max({worst, std::fabs(one.zmp[0] - batch[i].zmp[0]), std::fabs(one.zmp[1] - batch[i].zmp[1]),
                        std::fabs(one.force_z - batch[i].force_z)});
      EXPECT_TRUE(a.lastIter(i) == b.lastIter());
    }
    EXPECT_LT(worst, 1e-300);
    std::printf("DdpZmp planBatch(24) vs planOnce: max diff %g\n", worst);
  }
  return finish("TestDdpZmp");
}
