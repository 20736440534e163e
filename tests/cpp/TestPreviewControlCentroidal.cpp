/* tests/cpp/TestPreviewControlCentroidal.cpp — the reference's TestPreviewControlCentroidal closed loop (reference
 * tests/src/TestPreviewControlCentroidal.cpp:15-150) through the drop-in class CCC::PreviewControlCentroidal, plus
 * planBatch == repeated planOnce.
 */
#include "../../centroidalcontrolcollection_b200/include/CCC/PreviewControlCentroidal.h"
#include "TestFixtures.h"

using namespace fixtures;
using Pcc = CCC::PreviewControlCentroidal;

int main()
{
  const double horizon_duration = 2.0, horizon_dt = 0.01, sim_dt = 0.005, mass = 100.0;
  const Vec3 moment_of_inertia = {40.0, 20.0, 10.0};
  Pcc pc(mass, moment_of_inertia, horizon_duration, horizon_dt);

  auto motion_param_func = [](double t) {
    t += 1e-6;
    Pcc::MotionParam mp;
    if(t < 1.4)
      mp.contact_list.push_back(makeContactFromRect(-0.1, -0.1, 0.1, 0.1));
    else if(t < 1.6)
      mp.contact_list.push_back(makeContactFromRect(0.15, 0.15, 0.35, 0.35));
    else
      mp.contact_list.push_back(makeContactFromRect(0.4, -0.1, 0.6, 0.1));
    return mp;
  };
  auto ref_data_func = [](double t) {
    t += 1e-6;
    Pcc::RefData rd;
    if(t < 1.4)
      rd.pos.linear() = {0.0, 0.0, 1.0};
    else if(t < 1.6)
      rd.pos.linear() = {0.25, 0.0, 1.2};
    else
      rd.pos.linear() = {0.5, 0.0, 1.0};
    return rd;
  };

  CentroidalSim sim{mass, sim_dt, moment_of_inertia};
  sim.pos = ref_data_func(0.0).pos.linear();
  CCC::sva::ForceVecd planned_wrench({0, 0, 0}, {0, 0, mass * kG});
  auto initial_param_of = [&](const CentroidalSim & s, const CCC::sva::ForceVecd & w) {
    Pcc::InitialParam ip;
    ip.pos = CCC::sva::MotionVecd(s.ang, s.pos);
    ip.vel = CCC::sva::MotionVecd(s.omega, s.vel);
    ip.acc.linear() = {w.force()[0] / mass, w.force()[1] / mass, w.force()[2] / mass - kG};
    ip.acc.angular() = {w.moment()[0] / moment_of_inertia[0], w.moment()[1] / moment_of_inertia[1], w.moment()[2] / moment_of_inertia[2]};
    return ip;
  };
  double t = 0;
  while(t < 3.0)
  {
    const Pcc::InitialParam ip = initial_param_of(sim, planned_wrench);
    planned_wrench = pc.planOnce(motion_param_func(t), ref_data_func, ip, t, sim_dt);
    EXPECT_TRUE(pc.lastStatus() == 0);
    const auto rd = ref_data_func(t);
    const Vec3 dl = sub3(sim.pos, rd.pos.linear());
    EXPECT_LT(std::sqrt(dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2] + sim.ang[0] * sim.ang[0] + sim.ang[1] * sim.ang[1] + sim.ang[2] * sim.ang[2]), 2.0);
    EXPECT_LT(std::sqrt(norm3(sim.vel) * norm3(sim.vel) + norm3(sim.omega) * norm3(sim.omega)), 2.0);
    t += sim_dt;
    sim.update(planned_wrench.force(), planned_wrench.moment());
    if(1.0 <= t && t < 1.0 + sim_dt)
    {
      sim.vel[0] += 0.05;
      sim.vel[1] += 0.05;
    }
  }
  {
    const auto rd = ref_data_func(t);
    const Vec3 dl = sub3(sim.pos, rd.pos.linear());
    const double perr = std::sqrt(dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2] + sim.ang[0] * sim.ang[0] + sim.ang[1] * sim.ang[1] + sim.ang[2] * sim.ang[2]);
    const double verr = std::sqrt(norm3(sim.vel) * norm3(sim.vel) + norm3(sim.omega) * norm3(sim.omega));
    EXPECT_LT(perr, 0.1);
    EXPECT_LT(verr, 0.1);
    std::printf("PreviewControlCentroidal closed loop: final position error %.3e, velocity %.3e\n", perr, verr);
  }
  {
    std::vector<Pcc::InitialParam> ips(10);
    for(int i = 0; i < 10; i++)
    {
      ips[i].pos.linear() = {0.01 * i, -0.005 * i, 1.0 + 0.002 * i};
      ips[i].pos.angular() = {0.01 * i, 0.0, -0.01 * i};
      ips[i].vel.linear() = {0.02 * (i % 3), 0.0, 0.01 * (i % 2)};
    }
    const auto batch = pc.planBatch(motion_param_func(0.5), ref_data_func, ips, 0.5, sim_dt);
    double worst = 0;
    for(int i = 0; i < 10; i += 3)
    {
      const auto one = pc.planOnce(motion_param_func(0.5), ref_data_func, ips[i], 0.5, sim_dt);
      for(int k = 0; k < 6; k++) worst = std::max(worst, std::fabs(one.vector()[k] - batch[i].vector()[k]));
    }
    EXPECT_LT(worst, 1e-300);
    std::printf("PreviewControlCentroidal planBatch(10) vs planOnce: max diff %g\n", worst);
  }
  return finish("TestPreviewControlCentroidal");
}
