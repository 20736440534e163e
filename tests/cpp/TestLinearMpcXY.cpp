/* tests/cpp/TestLinearMpcXY.cpp — the reference's TestLinearMpcXY closed loop (reference
 * tests/src/TestLinearMpcXY.cpp:15-146) through the drop-in class CCC::LinearMpcXY (n = 240 ridge force
 * scales, 15 equalities, 480 bound rows per QP), plus planBatch == planOnce.
 */
#include "../../centroidalcontrolcollection_b200/include/CCC/LinearMpcXY.h"
#include "TestFixtures.h"

using namespace fixtures;
using Xy = CCC::LinearMpcXY;

int main(int argc, char ** argv)
{
  const double horizon_dt = 0.1, sim_dt = 0.05, mass = 100.0;
  const double end_time = argc > 1 ? std::atof(argv[1]) : 8.0;
  const int horizon_steps = static_cast<int>(1.5 / horizon_dt);
  Xy mpc(mass, horizon_dt, horizon_steps);

  auto motion_param_func = [&](double t) {
    Xy::MotionParam mp;
    mp.com_z = 1.0;
    mp.total_force_z = mass * kG;
    if(t < 3.0)
      mp.contact_list.push_back(makeContactFromRect(0.9, -0.15, 1.1, 0.15));
    else if(t < 4.0)
      mp.contact_list.push_back(makeContactFromRect(0.9, 0.05, 1.1, 0.15));
    else if(t < 5.0)
      mp.contact_list.push_back(makeContactFromRect(1.15, -0.15, 1.35, -0.05));
    else if(t < 6.0)
      mp.contact_list.push_back(makeContactFromRect(1.4, 0.05, 1.6, 0.15));
    else
      mp.contact_list.push_back(makeContactFromRect(1.4, -0.15, 1.6, 0.15));
    return mp;
  };
  auto ref_data_func = [](double t) {
    Xy::RefData rd;
    if(t < 3.0)
      rd.pos = {1.0, 0.0};
    else if(t < 4.0)
      rd.pos = {1.0, 0.1};
    else if(t < 5.0)
      rd.pos = {1.25, -0.1};
    else if(t < 6.0)
      rd.pos = {1.5, 0.1};
    else
      rd.pos = {1.5, 0.0};
    return rd;
  };

  CentroidalSim sim{mass, sim_dt, {40.0, 20.0, 10.0}};
  sim.pos = {ref_data_func(0.0).pos[0], ref_data_func(0.0).pos[1], 1.0};
  double t = 0;
  long iters = 0;
  int ticks = 0;
  while(t < end_time)
  {
    Xy::InitialParam ip;
    ip.pos = {sim.pos[0], sim.pos[1]};
    ip.vel = {sim.vel[0], sim.vel[1]};
    ip.angular_momentum = {sim.L[0], sim.L[1]};
    const Xy::VectorXd scales = mpc.planOnce(motion_param_func, ref_data_func, ip, t);
    EXPECT_TRUE(mpc.lastStatus() == 0);
    iters += mpc.lastIter();
    ticks++;
    const auto mp = motion_param_func(t);
    const auto rd = ref_data_func(t);
    Vec3 f, n;
    totalWrench(mp.contact_list, scales, sim.pos, f, n);
    EXPECT_LT(norm3(sub3(sim.pos, Vec3{rd.pos[0], rd.pos[1], mp.com_z})), 2.0);
    EXPECT_LT(norm3(sim.vel), 2.0);
    EXPECT_LT(norm3(sim.L), 5.0);
    t += sim_dt;
    sim.update(f, n);
  }
  if(end_time >= 8.0)
  {
    const auto rd = ref_data_func(t);
    EXPECT_LT(norm3(sub3(sim.pos, Vec3{rd.pos[0], rd.pos[1], 1.0})), 0.1);
    EXPECT_LT(norm3(sim.vel), 0.1);
    EXPECT_LT(norm3(sim.L), 0.1);
  }
  std::printf("LinearMpcXY closed loop: %d control cycles, mean %.1f active-set iterations per QP\n", ticks, double(iters) / ticks);

  {
    std::vector<Xy::InitialParam> ips(12);
    for(int i = 0; i < 12; i++)
    {
      ips[i].pos = {1.0 + 0.002 * i, 0.001 * i};
      ips[i].vel = {0.01 * (i % 3), -0.01 * (i % 2)};
    }
    const auto batch = mpc.planBatch(motion_param_func, ref_data_func, ips, 2.5);
    double worst = 0;
    for(int i = 0; i < 12; i += 4)
    {
      const auto one = mpc.planOnce(motion_param_func, ref_data_func, ips[i], 2.5);
      for(size_t j = 0; j < one.size(); j++) worst = std::max(worst, std::fabs(one[j] - batch[i][j]));
    }
    EXPECT_LT(worst, 1e-300);
    std::printf("LinearMpcXY planBatch(12) vs planOnce: max diff %g\n", worst);
  }
  {
    // planSweep: three schedules (the scenario sampled at three times), two initial states each == planOnce
    std::vector<Xy::Schedule> sch;
    for(double t0 : {2.5, 3.45, 5.2}) sch.push_back({motion_param_func, ref_data_func, t0});
    std::vector<Xy::InitialParam> ips(6);
    std::vector<int> sid(6);
    for(int i = 0; i < 6; i++)
    {
      sid[i] = i % 3;
      const auto rd = ref_data_func(sch[sid[i]].current_time);
      ips[i].pos = {rd.pos[0] + 0.003 * i, rd.pos[1] - 0.002 * i};
      ips[i].vel = {0.02 * (i % 2), 0.01 * i};
    }
    const auto sweep = mpc.planSweep(sch, ips, sid);
    EXPECT_TRUE(mpc.lastFailed() == 0);
    double worst = 0;
    for(int i = 0; i < 6; i++)
    {
      const auto one = mpc.planOnce(motion_param_func, ref_data_func, ips[i], sch[sid[i]].current_time);
      EXPECT_TRUE(one.size() == sweep[i].size());
      for(size_t j = 0; j < one.size(); j++) worst = std::max(worst, std::fabs(one[j] - sweep[i][j]));
    }
    EXPECT_LT(worst, 1e-300);
    std::printf("LinearMpcXY planSweep(3 schedules x 2) vs planOnce: max diff %g\n", worst);
  }
  return finish("TestLinearMpcXY");
}
