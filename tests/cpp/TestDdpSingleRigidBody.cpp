/* tests/cpp/TestDdpSingleRigidBody.cpp — the reference's TestDdpSingleRigidBody.PlanOnce closed loop
 * (reference tests/src/TestDdpSingleRigidBody.cpp:15-175) through the drop-in class CCC::DdpSingleRigidBody,
 * plus planBatch == planOnce.  One DDP iteration per control cycle after the first, as the reference (:133); the
 * sensitivity of this scenario to rounding is measured in profiles/r02_srb_robustness.txt.
 */
#include "../../centroidalcontrolcollection_b200/include/CCC/DdpSingleRigidBody.h"
#include "TestFixtures.h"

using namespace fixtures;
using Srb = CCC::DdpSingleRigidBody;

int main()
{
  const double horizon_dt = 0.03, sim_dt = 0.005, mass = 100.0;
  const int horizon_steps = 100;
  Srb::WeightParam wp;
  wp.running_pos = {1.0, 1.0, 10.0};
  wp.running_ori = {0.5, 0.5, 0.5};
  wp.terminal_pos = {1.0, 1.0, 10.0};
  wp.terminal_ori = {0.5, 0.5, 0.5};
  Srb ddp(mass, horizon_dt, horizon_steps, wp);

  auto motion_param_func = [](double t) {
    t += 1e-6;
    Srb::MotionParam mp;
    if(t < 1.4)
      mp.contact_list.push_back(makeContactFromRect(-0.1, -0.5, 0.1, 0.5));
    else if(t >= 1.6)
      mp.contact_list.push_back(makeContactFromRect(0.4, -0.5, 0.6, 0.5));
    mp.inertia_mat = {40.0, 0, 0, 0, 20.0, 0, 0, 0, 10.0};
    return mp;
  };
  auto ref_data_func = [](double t) {
    t += 1e-6;
    Srb::RefData rd;
    if(t < 1.4)
      rd.pos = {0.0, 0.0, 1.0};
    else if(t < 1.6)
      rd.pos = {0.25, 0.0, 1.2};
    else
      rd.pos = {0.5, 0.0, 1.0};
    if(2.2 < t && t < 2.4) rd.ori = {0.0, 0.0, 0.3};
    return rd;
  };

  CentroidalSim sim{mass, sim_dt, {40.0, 20.0, 10.0}};
  sim.pos = ref_data_func(0.0).pos;
  double t = 0;
  int first_iter = 0;
  while(t < 3.0)
  {
    Srb::InitialParam ip;
    ip.pos = sim.pos;
    ip.ori = {sim.ang[2], sim.ang[1], sim.ang[0]}; // the plant keeps (X, Y, Z), the controller (Z, Y, X) (reference :92)
    ip.linear_vel = sim.vel;
    ip.angular_vel = sim.omega;
    ip.u_list = ddp.ddp_solver_->controlData().u_list; // reference :118-130 through the ddp_solver_ / ddp_problem_ views
    if(!ip.u_list.empty())
    {
      for(int i = 0; i < horizon_steps; i++)
      {
        const int input_dim = ddp.ddp_problem_->inputDim(t + i * ddp.ddp_problem_->dt());
        if(static_cast<int>(ip.u_list[i].size()) != input_dim) ip.u_list[i].assign(input_dim, 0.0);
      }
    }
    else
      first_iter = -1;
    const Srb::VectorXd scales = ddp.planOnce(motion_param_func, ref_data_func, ip, t);
    if(first_iter == -1) first_iter = ddp.ddp_solver_->traceDataList().back().iter; // :146
    ddp.ddp_solver_->config().max_iter = 1; // Set max_iter from second simulation iteration (:132)

    const auto mp = motion_param_func(t);
    const auto rd = ref_data_func(t);
    Vec3 f, n;
    totalWrench(mp.contact_list, scales, sim.pos, f, n);
    EXPECT_LT(norm3(sub3(sim.pos, rd.pos)), 2.0);
    EXPECT_LT(norm3(sub3(sim.ang, Vec3{rd.ori[2], rd.ori[1], rd.ori[0]})), 1.0);
    EXPECT_LT(norm3(sim.vel), 2.0);
    EXPECT_LT(norm3(sim.omega), 2.0);
    t += sim_dt;
    sim.update(f, n);
    if(1.0 <= t && t < 1.0 + sim_dt)
    {
      sim.vel[0] += 0.05;
      sim.vel[1] += 0.05;
    }
  }
  const auto rd = ref_data_func(t);
  EXPECT_LT(norm3(sub3(sim.pos, rd.pos)), 0.1);
  EXPECT_LT(norm3(sub3(sim.ang, Vec3{rd.ori[2], rd.ori[1], rd.ori[0]})), 0.1);
  EXPECT_LT(norm3(sim.vel), 0.1);
  EXPECT_LT(norm3(sim.omega), 0.1);
  std::printf("DdpSingleRigidBody closed loop done: first solve %d iterations, final pos error %.4f\n", first_iter,
              norm3(sub3(sim.pos, rd.pos)));

  {
    Srb a(mass, horizon_dt, 40, wp), b(mass, horizon_dt, 40, wp);
    a.config().max_iter = 30;
    b.config().max_iter = 30;
    std::vector<Srb::BatchItem> items(16);
    for(int i = 0; i < 16; i++)
    {
      items[i].initial_param.pos = {0.002 * i, -0.001 * i, 1.0};
      items[i].initial_param.ori = {0.01 * (i % 3), 0.0, -0.01 * (i % 2)};
      items[i].initial_param.angular_vel = {0.0, 0.02 * (i % 4), 0.0};
    }
    const auto batch = a.planBatch({motion_param_func}, {ref_data_func}, items, 0.0);
    double worst = 0;
    for(int i = 0; i < 16; i += 5)
    {
      const auto one = b.planOnce(motion_param_func, ref_data_func, items[i].initial_param, 0.0);
      for(size_t j = 0; j < one.size(); j++) worst = std::max(worst, std::fabs(one[j] - batch[i][j]));
      EXPECT_TRUE(a.lastIter(i) == b.lastIter());
    }
    EXPECT_LT(worst, 1e-300);
    std::printf("DdpSingleRigidBody planBatch(16) vs planOnce: max diff %g\n", worst);
  }
  return finish("TestDdpSingleRigidBody");
}
