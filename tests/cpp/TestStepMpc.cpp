/* tests/cpp/TestStepMpc.cpp — the reference's closed-loop test of CCC::StepMpc with online footstep update (reference
 * tests/src/TestStepMpc.cpp:16-141) through the drop-in class, plus planBatch == repeated planOnce and the 1-D class.
 */
#include "../../centroidalcontrolcollection_b200/include/CCC/StepMpc.h"
#include "TestFixtures.h"

using namespace fixtures;

static CCC::StepMpc::RefData refData(const FootstepManager & fm, double t)
{
  CCC::StepMpc::RefData rd;
  for(const auto & e : fm.makeStepMpcRefData(t)) rd.element_list.push_back({e.is_single_support, e.zmp, e.end_time});
  return rd;
}

int main()
{
  const double sim_dt = 0.005, com_height = 1.0;
  CCC::StepMpc step_mpc(com_height);
  FootstepManager fm = walkingPlan();
  ComZmpSim2d sim(com_height, sim_dt);
  Vec2 planned_zmp = {0, 0};
  double t = 0;
  int updates = 0;
  while(t < 10.0)
  {
    fm.update(t);
    CCC::StepMpc::InitialParam ip;
    ip.pos = sim.pos();
    ip.vel = sim.vel();
    const auto planned = step_mpc.planOnce(refData(fm, t), ip, t);
    planned_zmp = planned.current_zmp;
    EXPECT_LT(norm(sub(planned_zmp, fm.refZmp(t))), 0.2);
    t += sim_dt;
    sim.update(planned_zmp);
    for(double td : {4.5, 8.5})
      if(td <= t && t < td + sim_dt) sim.addDisturb({0.05, 0.05});
    // online footstep update (:113-128)
    if(!fm.footstep_list_.empty())
    {
      const double pre = fm.footstep_list_.front().swing_end_time - 0.1;
      if(pre <= t && t < pre + sim_dt && planned.next_foot_zmp)
      {
        fm.footstep_list_.front().pos = *planned.next_foot_zmp;
        updates++;
      }
    }
  }
  const Vec2 ref = fm.refZmp(t);
  EXPECT_LT(norm(sub(planned_zmp, ref)), 1e-2);
  EXPECT_LT(norm(sub(sim.pos(), ref)), 1e-2);
  EXPECT_LT(norm(sim.vel()), 1e-2);
  EXPECT_TRUE(updates >= 5);
  std::printf("StepMpc closed loop: final |zmp - ref| = %.2e, %d footsteps updated online\n", norm(sub(planned_zmp, ref)), updates);

  // planBatch over three records == planOnce; the 1-D class is one axis
  {
    std::vector<CCC::StepMpc::RefData> rds;
    std::vector<double> times = {0.4, 2.65, 6.2};
    for(double t0 : times)
    {
      FootstepManager f = walkingPlan();
      for(int tick = 0; tick * 0.005 <= t0; tick++) f.update(tick * 0.005);
      f.update(t0);
      rds.push_back(refData(f, t0));
    }
    std::vector<CCC::StepMpc::InitialParam> ips(9);
    std::vector<int> pid(9);
    for(int i = 0; i < 9; i++)
    {
      pid[i] = i % 3;
      ips[i].pos = {0.05 * i, 0.01 * i - 0.03};
      ips[i].vel = {0.1 - 0.02 * i, 0.03};
    }
    const auto pb = step_mpc.planBatch(rds, times, ips, pid);
    double worst = 0;
    for(int i = 0; i < 9; i++)
    {
      const auto one = step_mpc.planOnce(rds[pid[i]], ips[i], times[pid[i]]);
      worst = std::max(worst, norm(sub(pb[i].current_zmp, one.current_zmp)));
      EXPECT_TRUE(pb[i].next_foot_zmp.has_value() == one.next_foot_zmp.has_value());
      if(one.next_foot_zmp) worst = std::max(worst, norm(sub(*pb[i].next_foot_zmp, *one.next_foot_zmp)));
    }
    EXPECT_LT(worst, 1e-300);
    CCC::StepMpc1d mpc1(com_height);
    CCC::StepMpc1d::RefData rd1;
    for(const auto & e : rds[1].element_list) rd1.element_list.push_back({e.is_single_support, e.zmp[1], e.end_time});
    const auto p1 = mpc1.planOnce(rd1, {ips[1].pos[1], ips[1].vel[1]}, times[1]);
    EXPECT_LT(std::abs(p1.current_zmp - pb[1].current_zmp[1]), 1e-300);
    // an empty reference is refused
    bool threw = false;
    try
    {
      step_mpc.planOnce(CCC::StepMpc::RefData(), ips[0], 0.0);
    }
    catch(const std::invalid_argument &)
    {
      threw = true;
    }
    EXPECT_TRUE(threw);
  }
  return finish("TestStepMpc");
}
