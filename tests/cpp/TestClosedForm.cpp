/* tests/cpp/TestClosedForm.cpp — the reference's closed-loop tests of the closed-form ZMP controllers (reference
 * tests/src/TestDcmTracking.cpp:15-104, tests/src/TestFootGuidedControl.cpp:15-104, tests/src/TestSingularPreviewControlZmp.cpp:14-113)
 * through the drop-in classes CCC::DcmTracking, CCC::FootGuidedControl and CCC::SingularPreviewControlZmp, plus
 * planBatch == repeated planOnce and the reference's exceptions.
 */
#include "../../centroidalcontrolcollection_b200/include/CCC/DcmTracking.h"
#include "../../centroidalcontrolcollection_b200/include/CCC/FootGuidedControl.h"
#include "../../centroidalcontrolcollection_b200/include/CCC/SingularPreviewControlZmp.h"
#include "TestFixtures.h"

using namespace fixtures;

template<class Plan, class RefZmp>
void closedLoop(const char * name, Plan plan, RefZmp ref_zmp_of)
{
  const double sim_dt = 0.005, com_height = 1.0;
  FootstepManager fm = walkingPlan();
  ComZmpSim2d sim(com_height, sim_dt);
  Vec2 planned_zmp = {0, 0};
  double t = 0;
  while(t < 10.0)
  {
    fm.update(t);
    const Vec2 ip = add(sim.pos(), scale(std::sqrt(com_height / kG), sim.vel()));
    planned_zmp = plan(fm, ip, t);
    EXPECT_LT(norm(sub(planned_zmp, ref_zmp_of(fm, t))), 0.1);
    t += sim_dt;
    sim.update(planned_zmp);
    for(double td : {4.5, 8.5})
      if(td <= t && t < td + sim_dt) sim.addDisturb({0.05, 0.05});
  }
  const Vec2 ref = ref_zmp_of(fm, t);
  EXPECT_LT(norm(sub(planned_zmp, ref)), 1e-2);
  EXPECT_LT(norm(sub(sim.pos(), ref)), 1e-2);
  EXPECT_LT(norm(sim.vel()), 1e-2);
  std::printf("%s closed loop: final |zmp - ref| = %.2e\n", name, norm(sub(planned_zmp, ref)));
}

int main()
{
  CCC::DcmTracking dcm_tracking(1.0);
  CCC::FootGuidedControl foot_guided(1.0);
  auto dcm_ref = [](const FootstepManager & fm, double t) {
    CCC::DcmTracking::RefData rd;
    fm.makeDcmTrackingRefData(t, rd.current_zmp, rd.time_zmp_list);
    return rd;
  };
  auto fgc_ref = [](const FootstepManager & fm, double t) {
    CCC::FootGuidedControl::RefData rd;
    fm.makeFootGuidedControlRefData(t, rd.transit_start_zmp, rd.transit_end_zmp, rd.transit_start_time, rd.transit_duration);
    return rd;
  };
  closedLoop(
      "DcmTracking", [&](const FootstepManager & fm, const Vec2 & ip, double t) { return dcm_tracking.planOnce(dcm_ref(fm, t), ip, t); },
      [&](const FootstepManager & fm, double t) { return dcm_ref(fm, t).current_zmp; });
  closedLoop(
      "FootGuidedControl", [&](const FootstepManager & fm, const Vec2 & ip, double t) { return foot_guided.planOnce(fgc_ref(fm, t), ip, t); },
      [&](const FootstepManager & fm, double t) { return fm.refZmp(t); });

  // SingularPreviewControlZmp (reference tests/src/TestSingularPreviewControlZmp.cpp:14-113): the planned ZMP of the previous
  // cycle is part of the initial parameter, control_dt = sim_dt
  {
    const double sim_dt = 0.005, com_height = 1.0;
    CCC::SingularPreviewControlZmp pc(com_height, 2.0, 0.01);
    FootstepManager fm = walkingPlan();
    ComZmpSim2d sim(com_height, sim_dt);
    Vec2 planned_zmp = sim.pos();
    double t = 0;
    while(t < 10.0)
    {
      fm.update(t);
      CCC::SingularPreviewControlZmp::InitialParam ip;
      ip.pos = sim.pos();
      ip.vel = sim.vel();
      ip.planned_zmp = planned_zmp;
      planned_zmp = pc.planOnce([&](double tt) { return fm.refZmp(tt); }, ip, t, sim_dt);
      EXPECT_LT(norm(sub(planned_zmp, fm.refZmp(t))), 0.1);
      t += sim_dt;
      sim.update(planned_zmp);
      for(double td : {4.5, 8.5})
        if(td <= t && t < td + sim_dt) sim.addDisturb({0.05, 0.05});
    }
    const Vec2 ref = fm.refZmp(t);
    EXPECT_LT(norm(sub(planned_zmp, ref)), 1e-2);
    EXPECT_LT(norm(sub(sim.pos(), ref)), 1e-2);
    EXPECT_LT(norm(sim.vel()), 1e-2);
    std::printf("SingularPreviewControlZmp closed loop: final |zmp - ref| = %.2e\n", norm(sub(planned_zmp, ref)));
    // planBatch on two sampled sequences == planOnce; the 1-D class is the x axis
    FootstepManager fa = walkingPlan(), fb = walkingPlan();
    fa.update(0.0);
    for(int tick = 0; tick * 0.005 <= 2.3; tick++) fb.update(tick * 0.005);
    fb.update(2.3);
    const auto seq_a = pc.sample([&](double tt) { return fa.refZmp(tt); }, 0.0);
    const auto seq_b = pc.sample([&](double tt) { return fb.refZmp(tt); }, 2.3);
    std::vector<CCC::SingularPreviewControlZmp::InitialParam> ips(6);
    std::vector<int> pid(6);
    for(int i = 0; i < 6; i++)
    {
      pid[i] = i % 2;
      ips[i].pos = {0.03 * i, -0.01 * i};
      ips[i].vel = {0.1 - 0.02 * i, 0.05};
      ips[i].planned_zmp = {0.02 * i, 0.01};
    }
    const auto zb = pc.planBatch({seq_a, seq_b}, ips, pid, sim_dt);
    double worst = 0;
    for(int i = 0; i < 6; i++)
    {
      const Vec2 one = pid[i] == 0 ? pc.planOnce([&](double tt) { return fa.refZmp(tt); }, ips[i], 0.0, sim_dt)
                                   : pc.planOnce([&](double tt) { return fb.refZmp(tt); }, ips[i], 2.3, sim_dt);
      worst = std::max(worst, norm(sub(zb[i], one)));
    }
    EXPECT_LT(worst, 1e-300);
    CCC::SingularPreviewControlZmp1d pc1(com_height, 2.0, 0.01);
    CCC::SingularPreviewControlZmp1d::InitialParam ip1;
    ip1.pos = ips[2].pos[0];
    ip1.vel = ips[2].vel[0];
    ip1.planned_zmp = ips[2].planned_zmp[0];
    EXPECT_LT(std::abs(pc1.planOnce([&](double tt) { return fa.refZmp(tt)[0]; }, ip1, 0.0, sim_dt) - zb[2][0]), 1e-300);
  }

  // planBatch over three records of the plan == planOnce, and the reference's exceptions
  {
    std::vector<CCC::DcmTracking::RefData> drd;
    std::vector<CCC::FootGuidedControl::RefData> frd;
    std::vector<double> times = {0.4, 2.3, 5.05};
    for(double t0 : times)
    {
      FootstepManager fm = walkingPlan();
      for(int tick = 0; tick * 0.005 <= t0; tick++) fm.update(tick * 0.005);
      fm.update(t0);
      drd.push_back(dcm_ref(fm, t0));
      frd.push_back(fgc_ref(fm, t0));
    }
    std::vector<Vec2> ips(9);
    std::vector<int> pid(9);
    for(int i = 0; i < 9; i++)
    {
      pid[i] = i % 3;
      ips[i] = {0.05 * i - 0.1, 0.02 * i};
    }
    const auto db = dcm_tracking.planBatch(drd, times, ips, pid);
    const auto fb = foot_guided.planBatch(frd, times, ips, pid);
    double worst = 0;
    for(int i = 0; i < 9; i++)
    {
      worst = std::max(worst, norm(sub(db[i], dcm_tracking.planOnce(drd[pid[i]], ips[i], times[pid[i]]))));
      worst = std::max(worst, norm(sub(fb[i], foot_guided.planOnce(frd[pid[i]], ips[i], times[pid[i]]))));
    }
    EXPECT_LT(worst, 1e-300);
    std::printf("planBatch(9) vs planOnce: max diff %g\n", worst);
    bool threw = false;
    try
    {
      dcm_tracking.planOnce(drd[1], ips[0], times[1] + 100.0); // every switching time is in the past
    }
    catch(const std::runtime_error &)
    {
      threw = true;
    }
    EXPECT_TRUE(threw);
    threw = false;
    CCC::FootGuidedControl::RefData bad = frd[0];
    bad.transit_duration = -0.1;
    try
    {
      foot_guided.planOnce(bad, ips[0], times[0]);
    }
    catch(const std::runtime_error &)
    {
      threw = true;
    }
    EXPECT_TRUE(threw);
  }
  return finish("TestClosedForm");
}
