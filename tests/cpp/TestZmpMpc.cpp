/* tests/cpp/TestZmpMpc.cpp — closed loops of the reference's ZMP-based methods through the C++ drop-in
 * classes on the GPU engine:
 *   CCC::LinearMpcZmp            reference tests/src/TestLinearMpcZmp.cpp:15-125
 *   CCC::IntrinsicallyStableMpc  reference tests/src/TestIntrinsicallyStableMpc.cpp:15-123
 *   CCC::PreviewControlZmp       planBatch (GPU kernel) == planOnce (host) on a batch of perturbed states
 * plus planBatch == repeated planOnce for the two QP methods.  Needs a GPU (no CPU fallback).
 */
#include <functional>

#include "../../centroidalcontrolcollection_b200/include/CCC/IntrinsicallyStableMpc.h"
#include "../../centroidalcontrolcollection_b200/include/CCC/LinearMpcZmp.h"
#include "../../centroidalcontrolcollection_b200/include/CCC/PreviewControlZmp.h"
#include "TestFixtures.h"

using namespace fixtures;

static CCC::LinearMpcZmp::RefData lmpcRefData(const FootstepManager & fm, double t)
{
  // FootstepManager::makeLinearMpcZmpRefData (reference tests/src/FootstepManager.h:356-365; epsilon added twice)
  CCC::LinearMpcZmp::RefData rd;
  rd.zmp_limits = fm.zmpLimits(t + 1e-6);
  return rd;
}
static CCC::IntrinsicallyStableMpc::RefData ismpcRefData(const FootstepManager & fm, double t)
{
  // FootstepManager::makeIntrinsicallyStableMpcRefData (:370-380)
  t += 1e-6;
  CCC::IntrinsicallyStableMpc::RefData rd;
  rd.zmp = fm.refZmp(t);
  rd.zmp_limits = fm.zmpLimits(t);
  return rd;
}
static bool inside(const Vec2 & p, const std::array<Vec2, 2> & lim)
{
  return p[0] - lim[0][0] >= 0 && p[1] - lim[0][1] >= 0 && lim[1][0] - p[0] >= 0 && lim[1][1] - p[1] >= 0;
}

template<class Plan>
static void closedLoop(const char * name, Plan && plan)
{
  const double sim_dt = 0.005, com_height = 1.0;
  FootstepManager fm = walkingPlan();
  ComZmpSim2d sim(com_height, sim_dt);
  Vec2 planned_zmp = sim.pos();
  double t = 0;
  int ticks = 0;
  while(t < 10.0)
  {
    fm.update(t);
    planned_zmp = plan(fm, sim, planned_zmp, t, sim_dt);
    EXPECT_TRUE(inside(planned_zmp, fm.zmpLimits(t)));
    t += sim_dt;
    ticks++;
    sim.update(planned_zmp);
    for(double dtm : {4.5, 8.5})
      if(dtm <= t && t < dtm + sim_dt) sim.addDisturb({0.05, 0.05});
  }
  const auto lim = fm.zmpLimits(t);
  EXPECT_TRUE(inside(planned_zmp, lim));
  EXPECT_TRUE(inside(sim.pos(), lim));
  std::printf("%s: %d control cycles, final CoM (%.4f, %.4f), final ZMP (%.4f, %.4f)\n", name, ticks, sim.pos()[0], sim.pos()[1],
              planned_zmp[0], planned_zmp[1]);
}

int main()
{
  const double horizon_duration = 2.0, horizon_dt = 0.02, com_height = 1.0;

  {
    CCC::LinearMpcZmp mpc(com_height, horizon_duration, horizon_dt);
    closedLoop("LinearMpcZmp", [&](const FootstepManager & fm, const ComZmpSim2d & sim, const Vec2 & planned, double t, double sim_dt) {
      CCC::LinearMpcZmp::InitialParam ip;
      ip.pos = sim.pos();
      ip.vel = sim.vel();
      ip.acc = scale(kG / com_height, sub(sim.pos(), planned));
      const Vec2 z = mpc.planOnce([&](double tt) { return lmpcRefData(fm, tt); }, ip, t, sim_dt);
      EXPECT_TRUE(mpc.mpc_1d_->lastStatus(0) == 0 && mpc.mpc_1d_->lastStatus(1) == 0);
      return z;
    });
    // planBatch(2 schedules x 24 states) == planOnce one by one
    FootstepManager fm = walkingPlan();
    fm.update(1.8);
    std::vector<std::function<CCC::LinearMpcZmp::RefData(double)>> scheds = {
        [&](double tt) { return lmpcRefData(fm, tt); },
        [&](double tt) {
          auto rd = lmpcRefData(fm, tt);
          rd.zmp_limits[0][0] -= 0.01;
          rd.zmp_limits[1][0] += 0.01;
          return rd;
        }};
    std::vector<CCC::LinearMpcZmp::BatchItem> items(48);
    for(int i = 0; i < 48; i++)
    {
      items[i].schedule = i % 2;
      items[i].initial_param.pos = {0.001 * i, -0.0005 * i};
      items[i].initial_param.vel = {0.01 * (i % 7), -0.01 * (i % 3)};
      items[i].initial_param.acc = {0.02 * (i % 5), 0.0};
    }
    const auto batch = mpc.planBatch(scheds, items, 1.8, 0.005);
    double worst = 0;
    for(int i = 0; i < 48; i += 5)
    {
      const Vec2 one = mpc.planOnce(scheds[items[i].schedule], items[i].initial_param, 1.8, 0.005);
      worst = std::max(worst, norm(sub(one, batch[i])));
    }
    EXPECT_LT(worst, 1e-300);
    std::printf("LinearMpcZmp planBatch(48) vs planOnce: max diff %g\n", worst);
  }

  {
    CCC::IntrinsicallyStableMpc mpc(com_height, horizon_duration, horizon_dt);
    closedLoop("IntrinsicallyStableMpc",
               [&](const FootstepManager & fm, const ComZmpSim2d & sim, const Vec2 & planned, double t, double sim_dt) {
                 CCC::IntrinsicallyStableMpc::InitialParam ip;
                 ip.capture_point = add(sim.pos(), scale(std::sqrt(com_height / kG), sim.vel()));
                 ip.planned_zmp = planned;
                 const Vec2 z = mpc.planOnce([&](double tt) { return ismpcRefData(fm, tt); }, ip, t, sim_dt);
                 EXPECT_TRUE(mpc.mpc_1d_->lastStatus(0) == 0 && mpc.mpc_1d_->lastStatus(1) == 0);
                 return z;
               });
    FootstepManager fm = walkingPlan();
    fm.update(1.8);
    std::vector<std::function<CCC::IntrinsicallyStableMpc::RefData(double)>> scheds = {[&](double tt) { return ismpcRefData(fm, tt); }};
    std::vector<CCC::IntrinsicallyStableMpc::BatchItem> items(32);
    for(int i = 0; i < 32; i++)
    {
      items[i].initial_param.capture_point = {0.002 * i - 0.03, 0.001 * i - 0.015};
      items[i].initial_param.planned_zmp = {0.0, 0.0};
    }
    const auto batch = mpc.planBatch(scheds, items, 1.8, 0.005);
    double worst = 0;
    for(int i = 0; i < 32; i += 3)
    {
      const Vec2 one = mpc.planOnce(scheds[0], items[i].initial_param, 1.8, 0.005);
      worst = std::max(worst, norm(sub(one, batch[i])));
    }
    EXPECT_LT(worst, 1e-300);
    std::printf("IntrinsicallyStableMpc planBatch(32) vs planOnce: max diff %g\n", worst);
  }

  {
    // preview control: the batched GPU row kernel against the host dot product
    CCC::PreviewControlZmp pc(com_height, 2.0, 0.01);
    FootstepManager fm = walkingPlan();
    fm.update(1.5);
    std::vector<std::function<Vec2(double)>> scheds = {[&](double tt) { return fm.refZmp(tt); }};
    std::vector<CCC::PreviewControlZmp::BatchItem> items(100);
    for(int i = 0; i < 100; i++)
    {
      items[i].initial_param.pos = {0.001 * i, 0.0003 * i};
      items[i].initial_param.vel = {0.002 * (i % 11), -0.001 * (i % 13)};
      items[i].initial_param.acc = {0.01 * (i % 3), 0.01 * (i % 4)};
    }
    const auto batch = pc.planBatch(scheds, items, 1.5, 0.005);
    double worst = 0;
    for(int i = 0; i < 100; i++) worst = std::max(worst, norm(sub(pc.planOnce(scheds[0], items[i].initial_param, 1.5, 0.005), batch[i])));
    EXPECT_LT(worst, 1e-9); // different summation order of the 200-term dot product
    std::printf("PreviewControlZmp planBatch(100, GPU) vs planOnce (host): max diff %g\n", worst);
  }
  return finish("TestZmpMpc");
}
