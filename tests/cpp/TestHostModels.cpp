/* tests/cpp/TestHostModels.cpp — host-side formulation classes of the C++ drop-in (no GPU needed):
 *  - StateSpaceModel zero-order hold vs Euler, with and without the affine term
 *    (reference tests/src/TestStateSpaceModel.cpp:64-136) and vs the closed form of the jerk model;
 *  - InvariantSequentialExtension / VariantSequentialExtension: condensed == iterated
 *    (reference tests/src/TestInvariantSequentialExtension.cpp:68-92, TestVariantSequentialExtension.cpp:178-224);
 *  - PreviewControl gains: Riccati residual, and the whole reference closed loop of
 *    tests/src/TestPreviewControlZmp.cpp:15-105 (BASELINE config 1: single problem on the CPU).
 */
#include <functional>
#include <memory>

#include "../../centroidalcontrolcollection_b200/include/CCC/CommonModels.h"
#include "../../centroidalcontrolcollection_b200/include/CCC/InvariantSequentialExtension.h"
#include "../../centroidalcontrolcollection_b200/include/CCC/PreviewControlZmp.h"
#include "../../centroidalcontrolcollection_b200/include/CCC/VariantSequentialExtension.h"
#include "TestFixtures.h"

using namespace fixtures;
using CCC::detail::Matrix;

static double diffNorm(const std::vector<double> & a, const std::vector<double> & b)
{
  double s = 0;
  for(size_t i = 0; i < a.size(); i++) s += (a[i] - b[i]) * (a[i] - b[i]);
  return std::sqrt(s);
}
static std::vector<double> condensed(const Matrix & A, const Matrix & B, const std::vector<double> & E,
                                     const std::vector<double> & x0, const std::vector<double> & u)
{
  std::vector<double> y(E);
  for(int i = 0; i < A.rows(); i++)
  {
    for(int j = 0; j < A.cols(); j++) y[i] += A(i, j) * x0[j];
    for(int j = 0; j < B.cols(); j++) y[i] += B(i, j) * u[j];
  }
  return y;
}

static void testStateSpaceModel()
{
  const double dt = 0.01;
  CCC::ComZmpModelJerkInput jerk(1.0);
  jerk.calcDiscMatrix(dt);
  const double ad[3][3] = {{1, dt, dt * dt / 2}, {0, 1, dt}, {0, 0, 1}};
  const double bd[3] = {dt * dt * dt / 6, dt * dt / 2, dt};
  for(int i = 0; i < 3; i++)
  {
    for(int j = 0; j < 3; j++) EXPECT_LT(std::fabs(jerk.Ad_(i, j) - ad[i][j]), 1e-15);
    EXPECT_LT(std::fabs(jerk.Bd_(i, 0) - bd[i]), 1e-15);
  }
  for(int with_e = 0; with_e < 2; with_e++)
  {
    CCC::StateSpaceModel s(3, 1, 1);
    s.A_(0, 1) = 1;
    s.A_(1, 2) = 1;
    s.B_(2, 0) = 1;
    if(with_e) s.E_ = {-1.0, 2.0, -3.0};
    s.calcDiscMatrix(dt);
    const std::vector<double> x = {1.0, 2.0, 3.0}, u = {-0.5};
    std::vector<double> euler = s.stateEq(x, u);
    for(int i = 0; i < 3; i++) euler[i] = x[i] + dt * euler[i];
    EXPECT_LT(diffNorm(s.stateEqDisc(x, u), euler), 1e-3);
  }
  bool threw = false;
  try
  {
    CCC::StateSpaceModel bad(0, 1, 1);
  }
  catch(const std::runtime_error &)
  {
    threw = true;
  }
  EXPECT_TRUE(threw);
  // a model with genuinely non-nilpotent dynamics: exp of [[0,1],[w^2,0]] dt has cosh / sinh entries
  CCC::StateSpaceModel lip(2, 1, 0);
  const double w = 3.0;
  lip.A_(0, 1) = 1;
  lip.A_(1, 0) = w * w;
  lip.B_(1, 0) = -w * w;
  lip.calcDiscMatrix(0.05);
  EXPECT_LT(std::fabs(lip.Ad_(0, 0) - std::cosh(w * 0.05)), 1e-14);
  EXPECT_LT(std::fabs(lip.Ad_(0, 1) - std::sinh(w * 0.05) / w), 1e-14);
  EXPECT_LT(std::fabs(lip.Bd_(0, 0) - (1 - std::cosh(w * 0.05))), 1e-14);
}

static void testInvariantSequentialExtension()
{
  const double dt = 0.01;
  const int N = 5;
  auto model = std::make_shared<CCC::ComZmpModelJerkInput>(1.0);
  model->calcDiscMatrix(dt);
  const std::vector<double> x0 = {1.0, 2.0, 3.0}, u = {5.0, 2.5, 0.0, -1.0, -2.0};
  CCC::InvariantSequentialExtension ext(model, N, false), ext_o(model, N, true);
  std::vector<double> xs, ys, x = x0;
  for(int k = 0; k < N; k++)
  {
    x = model->stateEqDisc(x, {u[k]});
    xs.insert(xs.end(), x.begin(), x.end());
    ys.push_back(model->observEq(x)[0]);
  }
  EXPECT_LT(diffNorm(condensed(ext.A_seq_, ext.B_seq_, ext.E_seq_, x0, u), xs), 1e-10);
  EXPECT_LT(diffNorm(condensed(ext_o.A_seq_, ext_o.B_seq_, ext_o.E_seq_, x0, u), ys), 1e-10);
  EXPECT_TRUE(ext.totalStateDim() == 15 && ext.totalInputDim() == 5 && ext_o.totalOutputDim() == 5);
}

static void testVariantSequentialExtension()
{
  // input dimension changes per stage, stages {4, 5, 6, 8} of 10 have no input
  // (reference tests/src/TestVariantSequentialExtension.cpp:178-224)
  const double dt = 0.01;
  std::vector<std::shared_ptr<CCC::StateSpaceModel>> models;
  std::vector<int> dims = {1, 1, 1, 1, 0, 0, 0, 1, 0, 1};
  for(int d : dims)
  {
    auto m = std::make_shared<CCC::StateSpaceModel>(3, d, 1);
    m->A_(0, 1) = 1;
    m->A_(1, 2) = 1;
    if(d) m->B_(2, 0) = 1;
    m->C_(0, 0) = 1;
    m->C_(0, 2) = -0.1;
    m->E_ = {0.0, 0.0, d ? 0.0 : -0.3};
    m->calcDiscMatrix(dt);
    models.push_back(m);
  }
  const std::vector<double> x0 = {1.0, 2.0, 3.0}, u = {0.5, 1.0, 0.5, 1.0, -1.0, -2.0};
  CCC::VariantSequentialExtension ext(models, false), ext_o(models, true);
  EXPECT_TRUE(ext.totalInputDim() == 6 && ext.totalStateDim() == 30 && ext_o.totalOutputDim() == 10);
  std::vector<double> xs, ys, x = x0;
  int acc = 0;
  for(size_t k = 0; k < models.size(); k++)
  {
    std::vector<double> uk(u.begin() + acc, u.begin() + acc + dims[k]);
    x = models[k]->stateEqDisc(x, uk);
    acc += dims[k];
    xs.insert(xs.end(), x.begin(), x.end());
    ys.push_back(models[k]->observEq(x)[0]);
  }
  EXPECT_LT(diffNorm(condensed(ext.A_seq_, ext.B_seq_, ext.E_seq_, x0, u), xs), 1e-10);
  EXPECT_LT(diffNorm(condensed(ext_o.A_seq_, ext_o.B_seq_, ext_o.E_seq_, x0, u), ys), 1e-10);
}

static void testPreviewControlZmp()
{
  const double horizon_duration = 2.0, horizon_dt = 0.01, sim_dt = 0.005, com_height = 1.0;
  CCC::PreviewControlZmp pc(com_height, horizon_duration, horizon_dt);
  EXPECT_TRUE(pc.preview_control_1d_->riccati_converged_);
  EXPECT_LT(pc.preview_control_1d_->riccati_error_, 1e-6);
  EXPECT_TRUE(pc.preview_control_1d_->horizon_steps_ == 200);
  bool threw = false;
  try
  {
    CCC::PreviewControlZmp bad(1.0, -1.0, 0.01);
  }
  catch(const std::runtime_error &)
  {
    threw = true;
  }
  EXPECT_TRUE(threw);

  FootstepManager fm = walkingPlan();
  ComZmpSim2d sim(com_height, sim_dt);
  Vec2 planned_zmp = sim.pos();
  double t = 0;
  while(t < 10.0)
  {
    fm.update(t);
    CCC::PreviewControlZmp::InitialParam ip;
    ip.pos = sim.pos();
    ip.vel = sim.vel();
    ip.acc = scale(kG / com_height, sub(sim.pos(), planned_zmp));
    planned_zmp = pc.planOnce([&](double tt) { return fm.refZmp(tt); }, ip, t, sim_dt);
    EXPECT_LT(norm(sub(planned_zmp, fm.refZmp(t))), 0.1);
    t += sim_dt;
    sim.update(planned_zmp);
    for(double dtm : {4.5, 8.5})
      if(dtm <= t && t < dtm + sim_dt) sim.addDisturb({0.05, 0.05});
  }
  const Vec2 ref_zmp = fm.refZmp(t);
  EXPECT_LT(norm(sub(planned_zmp, ref_zmp)), 1e-2);
  EXPECT_LT(norm(sub(sim.pos(), ref_zmp)), 1e-2);
  EXPECT_LT(norm(sim.vel()), 1e-2);
}

int main()
{
  testStateSpaceModel();
  testInvariantSequentialExtension();
  testVariantSequentialExtension();
  testPreviewControlZmp();
  return finish("TestHostModels");
}
