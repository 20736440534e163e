/* tests/cpp/TestDdpCentroidal.cpp — the reference's TestDdpCentroidal.PlanOnce closed loop
 * (reference tests/src/TestDdpCentroidal.cpp:15-163) driven through the drop-in C++ class
 * CCC::DdpCentroidal of this repository (GoogleTest is absent here: plain checks, exit code).
 * Also exercises planBatch (new API) against repeated planOnce calls.
 *
 * Build: g++ -std=c++17 -O2 TestDdpCentroidal.cpp -L<pkg> -lccc_b200 -Wl,-rpath,<pkg>   (needs a GPU to run)
 */
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../centroidalcontrolcollection_b200/include/CCC/DdpCentroidal.h"

namespace
{
constexpr double kG = 9.80665;
int g_failures = 0;
#define EXPECT_LT(a, b)                                                                            \
  do                                                                                               \
  {                                                                                                \
    if(!((a) < (b)))                                                                               \
    {                                                                                              \
      std::printf("FAILED %s:%d: %s = %g is not < %s = %g\n", __FILE__, __LINE__, #a, (double)(a), #b, (double)(b)); \
      g_failures++;                                                                                \
    }                                                                                              \
  } while(0)

using CCC::Vector3d;
double norm(const Vector3d & v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
Vector3d sub(const Vector3d & a, const Vector3d & b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
Vector3d cross(const Vector3d & a, const Vector3d & b)
{
  return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}

/** makeContactFromRect, reference tests/src/ContactManager.h:10-21 */
std::shared_ptr<ForceColl::Contact> makeContactFromRect(double x0, double y0, double x1, double y1)
{
  std::vector<Vector3d> v = {{x0, y0, 0.0}, {x0, y1, 0.0}, {x1, y1, 0.0}, {x1, y0, 0.0}};
  return std::make_shared<ForceColl::SurfaceContact>("ContactFromRect", 0.5, v);
}

/** CentroidalSim, reference tests/src/SimModels.h:233-332 (A^2 = 0: exact ZOH polynomials). */
struct CentroidalSim
{
  double mass, dt;
  Vector3d inertia;
  Vector3d pos{0, 0, 0}, ang{0, 0, 0}, vel{0, 0, 0}, omega{0, 0, 0}, P{0, 0, 0}, L{0, 0, 0};
  void update(const Vector3d & f, const Vector3d & n)
  {
    for(int a = 0; a < 3; a++)
    {
      const double acc = f[a] / mass + (a == 2 ? -kG : 0.0);
      const double aacc = n[a] / inertia[a];
      pos[a] += vel[a] * dt + 0.5 * dt * dt * acc;
      ang[a] += omega[a] * dt + 0.5 * dt * dt * aacc;
      vel[a] += dt * acc;
      omega[a] += dt * aacc;
      P[a] += dt * (f[a] + (a == 2 ? -mass * kG : 0.0));
      L[a] += dt * n[a];
    }
  }
};

/** ForceColl::calcTotalWrench about `origin` */
void totalWrench(const CCC::DdpCentroidal::MotionParam & mp, const CCC::VectorXd & scales, const Vector3d & origin, Vector3d & f, Vector3d & n)
{
  f = {0, 0, 0};
  n = {0, 0, 0};
  int j = 0;
  for(const auto & c : mp.contact_list)
    for(const auto & vr : c->vertexWithRidgeList_)
      for(const auto & r : vr.ridgeList)
      {
        const Vector3d m = cross(sub(vr.vertex, origin), r);
        for(int a = 0; a < 3; a++)
        {
          f[a] += scales[j] * r[a];
          n[a] += scales[j] * m[a];
        }
        j++;
      }
}
} // namespace

int main()
{
  const double horizon_dt = 0.03, horizon_duration = 3.0, sim_dt = 0.005, mass = 100.0;
  const int horizon_steps = static_cast<int>(horizon_duration / horizon_dt);

  CCC::DdpCentroidal::WeightParam weight_param;
  weight_param.running_pos = {1.0, 1.0, 10.0};
  weight_param.terminal_pos = {1.0, 1.0, 10.0};
  CCC::DdpCentroidal ddp(mass, horizon_dt, horizon_steps, weight_param);

  auto motion_param_func = [](double t) {
    t += 1e-6;
    CCC::DdpCentroidal::MotionParam mp;
    if(t < 1.4)
      mp.contact_list.push_back(makeContactFromRect(-0.1, -0.1, 0.1, 0.1));
    else if(t < 1.6)
    {
    }
    else
      mp.contact_list.push_back(makeContactFromRect(0.4, -0.1, 0.6, 0.1));
    return mp;
  };
  auto ref_data_func = [](double t) {
    t += 1e-6;
    CCC::DdpCentroidal::RefData rd;
    if(t < 1.4)
      rd.pos = {0.0, 0.0, 1.0};
    else if(t < 1.6)
      rd.pos = {0.25, 0.0, 1.2};
    else
      rd.pos = {0.5, 0.0, 1.0};
    return rd;
  };

  CentroidalSim sim{mass, sim_dt, {40.0, 20.0, 10.0}};
  sim.pos = ref_data_func(0.0).pos;

  std::vector<double> durations;
  double t = 0;
  int first_iter = 0;
  Vector3d sim_pos_at_600 = {0, 0, 0};
  while(t < 3.0)
  {
    if(durations.size() == 600) sim_pos_at_600 = sim.pos;
    auto start = std::chrono::steady_clock::now();
    CCC::DdpCentroidal::InitialParam ip;
    ip.pos = sim.pos;
    ip.vel = sim.vel;
    ip.angular_momentum = sim.L;
    // the reference's own expressions (tests/src/TestDdpCentroidal.cpp:102-116), through the ddp_solver_ / ddp_problem_ views
    ip.u_list = ddp.ddp_solver_->controlData().u_list;
    if(!ip.u_list.empty())
    {
      for(int i = 0; i < ddp.ddp_solver_->config().horizon_steps; i++)
      {
        double tmp_time = t + i * ddp.ddp_problem_->dt();
        int input_dim = ddp.ddp_problem_->inputDim(tmp_time);
        if(static_cast<int>(ip.u_list[i].size()) != input_dim) ip.u_list[i].assign(input_dim, 0.0); // Eigen: setZero(input_dim)
      }
    }
    CCC::VectorXd scales = ddp.planOnce(motion_param_func, ref_data_func, ip, t);
    if(durations.empty()) first_iter = ddp.ddp_solver_->traceDataList().back().iter; // :129
    ddp.ddp_solver_->config().max_iter = 1; // Set max_iter from second simulation iteration (:116)
    durations.push_back(1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count());

    const auto mp = motion_param_func(t);
    const auto rd = ref_data_func(t);
    Vector3d f, n;
    totalWrench(mp, scales, sim.pos, f, n);
    EXPECT_LT(norm(sub(sim.pos, rd.pos)), 2.0);
    EXPECT_LT(norm(sim.vel), 2.0);
    EXPECT_LT(norm(sim.L), 1.0);

    t += sim_dt;
    sim.update(f, n);
    if(1.0 <= t && t < 1.0 + sim_dt)
    {
      sim.vel[0] += 0.05;
      sim.vel[1] += 0.05;
    }
  }
  const auto rd = ref_data_func(t);
  EXPECT_LT(norm(sub(sim.pos, rd.pos)), 0.1);
  EXPECT_LT(norm(sim.vel), 0.1);
  EXPECT_LT(norm(sim.L), 0.01);

  double mean = 0, mx = 0;
  for(double d : durations)
  {
    mean += d;
    mx = std::max(mx, d);
  }
  mean /= durations.size();
  std::printf("PlanOnce: %zu control cycles, first solve %d DDP iterations\n", durations.size(), first_iter);
  std::printf("Computation time per control cycle:\n  mean: %.3f [ms], max: %.3f [ms]\n", mean, mx);

  // ---- planBatch: 64 perturbed initial states on the same schedule == 64 separate planOnce calls ----
  {
    CCC::DdpCentroidal a(mass, horizon_dt, 50, weight_param), b(mass, horizon_dt, 50, weight_param);
    std::vector<CCC::DdpCentroidal::BatchItem> items(64);
    for(int i = 0; i < 64; i++)
    {
      items[i].initial_param.pos = {0.001 * i, -0.0005 * i, 1.0};
      items[i].initial_param.vel = {0.01 * (i % 5), 0.0, 0.0};
    }
    auto batch = a.planBatch({motion_param_func}, {ref_data_func}, items, 0.0);
    double worst = 0;
    for(int i = 0; i < 64; i += 9)
    {
      auto one = b.planOnce(motion_param_func, ref_data_func, items[i].initial_param, 0.0);
      for(size_t j = 0; j < one.size(); j++) worst = std::max(worst, std::fabs(one[j] - batch[i][j]));
      if(a.lastIter(i) != b.lastIter())
      {
        std::printf("FAILED: batch item %d: %d iterations, single %d\n", i, a.lastIter(i), b.lastIter());
        g_failures++;
      }
    }
    EXPECT_LT(worst, 1e-300); // identical bits
    std::printf("planBatch(64) vs planOnce: max |du0| = %g\n", worst);
  }

  // ---- runClosedLoopBatch: the same loop for 8 plants on the device in one call; plant 0 starts like the loop above ----
  {
    CCC::DdpCentroidal loop(mass, horizon_dt, horizon_steps, weight_param);
    std::vector<CCC::DdpCentroidal::BatchItem> items(8);
    for(int i = 0; i < 8; i++)
    {
      items[i].initial_param.pos = ref_data_func(0.0).pos;
      items[i].initial_param.pos[0] += 0.002 * i;
      items[i].initial_param.vel = {0.0, 0.005 * i, 0.0};
    }
    const int ticks = 600;
    auto start = std::chrono::steady_clock::now();
    const auto res = loop.runClosedLoopBatch({motion_param_func}, {ref_data_func}, items, 0.0, sim_dt, ticks, 1, 199, {0.05, 0.05, 0.0});
    const double ms = 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    for(int b = 0; b < 8; b++)
    {
      const double * end = res.state(b, ticks);
      const auto rd_end = ref_data_func(ticks * sim_dt);
      EXPECT_LT(norm(sub({end[0], end[1], end[2]}, rd_end.pos)), 0.1);
      EXPECT_LT(norm({end[3], end[4], end[5]}), 0.1);
      EXPECT_LT(norm({end[6], end[7], end[8]}), 0.01);
    }
    // plant 0 against the host-driven loop above (different plant arithmetic order: compare to 1e-6)
    EXPECT_LT(std::fabs(res.state(0, ticks)[0] - sim_pos_at_600[0]) + std::fabs(res.state(0, ticks)[1] - sim_pos_at_600[1])
                  + std::fabs(res.state(0, ticks)[2] - sim_pos_at_600[2]),
              1e-6);
    std::printf("runClosedLoopBatch: 8 plants x %d cycles in %.1f ms (%.3f ms per control cycle of the batch), first cycle %d iterations\n",
                ticks, ms, ms / ticks, res.iter(0, 0));
  }

  if(g_failures)
  {
    std::printf("%d check(s) FAILED\n", g_failures);
    return 1;
  }
  std::printf("ALL CHECKS PASSED\n");
  return 0;
}
