/* tests/cpp/TestFixtures.h — fixtures shared by the C++ drop-in tests: check macros, the footstep manager
 * (reference tests/src/FootstepManager.h:19-254, 356-380) and the CoM-ZMP plants (reference
 * tests/src/SimModels.h:11-226), restated without Eigen / GoogleTest.  Test infrastructure only.
 */
#pragma once
#include <array>
#include <cmath>
#include <cstdio>
#include <deque>
#include <map>
#include <stdexcept>
#include <vector>

namespace fixtures
{
constexpr double kG = 9.80665;
inline int & failures()
{
  static int n = 0;
  return n;
}
#define EXPECT_TRUE(cond)                                                          \
  do                                                                               \
  {                                                                                \
    if(!(cond))                                                                    \
    {                                                                              \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);                \
      fixtures::failures()++;                                                      \
    }                                                                              \
  } while(0)
#define EXPECT_LT(a, b)                                                                                               \
  do                                                                                                                  \
  {                                                                                                                   \
    if(!((a) < (b)))                                                                                                  \
    {                                                                                                                 \
      std::printf("FAILED %s:%d: %s = %g is not < %s = %g\n", __FILE__, __LINE__, #a, (double)(a), #b, (double)(b)); \
      fixtures::failures()++;                                                                                         \
    }                                                                                                                 \
  } while(0)
inline int finish(const char * name)
{
  if(failures())
  {
    std::printf("%s: %d check(s) FAILED\n", name, failures());
    return 1;
  }
  std::printf("%s: ALL CHECKS PASSED\n", name);
  return 0;
}

using Vec2 = std::array<double, 2>;
inline Vec2 add(const Vec2 & a, const Vec2 & b) { return {a[0] + b[0], a[1] + b[1]}; }
inline Vec2 sub(const Vec2 & a, const Vec2 & b) { return {a[0] - b[0], a[1] - b[1]}; }
inline Vec2 scale(double s, const Vec2 & a) { return {s * a[0], s * a[1]}; }
inline double norm(const Vec2 & a) { return std::sqrt(a[0] * a[0] + a[1] * a[1]); }

enum class Foot
{
  Left = 0,
  Right
};
inline Foot opposite(Foot f) { return f == Foot::Left ? Foot::Right : Foot::Left; }

struct Footstep
{
  Foot foot;
  Vec2 pos;
  double transit_start_time, swing_start_time, swing_end_time, transit_end_time;
  Footstep(Foot _foot, const Vec2 & _pos, double _transit_start_time, double transit_duration, double swing_duration)
  : foot(_foot), pos(_pos), transit_start_time(_transit_start_time), swing_start_time(_transit_start_time + 0.5 * transit_duration),
    swing_end_time(_transit_start_time + 0.5 * transit_duration + swing_duration),
    transit_end_time(_transit_start_time + transit_duration + swing_duration)
  {
  }
};

using Footstance = std::map<Foot, Vec2>;
inline Vec2 midPos(const Footstance & s)
{
  if(s.size() == 1) return s.begin()->second;
  return scale(0.5, add(s.at(Foot::Left), s.at(Foot::Right)));
}
inline std::array<Vec2, 2> supportRegion(const Footstance & s)
{
  if(s.size() == 1) return {s.begin()->second, s.begin()->second};
  const Vec2 & l = s.at(Foot::Left);
  const Vec2 & r = s.at(Foot::Right);
  return {Vec2{std::min(l[0], r[0]), std::min(l[1], r[1])}, Vec2{std::max(l[0], r[0]), std::max(l[1], r[1])}};
}

class FootstepManager
{
public:
  explicit FootstepManager(const Footstance & initial = {{Foot::Left, {0.0, 0.1}}, {Foot::Right, {0.0, -0.1}}}) : footstance_(initial) {}

  void appendFootstep(const Footstep & fs)
  {
    if(!footstep_list_.empty() && fs.transit_start_time < footstep_list_.back().transit_end_time)
      throw std::runtime_error("transit_start_time of specified footstep must be after transit_end_time of last footstep");
    footstep_list_.push_back(fs);
  }

  void update(double t)
  {
    auto & fl = footstep_list_;
    if(!fl.empty() && fl.front().swing_end_time <= t) footstance_[fl.front().foot] = fl.front().pos;
    while(!fl.empty() && fl.front().transit_end_time < t)
    {
      prev_footstep_ = std::make_shared<Footstep>(fl.front());
      fl.pop_front();
    }
    ref_zmp_list_.clear();
    ref_footstance_list_.clear();
    if(fl.empty())
    {
      ref_zmp_list_.emplace(t, midPos(footstance_));
      ref_zmp_list_.emplace(t + horizon_duration_, midPos(footstance_));
      ref_footstance_list_.emplace(t, footstance_);
      ref_footstance_list_.emplace(t + horizon_duration_, footstance_);
      return;
    }
    if(t < fl.front().transit_start_time)
    {
      ref_zmp_list_.emplace(t, midPos(footstance_));
      ref_footstance_list_.emplace(t, footstance_);
    }
    Footstance tmp = footstance_;
    for(const auto & fs : fl)
    {
      if(fs.transit_start_time > t + horizon_duration_) break;
      ref_zmp_list_.emplace(fs.transit_start_time, midPos(tmp));
      ref_footstance_list_.emplace(fs.transit_start_time, tmp);
      tmp.erase(fs.foot);
      ref_zmp_list_.emplace(fs.swing_start_time, tmp.at(opposite(fs.foot)));
      ref_footstance_list_.emplace(fs.swing_start_time, tmp);
      tmp[fs.foot] = fs.pos;
      ref_zmp_list_.emplace(fs.swing_end_time, tmp.at(opposite(fs.foot)));
      ref_footstance_list_.emplace(fs.swing_end_time, tmp);
      ref_zmp_list_.emplace(fs.transit_end_time, midPos(tmp));
    }
    if(ref_zmp_list_.rbegin()->first < t + horizon_duration_)
    {
      ref_zmp_list_.emplace(t + horizon_duration_, midPos(tmp));
      ref_footstance_list_.emplace(t + horizon_duration_, tmp);
    }
  }

  Vec2 refZmp(double t) const
  {
    t += 1e-6;
    auto hi = ref_zmp_list_.upper_bound(t);
    auto lo = std::prev(hi);
    const double ratio = (t - lo->first) / (hi->first - lo->first);
    return add(scale(1 - ratio, lo->second), scale(ratio, hi->second));
  }

  std::array<Vec2, 2> zmpLimits(double t) const
  {
    t += 1e-6;
    auto it = std::prev(ref_footstance_list_.upper_bound(t));
    auto region = supportRegion(it->second);
    return {sub(region[0], scale(0.5, foot_size_)), add(region[1], scale(0.5, foot_size_))};
  }

  /** FootstepManager::makeDcmTrackingRefData (reference tests/src/FootstepManager.h:260-274; horizon_duration is an int
   *  there): current ZMP and the switching list within the horizon. */
  void makeDcmTrackingRefData(double current_time, Vec2 & current_zmp, std::map<double, Vec2> & time_zmp_list, int horizon_duration = 5) const
  {
    current_time += 1e-6;
    current_zmp = std::prev(ref_zmp_list_.upper_bound(current_time))->second;
    time_zmp_list.clear();
    for(auto it = ref_zmp_list_.upper_bound(current_time); it != ref_zmp_list_.end() && it->first < current_time + horizon_duration; it++)
      time_zmp_list.emplace(*it);
  }

  /** FootstepManager::makeFootGuidedControlRefData (:279-351). */
  void makeFootGuidedControlRefData(double current_time, Vec2 & start_zmp, Vec2 & end_zmp, double & start_time, double & duration) const
  {
    current_time += 1e-6;
    const double constant_zmp_duration = 1.0, concat_thre = 0.1, horizon_margin = 1e-3;
    if(footstep_list_.empty())
    {
      start_zmp = midPos(footstance_);
      end_zmp = start_zmp;
      start_time = current_time + constant_zmp_duration;
      duration = 0;
      return;
    }
    const Footstep & fs = footstep_list_.front();
    if(current_time < fs.swing_start_time)
    {
      start_zmp = midPos(footstance_);
      end_zmp = footstance_.at(opposite(fs.foot));
      start_time = fs.transit_start_time;
      duration = fs.swing_start_time - fs.transit_start_time;
      if(prev_footstep_ && prev_footstep_->foot == opposite(fs.foot) && fs.transit_start_time - prev_footstep_->transit_end_time < concat_thre)
      {
        start_zmp = footstance_.at(opposite(prev_footstep_->foot));
        start_time = prev_footstep_->swing_end_time;
        duration = fs.swing_start_time - prev_footstep_->swing_end_time;
      }
    }
    else
    {
      start_zmp = footstance_.at(opposite(fs.foot));
      Footstance tmp = footstance_;
      tmp[fs.foot] = fs.pos;
      end_zmp = midPos(tmp);
      start_time = fs.swing_end_time;
      duration = fs.transit_end_time - fs.swing_end_time;
      if(footstep_list_.size() >= 2)
      {
        const Footstep & nxt = footstep_list_[1];
        if(nxt.foot == opposite(fs.foot) && nxt.transit_start_time - fs.transit_end_time < concat_thre)
        {
          end_zmp = fs.pos;
          duration = nxt.swing_start_time - fs.swing_end_time;
        }
      }
    }
    if(start_time + duration < current_time + horizon_margin) duration += horizon_margin;
  }

  /** One support phase of FootstepManager::makeStepMpcRefData. */
  struct StepElement
  {
    bool is_single_support;
    Vec2 zmp;
    double end_time;
  };

  /** FootstepManager::makeStepMpcRefData (:385-448). */
  std::vector<StepElement> makeStepMpcRefData(double current_time) const
  {
    current_time -= 1e-6;
    const double constant_zmp_duration = 0.2, horizon_duration = 3.0;
    std::vector<StepElement> el;
    if(footstep_list_.empty())
    {
      el.push_back({false, midPos(footstance_), current_time + constant_zmp_duration});
      return el;
    }
    Footstance tmp = footstance_;
    if(current_time < footstep_list_.front().swing_start_time) el.push_back({false, midPos(tmp), footstep_list_.front().swing_start_time});
    for(size_t i = 0; i < footstep_list_.size(); i++)
    {
      const Footstep & fs = footstep_list_[i];
      if(i > 0 || current_time < fs.swing_end_time)
      {
        el.push_back({true, tmp.at(opposite(fs.foot)), fs.swing_end_time});
        tmp[fs.foot] = fs.pos;
      }
      el.push_back({false, midPos(tmp), i == footstep_list_.size() - 1 ? fs.transit_end_time : footstep_list_[i + 1].swing_start_time});
      if(el.back().end_time > current_time + horizon_duration) break;
    }
    return el;
  }

  Footstance footstance_;
  std::deque<Footstep> footstep_list_;
  std::shared_ptr<Footstep> prev_footstep_;
  double horizon_duration_ = 10.0;
  Vec2 foot_size_ = {0.1, 0.05};
  std::map<double, Vec2> ref_zmp_list_;
  std::map<double, Footstance> ref_footstance_list_;
};

/** The six-step walking plan of reference tests/src/TestLinearMpcZmp.cpp:30-41 (same in the other ZMP tests). */
inline FootstepManager walkingPlan()
{
  FootstepManager m;
  const double transit = 0.2, swing = 0.8;
  m.appendFootstep(Footstep(Foot::Left, {0.2, 0.1}, 2.0, transit, swing));
  m.appendFootstep(Footstep(Foot::Right, {0.4, -0.1}, 3.0, transit, swing));
  m.appendFootstep(Footstep(Foot::Left, {0.6, 0.1}, 4.0, transit, swing));
  m.appendFootstep(Footstep(Foot::Right, {0.8, -0.1}, 5.0, transit, swing));
  m.appendFootstep(Footstep(Foot::Left, {0.6, 0.1}, 6.0, transit, swing));
  m.appendFootstep(Footstep(Foot::Right, {0.6, -0.1}, 7.0, transit, swing));
  return m;
}

/** x'' = omega^2 (x - zmp) per axis, exact zero-order hold (reference tests/src/SimModels.h:11-138). */
struct ComZmpSim2d
{
  double ad[2][2], bd[2];
  Vec2 x = {0, 0}, y = {0, 0}; // (pos, vel) per axis
  ComZmpSim2d(double com_height, double sim_dt)
  {
    const double w = std::sqrt(kG / com_height), ch = std::cosh(w * sim_dt), sh = std::sinh(w * sim_dt);
    ad[0][0] = ch;
    ad[0][1] = sh / w;
    ad[1][0] = w * sh;
    ad[1][1] = ch;
    bd[0] = 1 - ch;
    bd[1] = -w * sh;
  }
  Vec2 pos() const { return {x[0], y[0]}; }
  Vec2 vel() const { return {x[1], y[1]}; }
  void update(const Vec2 & zmp)
  {
    x = {ad[0][0] * x[0] + ad[0][1] * x[1] + bd[0] * zmp[0], ad[1][0] * x[0] + ad[1][1] * x[1] + bd[1] * zmp[0]};
    y = {ad[0][0] * y[0] + ad[0][1] * y[1] + bd[0] * zmp[1], ad[1][0] * y[0] + ad[1][1] * y[1] + bd[1] * zmp[1]};
  }
  /** the reference adds impulse.x() to both axes (SimModels.h:127-128) */
  void addDisturb(const Vec2 & impulse_per_mass)
  {
    x[1] += impulse_per_mass[0];
    y[1] += impulse_per_mass[0];
  }
};

/** Horizontal model rebuilt with the current CoM height every tick; vertical double integrator with
 *  gravity (reference tests/src/SimModels.h:140-226). */
struct ComZmpSim3d
{
  double mass, dt;
  Vec2 x = {0, 0}, y = {0, 0}, z = {0, 0};
  ComZmpSim3d(double _mass, double _dt) : mass(_mass), dt(_dt) {}
  void update(const Vec2 & zmp, double force_z)
  {
    const double w = std::sqrt(kG / z[0]), ch = std::cosh(w * dt), sh = std::sinh(w * dt);
    x = {ch * x[0] + sh / w * x[1] + (1 - ch) * zmp[0], w * sh * x[0] + ch * x[1] - w * sh * zmp[0]};
    y = {ch * y[0] + sh / w * y[1] + (1 - ch) * zmp[1], w * sh * y[0] + ch * y[1] - w * sh * zmp[1]};
    const double a = force_z / mass - kG;
    z = {z[0] + dt * z[1] + 0.5 * dt * dt * a, z[1] + dt * a};
  }
  void addDisturb(const Vec2 & impulse_per_mass)
  {
    x[1] += impulse_per_mass[0];
    y[1] += impulse_per_mass[0];
  }
};
} // namespace fixtures

// ---- 3-D fixtures for the wrench-based methods -------------------------------------------------------
#include <memory>

#include "../../centroidalcontrolcollection_b200/include/CCC/Contact.h"

namespace fixtures
{
using Vec3 = std::array<double, 3>;
inline Vec3 sub3(const Vec3 & a, const Vec3 & b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
inline double norm3(const Vec3 & a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
inline Vec3 cross3(const Vec3 & a, const Vec3 & b)
{
  return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}

/** makeContactFromRect, reference tests/src/ContactManager.h:10-21 */
inline std::shared_ptr<ForceColl::Contact> makeContactFromRect(double x0, double y0, double x1, double y1)
{
  std::vector<Vec3> v = {{x0, y0, 0.0}, {x0, y1, 0.0}, {x1, y1, 0.0}, {x1, y0, 0.0}};
  return std::make_shared<ForceColl::SurfaceContact>("ContactFromRect", 0.5, v);
}

/** ForceColl::calcTotalWrench about `origin`. */
inline void totalWrench(const std::vector<std::shared_ptr<ForceColl::Contact>> & contact_list,
                        const std::vector<double> & scales,
                        const Vec3 & origin,
                        Vec3 & f,
                        Vec3 & n)
{
  f = {0, 0, 0};
  n = {0, 0, 0};
  int j = 0;
  for(const auto & c : contact_list)
    for(const auto & vr : c->vertexWithRidgeList_)
      for(const auto & r : vr.ridgeList)
      {
        const Vec3 m = cross3(sub3(vr.vertex, origin), r);
        for(int a = 0; a < 3; a++)
        {
          f[a] += scales[j] * r[a];
          n[a] += scales[j] * m[a];
        }
        j++;
      }
}

/** CentroidalSim, reference tests/src/SimModels.h:233-332 (the plant's A is nilpotent: exact ZOH polynomials).
 *  ang / omega are in (X, Y, Z) order. */
struct CentroidalSim
{
  double mass, dt;
  Vec3 inertia;
  Vec3 pos{0, 0, 0}, ang{0, 0, 0}, vel{0, 0, 0}, omega{0, 0, 0}, P{0, 0, 0}, L{0, 0, 0};
  void update(const Vec3 & f, const Vec3 & n)
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
} // namespace fixtures
