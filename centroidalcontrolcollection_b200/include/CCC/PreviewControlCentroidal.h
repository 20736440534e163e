/* CCC/PreviewControlCentroidal.h — drop-in host classes for CCC::PreviewControlCentroidal1d / CCC::PreviewControlCentroidal
 * (Murooka et al. 2022) on the C-ABI engine.
 *
 * Mirrors reference include/CCC/PreviewControlCentroidal.h and src/PreviewControlCentroidal.cpp: 1-D class =
 * PreviewControl<3,1,2> on CentroidalModel1d (src :10-21) with WeightParam (pos 2e2, wrench 5e-4, jerk 1e-8, :61);
 * PreviewControlCentroidal: MotionParam (:139-145), RefData (:148-157), InitialParam (:160-177), WeightParam
 * (:180-218), constructor (src :73-89), planOnce (:239-243, src :91-130).
 * SpaceVecAlg is absent: sva::MotionVecd / sva::ForceVecd are the small structs below with the same accessors
 * (angular(), linear() / moment(), force(), vector() = (angular | moment, linear | force)).  The wrench distribution
 * at the end of planOnce is ForceColl::WrenchDistribution (external; restated from its published formulation:
 * min ||G lambda - w||^2_W + eps ||lambda||^2 within ridge-force limits) as a QP on the engine; its configuration
 * (mc_rtc::Configuration in the reference) is the plain struct WrenchDistConfig.
 * planOnce = the six preview rows through ccc_preview_input + one QP through ccc_qp_solve_grouped; new: planBatch —
 * B initial states sharing the contacts and the reference (one QP matrix group per problem, because the grasp
 * matrix is taken about each problem's CoM).  Header-only; no CPU fallback.
 */
#pragma once
#include <array>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/ccc_b200.h"
#include "Contact.h"
#include "Gravity.h"
#include "PreviewControl.h"

namespace CCC
{
namespace sva
{
using Vector3d = std::array<double, 3>;
using Vector6d = std::array<double, 6>;
/** Spatial motion vector: (angular, linear). */
struct MotionVecd
{
  Vector3d angular_ = {0, 0, 0}, linear_ = {0, 0, 0};
  MotionVecd() {}
  MotionVecd(const Vector3d & angular, const Vector3d & linear) : angular_(angular), linear_(linear) {}
  static MotionVecd Zero() { return MotionVecd(); }
  Vector3d & angular() { return angular_; }
  const Vector3d & angular() const { return angular_; }
  Vector3d & linear() { return linear_; }
  const Vector3d & linear() const { return linear_; }
  Vector6d vector() const { return {angular_[0], angular_[1], angular_[2], linear_[0], linear_[1], linear_[2]}; }
};
/** Spatial force vector: (moment, force). */
struct ForceVecd
{
  Vector3d moment_ = {0, 0, 0}, force_ = {0, 0, 0};
  ForceVecd() {}
  ForceVecd(const Vector3d & moment, const Vector3d & force) : moment_(moment), force_(force) {}
  static ForceVecd Zero() { return ForceVecd(); }
  Vector3d & moment() { return moment_; }
  const Vector3d & moment() const { return moment_; }
  Vector3d & couple() { return moment_; }
  const Vector3d & couple() const { return moment_; }
  Vector3d & force() { return force_; }
  const Vector3d & force() const { return force_; }
  Vector6d vector() const { return {moment_[0], moment_[1], moment_[2], force_[0], force_[1], force_[2]}; }
};
} // namespace sva

class PreviewControlCentroidal1d : public PreviewControl
{
  friend class PreviewControlCentroidal;

public:
  using InitialParam = std::array<double, 3>;

  struct WeightParam
  {
    double pos, wrench, jerk;
    WeightParam(double _pos = 2e2, double _wrench = 5e-4, double _jerk = 1e-8) : pos(_pos), wrench(_wrench), jerk(_jerk) {}
    PreviewControl::WeightParam toPreviewControlWeightParam() const
    {
      PreviewControl::WeightParam w;
      w.output = {pos, wrench};
      w.input = {jerk};
      return w;
    }
  };

  /** m c'' = f with jerk input; outputs (c, m c'') (src/PreviewControlCentroidal.cpp:10-21). */
  class CentroidalModel1d : public StateSpaceModel
  {
  public:
    CentroidalModel1d(double inertia_param) : StateSpaceModel(3, 1, 2)
    {
      A_(0, 1) = 1;
      A_(1, 2) = 1;
      B_(2, 0) = 1;
      C_(0, 0) = 1;
      C_(1, 2) = inertia_param;
    }
  };

  PreviewControlCentroidal1d(double inertia_param, double horizon_duration, double horizon_dt, const WeightParam & weight_param = WeightParam())
  : PreviewControl(std::make_shared<CentroidalModel1d>(inertia_param), horizon_duration, horizon_dt, weight_param.toPreviewControlWeightParam())
  {
  }
};

class PreviewControlCentroidal
{
public:
  struct MotionParam
  {
    std::vector<std::shared_ptr<ForceColl::Contact>> contact_list;
  };

  struct RefData
  {
    sva::MotionVecd pos = sva::MotionVecd::Zero();
    sva::ForceVecd wrench = sva::ForceVecd::Zero();
  };

  struct InitialParam
  {
    sva::MotionVecd pos = sva::MotionVecd::Zero();
    sva::MotionVecd vel = sva::MotionVecd::Zero();
    sva::MotionVecd acc = sva::MotionVecd::Zero();
    /** (pos, vel, acc) of component idx: 0..2 rotational, 3..5 translational (src :46-57). */
    PreviewControlCentroidal1d::InitialParam toInitialParam1d(int idx) const { return {pos.vector()[idx], vel.vector()[idx], acc.vector()[idx]}; }
  };

  /** ForceColl::WrenchDistribution::Configuration (defaults as published by ForceColl). */
  struct WrenchDistConfig
  {
    sva::ForceVecd wrench_weight;
    double regular_weight;
    std::pair<double, double> ridge_force_min_max;
    WrenchDistConfig() : wrench_weight({1.0, 1.0, 1.0}, {1.0, 1.0, 1.0}), regular_weight(1e-8), ridge_force_min_max(3.0, 1000.0) {}
  };

  struct WeightParam
  {
    sva::MotionVecd pos;
    sva::ForceVecd wrench;
    sva::MotionVecd jerk;
    WrenchDistConfig wrench_dist_config;
    WeightParam(const sva::MotionVecd & _pos = sva::MotionVecd({1e2, 1e2, 1e2}, {2e2, 2e2, 2e2}),
                const sva::ForceVecd & _wrench = sva::ForceVecd({5e-3, 5e-3, 5e-3}, {5e-4, 5e-4, 5e-4}),
                const sva::MotionVecd & _jerk = sva::MotionVecd({1e-8, 1e-8, 1e-8}, {1e-8, 1e-8, 1e-8}),
                const WrenchDistConfig & _wrench_dist_config = WrenchDistConfig())
    : pos(_pos), wrench(_wrench), jerk(_jerk), wrench_dist_config(_wrench_dist_config)
    {
    }
    PreviewControlCentroidal1d::WeightParam toWeightParam1d(int idx) const
    {
      return PreviewControlCentroidal1d::WeightParam(pos.vector()[idx], wrench.vector()[idx], jerk.vector()[idx]);
    }
  };

public:
  PreviewControlCentroidal(double mass,
                           const sva::Vector3d & moment_of_inertia,
                           double horizon_duration,
                           double horizon_dt,
                           const WeightParam & weight_param = WeightParam())
  : mass_(mass), wrench_dist_config_(weight_param.wrench_dist_config)
  {
    for(int i = 0; i < 6; i++)
    {
      const double inertia_param = i < 3 ? moment_of_inertia[i] : mass_;
      preview_control_1d_[i] = std::make_shared<PreviewControlCentroidal1d>(inertia_param, horizon_duration, horizon_dt, weight_param.toWeightParam1d(i));
    }
  }
  ~PreviewControlCentroidal()
  {
    if(qp_ws_) ccc_qp_destroy(qp_ws_);
  }
  PreviewControlCentroidal(const PreviewControlCentroidal &) = delete;
  PreviewControlCentroidal & operator=(const PreviewControlCentroidal &) = delete;

  sva::ForceVecd planOnce(const MotionParam & motion_param,
                          const std::function<RefData(double)> & ref_data_func,
                          const InitialParam & initial_param,
                          double current_time,
                          double control_dt = -1)
  {
    return planBatch(motion_param, ref_data_func, {initial_param}, current_time, control_dt)[0];
  }

  /** planOnce for B initial states sharing the contacts and the reference. */
  std::vector<sva::ForceVecd> planBatch(const MotionParam & motion_param,
                                        const std::function<RefData(double)> & ref_data_func,
                                        const std::vector<InitialParam> & initial_params,
                                        double current_time,
                                        double control_dt = -1)
  {
    const int B = static_cast<int>(initial_params.size());
    if(B == 0) return {};
    const int N = preview_control_1d_[0]->horizon_steps_;
    const double horizon_dt = preview_control_1d_[0]->horizon_dt_;
    // ref_output_seq (src :101-110): row i = component i, columns (pos, wrench) interleaved over the horizon
    std::vector<double> ref(static_cast<size_t>(6) * 2 * N);
    for(int i = 0; i < N; i++)
    {
      const RefData rd = ref_data_func(current_time + (i + 1) * horizon_dt);
      const auto p = rd.pos.vector(), w = rd.wrench.vector();
      for(int c = 0; c < 6; c++)
      {
        ref[(static_cast<size_t>(c) * N + i) * 2] = p[c];
        ref[(static_cast<size_t>(c) * N + i) * 2 + 1] = w[c];
      }
    }
    if(control_dt < 0) control_dt = horizon_dt;
    // the six preview controllers: B rows each through the engine (src :112-124, 1-D procOnce :29-44)
    std::vector<double> desired(static_cast<size_t>(B) * 6), x(static_cast<size_t>(B) * 3), rows(static_cast<size_t>(B) * 2 * N), jerk(B);
    for(int c = 0; c < 6; c++)
    {
      const auto & pc = *preview_control_1d_[c];
      std::vector<double> K(3), F(static_cast<size_t>(2) * N);
      for(int j = 0; j < 3; j++) K[j] = pc.K_(0, j);
      for(int j = 0; j < 2 * N; j++) F[j] = pc.F_(0, j);
      for(int b = 0; b < B; b++)
      {
        const auto ip = initial_params[b].toInitialParam1d(c);
        for(int j = 0; j < 3; j++) x[static_cast<size_t>(b) * 3 + j] = ip[j];
        for(int j = 0; j < 2 * N; j++) rows[static_cast<size_t>(b) * 2 * N + j] = ref[static_cast<size_t>(c) * 2 * N + j];
      }
      if(ccc_preview_input(B, 2 * N, K.data(), F.data(), x.data(), rows.data(), jerk.data(), CCC_MEM_HOST, nullptr) != CCC_OK)
        throw std::runtime_error(std::string("[PreviewControlCentroidal] ") + ccc_last_error());
      for(int b = 0; b < B; b++)
      {
        const double acc = x[static_cast<size_t>(b) * 3 + 2] + control_dt * jerk[b];
        desired[static_cast<size_t>(b) * 6 + c] = pc.model_->C_(1, 2) * acc;
      }
    }
    for(int b = 0; b < B; b++) desired[static_cast<size_t>(b) * 6 + 5] += mass_ * constants::g; // src :125
    return distributeWrench(motion_param, desired, initial_params);
  }

  int lastStatus(int b = 0) const { return status_.at(b); }

public:
  //! Robot mass [kg]
  double mass_ = 0;
  //! One-dimensional preview controllers: three rotational, three translational components
  std::array<std::shared_ptr<PreviewControlCentroidal1d>, 6> preview_control_1d_;
  //! Configuration of the wrench distribution
  WrenchDistConfig wrench_dist_config_;

protected:
  /** ForceColl::WrenchDistribution::run(desired wrench, moment origin = CoM) for every problem (src :127-129). */
  std::vector<sva::ForceVecd> distributeWrench(const MotionParam & motion_param,
                                               const std::vector<double> & desired,
                                               const std::vector<InitialParam> & initial_params)
  {
    const int B = static_cast<int>(initial_params.size());
    std::vector<sva::Vector3d> vertex, ridge;
    for(const auto & contact : motion_param.contact_list)
      for(const auto & vr : contact->vertexWithRidgeList_)
        for(const auto & r : vr.ridgeList)
        {
          vertex.push_back(vr.vertex);
          ridge.push_back(r);
        }
    const int n = static_cast<int>(ridge.size());
    std::vector<sva::ForceVecd> out(B);
    status_.assign(B, 0);
    if(n == 0) return out; // no contact: zero wrench
    const auto w = wrench_dist_config_.wrench_weight.vector();
    std::vector<double> G(static_cast<size_t>(B) * 6 * n), Q(static_cast<size_t>(B) * n * n), c(static_cast<size_t>(B) * n), C(static_cast<size_t>(2) * n * n, 0.0),
        d(static_cast<size_t>(B) * 2 * n), xsol(static_cast<size_t>(B) * n);
    std::vector<int32_t> gid(B);
    for(int j = 0; j < n; j++)
    {
      C[static_cast<size_t>(j) * n + j] = -1.0;
      C[static_cast<size_t>(n + j) * n + j] = 1.0;
    }
    for(int b = 0; b < B; b++)
    {
      gid[b] = b;
      const auto & o = initial_params[b].pos.linear();
      double * Gb = G.data() + static_cast<size_t>(b) * 6 * n;
      for(int j = 0; j < n; j++)
      {
        const double rx = vertex[j][0] - o[0], ry = vertex[j][1] - o[1], rz = vertex[j][2] - o[2];
        Gb[0 * n + j] = ry * ridge[j][2] - rz * ridge[j][1];
        Gb[1 * n + j] = rz * ridge[j][0] - rx * ridge[j][2];
        Gb[2 * n + j] = rx * ridge[j][1] - ry * ridge[j][0];
        for(int a = 0; a < 3; a++) Gb[(3 + a) * n + j] = ridge[j][a];
      }
      double * Qb = Q.data() + static_cast<size_t>(b) * n * n;
      for(int i = 0; i < n; i++)
      {
        for(int j = 0; j < n; j++)
        {
          double s = 0;
          for(int k = 0; k < 6; k++) s += Gb[k * n + i] * w[k] * Gb[k * n + j];
          Qb[static_cast<size_t>(i) * n + j] = s + (i == j ? wrench_dist_config_.regular_weight : 0.0);
        }
        double s = 0;
        for(int k = 0; k < 6; k++) s += Gb[k * n + i] * w[k] * desired[static_cast<size_t>(b) * 6 + k];
        c[static_cast<size_t>(b) * n + i] = -1 * s;
        d[static_cast<size_t>(b) * 2 * n + i] = -wrench_dist_config_.ridge_force_min_max.first;
        d[static_cast<size_t>(b) * 2 * n + n + i] = wrench_dist_config_.ridge_force_min_max.second;
      }
    }
    if(!qp_ws_ || n != qp_n_ || B > qp_batch_)
    {
      if(qp_ws_) ccc_qp_destroy(qp_ws_);
      qp_ws_ = ccc_qp_create_grouped(n, 0, 2 * n, B, B);
      if(!qp_ws_) throw std::runtime_error(std::string("[PreviewControlCentroidal] ") + ccc_last_error());
      qp_n_ = n;
      qp_batch_ = B;
    }
    ccc_qp_batch_t bt{};
    bt.n = n;
    bt.n_eq = 0;
    bt.n_ineq = 2 * n;
    bt.batch = B;
    bt.Q = Q.data();
    bt.C = C.data();
    bt.c = c.data();
    bt.d = d.data();
    ccc_qp_result_t rs{};
    rs.x = xsol.data();
    rs.status = status_.data();
    if(ccc_qp_solve_grouped(qp_ws_, &bt, B, gid.data(), &rs, CCC_MEM_HOST, nullptr) != CCC_OK)
      throw std::runtime_error(std::string("[PreviewControlCentroidal] ") + ccc_last_error());
    for(int b = 0; b < B; b++)
    {
      double wr[6] = {0, 0, 0, 0, 0, 0};
      const double * Gb = G.data() + static_cast<size_t>(b) * 6 * n;
      for(int k = 0; k < 6; k++)
        for(int j = 0; j < n; j++) wr[k] += Gb[k * n + j] * xsol[static_cast<size_t>(b) * n + j];
      out[b] = sva::ForceVecd({wr[0], wr[1], wr[2]}, {wr[3], wr[4], wr[5]});
    }
    return out;
  }

  ccc_qp_ws_t * qp_ws_ = nullptr;
  int qp_n_ = 0, qp_batch_ = 0;
  std::vector<int32_t> status_;
};
} // namespace CCC
