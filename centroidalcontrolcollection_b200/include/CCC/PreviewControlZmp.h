/* CCC/PreviewControlZmp.h — drop-in host classes for CCC::PreviewControlZmp1d / CCC::PreviewControlZmp.
 *
 * Mirrors reference include/CCC/PreviewControlZmp.h and src/PreviewControlZmp.cpp: 1-D InitialParam (pos, vel,
 * acc), WeightParam (:31-50, defaults zmp 1.0, com_jerk 1e-8), planOnce (reference ZMP sampled at
 * t + (i+1) dt, src :15-29) / procOnce (src :31-49), 2-D class (:91-138, src :51-76).
 * The single-problem planOnce is a 203-term dot product per axis and runs on the host exactly like the
 * reference (BASELINE config 1 is "reference plumbing, no GPU"); planBatch sends the 2B rows of a batch
 * through the CUDA kernel behind ccc_preview_input (no CPU fallback there).
 */
#pragma once
#include <array>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/ccc_b200.h"
#include "CommonModels.h"
#include "PreviewControl.h"

namespace CCC
{
class PreviewControlZmp1d : public PreviewControl
{
public:
  using InitialParam = std::array<double, 3>;

  struct WeightParam
  {
    double zmp, com_jerk;
    WeightParam(double _zmp = 1.0, double _com_jerk = 1e-8) : zmp(_zmp), com_jerk(_com_jerk) {}
    PreviewControl::WeightParam toPreviewControlWeightParam() const
    {
      PreviewControl::WeightParam w;
      w.output = {zmp};
      w.input = {com_jerk};
      return w;
    }
  };

  PreviewControlZmp1d(double com_height, double horizon_duration, double horizon_dt, const WeightParam & weight_param = WeightParam())
  : PreviewControl(std::make_shared<ComZmpModelJerkInput>(com_height), horizon_duration, horizon_dt,
                   weight_param.toPreviewControlWeightParam())
  {
  }

  double planOnce(const std::function<double(double)> & ref_zmp_func,
                  const InitialParam & initial_param,
                  double current_time,
                  double control_dt = -1) const
  {
    std::vector<double> ref_zmp_seq(horizon_steps_);
    for(int i = 0; i < horizon_steps_; i++) ref_zmp_seq[i] = ref_zmp_func(current_time + (i + 1) * horizon_dt_);
    return procOnce(ref_zmp_seq, initial_param, current_time, control_dt);
  }

  double procOnce(const std::vector<double> & ref_zmp_seq, const InitialParam & initial_param, double, double control_dt) const
  {
    const double com_jerk = calcOptimalInput({initial_param[0], initial_param[1], initial_param[2]}, ref_zmp_seq)[0];
    return zmpFromJerk(initial_param, com_jerk, control_dt);
  }

  /** Batched procOnce on the GPU: initial_params [B], ref_zmp_seqs [B][N] row-major -> planned ZMP [B]. */
  std::vector<double> procBatch(const std::vector<InitialParam> & initial_params,
                                const std::vector<double> & ref_zmp_seqs,
                                double control_dt) const
  {
    const int B = static_cast<int>(initial_params.size()), N = horizon_steps_;
    if(ref_zmp_seqs.size() != static_cast<size_t>(B) * N) throw std::runtime_error("procBatch: ref_zmp_seqs must be [B][N]");
    std::vector<double> x(static_cast<size_t>(B) * 3), jerk(B, 0.0);
    for(int b = 0; b < B; b++)
      for(int i = 0; i < 3; i++) x[static_cast<size_t>(b) * 3 + i] = initial_params[b][i];
    const int rc = ccc_preview_input(B, N, K_.data(), F_.data(), x.data(), ref_zmp_seqs.data(), jerk.data(), CCC_MEM_HOST, nullptr);
    if(rc != CCC_OK) throw std::runtime_error(std::string("ccc_preview_input: ") + ccc_last_error());
    std::vector<double> zmp(B);
    for(int b = 0; b < B; b++) zmp[b] = zmpFromJerk(initial_params[b], jerk[b], control_dt);
    return zmp;
  }

protected:
  double zmpFromJerk(const InitialParam & ip, double com_jerk, double control_dt) const
  {
    if(control_dt < 0) control_dt = horizon_dt_;
    const double com_acc = ip[2] + control_dt * com_jerk;
    const double com_pos = ip[0] + control_dt * ip[1] + 0.5 * std::pow(control_dt, 2) * ip[2];
    return com_pos + model_->C_(0, 2) * com_acc;
  }
};

class PreviewControlZmp
{
public:
  using Vector2d = std::array<double, 2>;

  struct InitialParam
  {
    Vector2d pos = {0, 0};
    Vector2d vel = {0, 0};
    Vector2d acc = {0, 0};
  };
  struct BatchItem
  {
    int schedule = 0;
    InitialParam initial_param;
  };

  PreviewControlZmp(double com_height,
                    double horizon_duration,
                    double horizon_dt,
                    const PreviewControlZmp1d::WeightParam & weight_param = PreviewControlZmp1d::WeightParam())
  : preview_control_1d_(std::make_shared<PreviewControlZmp1d>(com_height, horizon_duration, horizon_dt, weight_param))
  {
  }

  Vector2d planOnce(const std::function<Vector2d(double)> & ref_zmp_func,
                    const InitialParam & initial_param,
                    double current_time,
                    double control_dt = -1) const
  {
    const int N = preview_control_1d_->horizon_steps_;
    std::vector<double> seq_x(N), seq_y(N);
    for(int i = 0; i < N; i++)
    {
      const Vector2d z = ref_zmp_func(current_time + (i + 1) * preview_control_1d_->horizon_dt_);
      seq_x[i] = z[0];
      seq_y[i] = z[1];
    }
    return {preview_control_1d_->procOnce(seq_x, {initial_param.pos[0], initial_param.vel[0], initial_param.acc[0]}, current_time, control_dt),
            preview_control_1d_->procOnce(seq_y, {initial_param.pos[1], initial_param.vel[1], initial_param.acc[1]}, current_time, control_dt)};
  }

  /** Batched planOnce on the GPU: items[b] follows the reference-ZMP schedule ref_zmp_funcs[items[b].schedule]. */
  std::vector<Vector2d> planBatch(const std::vector<std::function<Vector2d(double)>> & ref_zmp_funcs,
                                  const std::vector<BatchItem> & items,
                                  double current_time,
                                  double control_dt = -1) const
  {
    const int N = preview_control_1d_->horizon_steps_, S = static_cast<int>(ref_zmp_funcs.size()), B = static_cast<int>(items.size());
    if(S == 0 || B == 0) throw std::runtime_error("planBatch: empty schedule list or batch");
    for(const auto & item : items)
      if(item.schedule < 0 || item.schedule >= S) throw std::runtime_error("planBatch: schedule index out of range");
    std::vector<double> sched(static_cast<size_t>(2) * S * N);
    for(int s = 0; s < S; s++)
      for(int i = 0; i < N; i++)
      {
        const Vector2d z = ref_zmp_funcs[s](current_time + (i + 1) * preview_control_1d_->horizon_dt_);
        sched[(static_cast<size_t>(s)) * N + i] = z[0];
        sched[(static_cast<size_t>(S) + s) * N + i] = z[1];
      }
    std::vector<PreviewControlZmp1d::InitialParam> ips(2 * B);
    std::vector<double> refs(static_cast<size_t>(2) * B * N);
    for(int b = 0; b < B; b++)
      for(int a = 0; a < 2; a++)
      {
        const InitialParam & ip = items[b].initial_param;
        ips[a * B + b] = {ip.pos[a], ip.vel[a], ip.acc[a]};
        const double * src = &sched[(static_cast<size_t>(a) * S + items[b].schedule) * N];
        std::copy(src, src + N, refs.begin() + (static_cast<size_t>(a) * B + b) * N);
      }
    const std::vector<double> z = preview_control_1d_->procBatch(ips, refs, control_dt);
    std::vector<Vector2d> out(B);
    for(int b = 0; b < B; b++) out[b] = {z[b], z[B + b]};
    return out;
  }

public:
  std::shared_ptr<PreviewControlZmp1d> preview_control_1d_;
};
} // namespace CCC
