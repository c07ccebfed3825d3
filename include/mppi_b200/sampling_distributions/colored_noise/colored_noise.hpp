/*
 * ColoredNoiseDistribution — host side of include/mppi/sampling_distributions/colored_noise/colored_noise.cuh:41-…
 * (ColoredNoiseParamsImpl :44-60: exponents per control, offset_decay_rate, fmin on top of the Gaussian parameters).
 * generateSamples (colored_noise.cu:286-392: normals -> f^(-beta/2) spectrum -> cuFFT C2R(2T) -> rearrange -> the
 * Gaussian control rewrite) runs inside the engine (mppi-generic_b200/csrc/noise_colored.cuh).
 */
#pragma once
#include "../gaussian/gaussian.hpp"

namespace mppi
{
namespace sampling_distributions
{
template <int C_DIM, int MAX_DISTRIBUTIONS_T = 2>
struct ColoredNoiseParamsImpl : public GaussianParamsImpl<C_DIM, MAX_DISTRIBUTIONS_T>
{
  float exponents[C_DIM * MAX_DISTRIBUTIONS_T] = { 0.0f };
  float offset_decay_rate = 0.97f;
  float fmin = 0.0f;
  ColoredNoiseParamsImpl(int num_rollouts = 1, int num_timesteps = 1, int num_distributions = 1)
    : GaussianParamsImpl<C_DIM, MAX_DISTRIBUTIONS_T>(num_rollouts, num_timesteps, num_distributions)
  {
  }
};

template <class DYN_PARAMS_T, int C_DIM>
class ColoredNoiseDistributionImpl
{
public:
  static const int CONTROL_DIM = C_DIM;
  static const int SAMPLER_ID = MPPIB_SAMPLER_COLORED_NOISE;
  typedef ColoredNoiseParamsImpl<C_DIM, 2> SAMPLING_PARAMS_T;
  ColoredNoiseDistributionImpl(cudaStream_t stream = 0)
  {
  }
  ColoredNoiseDistributionImpl(const SAMPLING_PARAMS_T& params, cudaStream_t stream = 0) : params_(params)
  {
    if (params_.time_specific_std_dev)
      throw std::runtime_error("time_specific_std_dev is not supported by libmppi_b200");
  }
  void setParams(const SAMPLING_PARAMS_T& params, bool /*synchronize*/ = true)
  {
    params_ = params;
  }
  SAMPLING_PARAMS_T getParams() const
  {
    return params_;
  }
  std::string getSamplingDistributionName() const
  {
    return "Colored Noise";
  }
  float getOffsetDecayRate() const
  {
    return params_.offset_decay_rate;
  }
  void setOffsetDecayRate(float v)
  {
    params_.offset_decay_rate = v;
  }
  void GPUSetup()
  {
  }
  void freeCudaMem()
  {
  }
  mppib_gaussian_params blob() const
  {
    mppib_gaussian_params b{};
    for (int i = 0; i < MPPIB_MAX_CONTROL_DIM * MPPIB_MAX_DISTRIBUTIONS; i++)
      b.std_dev[i] = 1.0f;
    for (int d = 0; d < 2; d++)
      for (int c = 0; c < C_DIM; c++)
      {
        b.std_dev[d * C_DIM + c] = params_.std_dev[d * C_DIM + c];
        b.exponents[d * C_DIM + c] = params_.exponents[d * C_DIM + c];
      }
    for (int c = 0; c < C_DIM; c++)
      b.control_cost_coeff[c] = params_.control_cost_coeff[c];
    b.pure_noise_trajectories_percentage = params_.pure_noise_trajectories_percentage;
    b.std_dev_decay = params_.std_dev_decay;
    b.sum_strides = params_.sum_strides;
    b.use_same_noise_for_all_distributions = params_.use_same_noise_for_all_distributions ? 1 : 0;
    b.offset_decay_rate = params_.offset_decay_rate;
    b.fmin = params_.fmin;
    return b;
  }

protected:
  SAMPLING_PARAMS_T params_;
};

template <class DYN_PARAMS_T>
using ColoredNoiseDistribution =
    ColoredNoiseDistributionImpl<DYN_PARAMS_T, (int)DYN_PARAMS_T::ControlIndex::NUM_CONTROLS>;
}  // namespace sampling_distributions
}  // namespace mppi
