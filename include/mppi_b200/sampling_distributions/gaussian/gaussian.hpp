/*
 * GaussianDistribution — host side of include/mppi/sampling_distributions/gaussian/gaussian.cuh:21-61 (params) and
 * sampling_distribution.cuh:14-29. The N x T x C sample buffer, the cuRAND-identical draw and the weighted mean update
 * (generateSamples / updateDistributionParamsFromDevice, gaussian.cu:375-457) live in the engine.
 */
#pragma once
#include <string>
#include "../../utils/common.hpp"

namespace mppi
{
namespace sampling_distributions
{
template <int C_DIM, int MAX_DISTRIBUTIONS_T = 2>
struct GaussianParamsImpl
{
  static const int CONTROL_DIM = C_DIM;
  static const int MAX_DISTRIBUTIONS = MAX_DISTRIBUTIONS_T;
  bool use_same_noise_for_all_distributions = true;
  int num_rollouts = 1;
  int num_timesteps = 1;
  int num_distributions = 1;
  float std_dev[C_DIM * MAX_DISTRIBUTIONS_T];
  float control_cost_coeff[C_DIM];
  float pure_noise_trajectories_percentage = 0.01f;
  float std_dev_decay = 1.0f;
  dim3 rewrite_controls_block_dim = dim3(32, 16, 1);
  int sum_strides = 32;
  bool time_specific_std_dev = false;
  GaussianParamsImpl(int num_rollouts = 1, int num_timesteps = 1, int num_distributions = 1)
    : num_rollouts(num_rollouts), num_timesteps(num_timesteps), num_distributions(num_distributions)
  {
    for (int i = 0; i < C_DIM * MAX_DISTRIBUTIONS_T; i++)
      std_dev[i] = 1.0f;
    for (int i = 0; i < C_DIM; i++)
      control_cost_coeff[i] = 0.0f;
  }
};

template <class DYN_PARAMS_T, int C_DIM>
class GaussianDistributionImpl
{
public:
  static const int CONTROL_DIM = C_DIM;
  static const int SAMPLER_ID = MPPIB_SAMPLER_GAUSSIAN;
  typedef GaussianParamsImpl<C_DIM, 2> SAMPLING_PARAMS_T;
  GaussianDistributionImpl(cudaStream_t stream = 0)
  {
  }
  GaussianDistributionImpl(const SAMPLING_PARAMS_T& params, cudaStream_t stream = 0) : params_(params)
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
  {  // gaussian.cuh
    return "Gaussian";
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
        b.std_dev[d * C_DIM + c] = params_.std_dev[d * C_DIM + c];
    for (int c = 0; c < C_DIM; c++)
      b.control_cost_coeff[c] = params_.control_cost_coeff[c];
    b.pure_noise_trajectories_percentage = params_.pure_noise_trajectories_percentage;
    b.std_dev_decay = params_.std_dev_decay;
    b.sum_strides = params_.sum_strides;
    b.use_same_noise_for_all_distributions = params_.use_same_noise_for_all_distributions ? 1 : 0;
    b.offset_decay_rate = 0.97f;
    return b;
  }

protected:
  SAMPLING_PARAMS_T params_;
};

// GaussianDistribution<DYN_PARAMS_T>: CONTROL_DIM comes from the dynamics params' ControlIndex enum, as in the reference
// (sampling_distribution.cuh:36-40).
template <class DYN_PARAMS_T>
using GaussianDistribution = GaussianDistributionImpl<DYN_PARAMS_T, (int)DYN_PARAMS_T::ControlIndex::NUM_CONTROLS>;
}  // namespace sampling_distributions
}  // namespace mppi
