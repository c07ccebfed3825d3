/*
 * NLNDistribution — host side of include/mppi/sampling_distributions/nln/nln.cuh:20-74 (normal x log-normal noise with
 * the Gaussian parameters and the Gaussian control rewrite / likelihood-ratio cost). generateSamples (nln.cu:107-165:
 * CONTROL_DIM curandGenerateLogNormal planes, one curandGenerateNormal block, createNLNNoise) runs inside the engine
 * (mppi-generic_b200/csrc/engine.cu gen_draw, csrc/noise_colored.cuh nln_combine_kernel).
 */
#pragma once
#include <cmath>
#include <vector>

#include "../gaussian/gaussian.hpp"

namespace mppi
{
namespace sampling_distributions
{
template <class DYN_PARAMS_T, int C_DIM>
class NLNDistributionImpl : public GaussianDistributionImpl<DYN_PARAMS_T, C_DIM>
{
public:
  typedef GaussianDistributionImpl<DYN_PARAMS_T, C_DIM> PARENT_CLASS;
  typedef typename PARENT_CLASS::SAMPLING_PARAMS_T SAMPLING_PARAMS_T;
  static const int SAMPLER_ID = MPPIB_SAMPLER_NLN;
  NLNDistributionImpl(cudaStream_t stream = 0) : PARENT_CLASS(stream)
  {
    calculateLogMeanAndVariance();
  }
  NLNDistributionImpl(const SAMPLING_PARAMS_T& params, cudaStream_t stream = 0) : PARENT_CLASS(params, stream)
  {
    calculateLogMeanAndVariance();
  }
  std::string getSamplingDistributionName() const
  {
    return "NLN";
  }
  void setParams(const SAMPLING_PARAMS_T& params, bool synchronize = true)
  {
    PARENT_CLASS::setParams(params, synchronize);
    calculateLogMeanAndVariance();
  }
  // nln.cu:93-105 (informational: the engine draws the log-normal factors from std_dev directly, like the reference)
  void calculateLogMeanAndVariance()
  {
    log_noise_mean_.resize(C_DIM * 2);
    log_noise_std_dev_.resize(C_DIM * 2);
    for (int i = 0; i < C_DIM * 2; i++)
    {
      const float normal_variance = this->params_.std_dev[i] * this->params_.std_dev[i];
      log_noise_mean_[i] = expf(0.5 * normal_variance);
      const float log_variance = expf(normal_variance) * expf(normal_variance - 1.0f);
      log_noise_std_dev_[i] = sqrtf(log_variance);
    }
  }
  const std::vector<float>& getLogNoiseMean() const
  {
    return log_noise_mean_;
  }
  const std::vector<float>& getLogNoiseStdDev() const
  {
    return log_noise_std_dev_;
  }

protected:
  std::vector<float> log_noise_mean_;
  std::vector<float> log_noise_std_dev_;
};

template <class DYN_PARAMS_T>
using NLNDistribution = NLNDistributionImpl<DYN_PARAMS_T, (int)DYN_PARAMS_T::ControlIndex::NUM_CONTROLS>;
}  // namespace sampling_distributions
}  // namespace mppi
