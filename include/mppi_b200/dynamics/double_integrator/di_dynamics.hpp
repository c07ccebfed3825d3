/* DoubleIntegratorDynamics — host class of include/mppi/dynamics/double_integrator/di_dynamics.cuh. */
#pragma once
#include <cmath>
#include <random>

#include "../dynamics.hpp"

struct DoubleIntegratorParams
{  // di_dynamics.cuh:9-37
  enum class StateIndex : int { POS_X = 0, POS_Y, VEL_X, VEL_Y, NUM_STATES };
  enum class ControlIndex : int { ACCEL_X = 0, ACCEL_Y, NUM_CONTROLS };
  enum class OutputIndex : int { POS_X = 0, POS_Y, VEL_X, VEL_Y, NUM_OUTPUTS };
  float system_noise = 1;
  DoubleIntegratorParams() = default;
  DoubleIntegratorParams(float noise) : system_noise(noise){};
};

class DoubleIntegratorDynamics
  : public MPPI_internal::Dynamics<DoubleIntegratorDynamics, mppib_di_dyn_params, MPPIB_DYN_DOUBLE_INTEGRATOR, 4, 2, 4>
{
public:
  typedef DoubleIntegratorParams DYN_PARAMS_T;
  DoubleIntegratorDynamics(float system_noise = 1.0f, cudaStream_t stream = 0)
  {
    params_ = DoubleIntegratorParams(system_noise);
    std::random_device rd;
    gen.seed(rd());  // di_dynamics.cu:8-11
    setStateVariance(system_noise);
  }
  DoubleIntegratorParams getParams() const
  {
    return params_;
  }
  void setStateVariance(float system_variance = 1.0)
  {
    params_.system_noise = system_variance;
    normal_distribution = std::normal_distribution<float>(0, sqrtf(system_variance));
  }
  void computeStateDisturbance(float dt, Eigen::Ref<state_array> state)
  {  // di_dynamics.cu:59-65
    state(2) += normal_distribution(gen) * dt;
    state(3) += normal_distribution(gen) * dt;
  }
  std::string getDynamicsModelName() const override
  {
    return "2D Double Integrator Model";
  }
  void printState(const float* s) const
  {
    printf("X position: %f; Y position: %f; X velocity: %f; Y velocity: %f \n", s[0], s[1], s[2], s[3]);
  }
  mppib_di_dyn_params modelBlob() const
  {
    mppib_di_dyn_params b{};
    b.system_noise = params_.system_noise;
    return b;
  }

protected:
  DoubleIntegratorParams params_;
  std::mt19937 gen;
  std::normal_distribution<float> normal_distribution;
};
