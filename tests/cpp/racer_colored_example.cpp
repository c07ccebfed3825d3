// Host-layer check for config C5: RacerDubinsElevationLSTMSteering + ColoredNoiseDistribution + VanillaMPPIController
// written against the reference's include paths (include/mppi/...cuh forwarders), compiled with plain g++.
// Exit codes: 0 = closed loop reaches the speed set-point, 5 = no CUDA device (expected on the CPU-only box).
#include <mppi/controllers/ColoredMPPI/colored_mppi_controller.cuh>
#include <mppi/dynamics/racer_dubins/racer_dubins_elevation_lstm_steering.cuh>
#include <mppi/sampling_distributions/colored_noise/colored_noise.cuh>
#include <mppi_b200/cost_functions/racer/racer_quadratic_cost.hpp>

#include <cmath>
#include <cstdio>
#include <random>

using DYN = RacerDubinsElevationLSTMSteering;
using SAMPLER_T = mppi::sampling_distributions::ColoredNoiseDistribution<DYN::DYN_PARAMS_T>;
struct NoFeedback
{
};

int main()
{
  {  // fail-loudly probe: no device => status -5 from the C-ABI, no fallback
    mppib_engine* probe = nullptr;
    mppib_desc d{};
    d.dynamics_id = MPPIB_DYN_RACER_LSTM;
    d.cost_id = MPPIB_COST_RACER_QUADRATIC;
    d.sampler_id = MPPIB_SAMPLER_COLORED_NOISE;
    d.num_rollouts = 64;
    d.num_timesteps = 10;
    d.num_distributions = 1;
    d.world_size = 1;
    d.model_dims[0] = 4;
    d.model_dims[1] = 20;
    int rc = mppib_create(&probe, &d);
    if (rc == MPPIB_ERR_NO_DEVICE)
    {
      printf("no CUDA device: %s\n", mppib_last_error());
      return 5;
    }
    mppib_destroy(probe);
  }
  std::vector<int> init_output_layers = { 23, 100, 8 };
  std::vector<int> output_layers = { 8, 20, 1 };
  // tests/dynamics/racer_dubins_elevation_lstm_steering_model_test.cu:26-32
  DYN model(3, 20, init_output_layers, 4, 4, output_layers, 11);
  std::array<float2, 2> rngs = { float2{ -1.0f, 1.0f }, float2{ -1.0f, 1.0f } };
  model.setControlRanges(rngs);
  // synthetic weights U(-1,1)/sqrt(fan_in); initial hidden / cell zero
  std::mt19937 gen(2);
  std::uniform_real_distribution<float> uni(-1.0f, 1.0f);
  std::vector<float> lstm(model.lstmBlock(), 0.0f), head(8 * 20 + 20 + 20 + 1);
  for (int i = 0; i < model.lstmBlock() - 8; i++)
    lstm[i] = uni(gen) / sqrtf(8.0f);
  for (auto& v : head)
    v = uni(gen) / sqrtf(8.0f);
  model.setAllValues(lstm, head);

  RacerQuadraticCost cost;
  auto cp = cost.getParams();
  cp.desired_speed = 1.2f;  // full throttle saturates near (c_t + c_0) / c_v = 1.6 m/s with the reference's defaults
  cost.setParams(cp);

  auto sp = SAMPLER_T::SAMPLING_PARAMS_T();
  for (int i = 0; i < 2; i++)
  {
    sp.std_dev[i] = 0.3f;
    sp.exponents[i] = 1.0f;
  }
  SAMPLER_T sampler(sp);

  const int T = 60;
  const float dt = 0.02f;
  try
  {
    ColoredMPPIController<DYN, RacerQuadraticCost, NoFeedback, T, 4096> ctrl(&model, &cost, nullptr, &sampler, dt,
                                                                                     1, 1.0f, 0.0f);
    DYN::state_array x = DYN::state_array::Zero(), xn, xd;
    DYN::output_array y;
    x(0) = 3.0f;
    for (int i = 0; i < 4; i++)
      x(9 + i) = 1e-6f;
    DYN::control_array u0 = DYN::control_array::Zero();
    model.initializeDynamics(x, u0, y, 0.0f, dt);
    for (int it = 0; it < 80; it++)
    {
      ctrl.computeControl(x, 1);
      DYN::control_array u = ctrl.getControlSeq().col(0);
      model.step(x, xn, xd, u, y, it, dt);
      x = xn;
      ctrl.slideControlSequence(1);
    }
    printf("speed after 80 steps %f (set-point %f), baseline %f\n", x(0), cp.desired_speed, ctrl.getBaselineCost());
    int rc = fabsf(x(0) - cp.desired_speed) < 0.3f ? 0 : 2;
    // Tsallis weights (colored_mppi_controller.cu:199-209): the engine is re-created with the control write-back buffer
    ctrl.setGamma(50.0f);
    ctrl.setRExp(2.0f);
    for (int it = 0; it < 40; it++)
    {
      ctrl.computeControl(x, 1);
      DYN::control_array u = ctrl.getControlSeq().col(0);
      model.step(x, xn, xd, u, y, it, dt);
      x = xn;
      ctrl.slideControlSequence(1);
    }
    printf("with Tsallis weights: speed %f, baseline %f, normalizer %f\n", x(0), ctrl.getBaselineCost(),
           ctrl.getNormalizerCost());
    if (!(fabsf(x(0) - cp.desired_speed) < 0.4f) || !(ctrl.getNormalizerCost() > 0.0f))
      rc = 3;
    return rc;
  }
  catch (const std::exception& e)
  {
    printf("exception: %s\n", e.what());
    return std::string(e.what()).find("no CUDA device") != std::string::npos ? 5 : 4;
  }
}
