// Host-layer check: the reference's examples/cartpole_example.cu, line for line where the API allows, compiled against
// the header-only layer (include/mppi_b200/) and linked to libmppi_b200.so with plain g++ (no nvcc, no Eigen needed).
// Exit codes: 0 = swing-up cost below threshold, 5 = no CUDA device (expected on the CPU-only box), other = failure.
#include <mppi_b200/controllers/MPPI/mppi_controller.hpp>
#include <mppi_b200/controllers/Tube-MPPI/tube_mppi_controller.hpp>
#include <mppi/controllers/R-MPPI/robust_mppi_controller.cuh>
#include <mppi/sampling_distributions/nln/nln.cuh>
#include <mppi_b200/cost_functions/cartpole/cartpole_quadratic_cost.hpp>
#include <mppi_b200/cost_functions/double_integrator/double_integrator_circle_cost.hpp>
#include <mppi_b200/dynamics/cartpole/cartpole_dynamics.hpp>
#include <mppi_b200/dynamics/double_integrator/di_dynamics.hpp>

#include <chrono>
#include <cmath>
#include <iostream>

using SAMPLER_T = mppi::sampling_distributions::GaussianDistribution<CartpoleDynamics::DYN_PARAMS_T>;
struct NoFeedback
{
};

int main(int argc, char** argv)
{
  {  // fail-loudly probe: no device => status -5 from the C-ABI, no fallback
    mppib_engine* probe = nullptr;
    mppib_desc d{};
    d.dynamics_id = MPPIB_DYN_CARTPOLE;
    d.cost_id = MPPIB_COST_CARTPOLE_QUADRATIC;
    d.num_rollouts = 64;
    d.num_timesteps = 10;
    d.num_distributions = 1;
    d.world_size = 1;
    int rc = mppib_create(&probe, &d);
    if (rc == MPPIB_ERR_NO_DEVICE)
    {
      printf("no CUDA device: %s\n", mppib_last_error());
      return 5;
    }
    mppib_destroy(probe);
  }
  auto model = new CartpoleDynamics(1.0, 1.0, 1.0);
  auto cost = new CartpoleQuadraticCost;
  // parameters of the reference's swing-up integration test (tests/controllers/vanilla_mppi_test.cu:79-136): with them
  // the baseline cost must drop below 1.0 within 1000 control steps
  model->control_rngs_->x = -1e30f;  // written through the public member like examples/cartpole_example.cu:12-13
  model->control_rngs_->y = 1e30f;

  CartpoleQuadraticCostParams new_params;
  new_params.cart_position_coeff = 100;
  new_params.pole_angle_coeff = 200;
  new_params.cart_velocity_coeff = 10;
  new_params.pole_angular_velocity_coeff = 20;
  new_params.control_cost_coeff[0] = 1;
  new_params.terminal_cost_coeff = 0;
  new_params.desired_terminal_state[0] = -20;
  new_params.desired_terminal_state[1] = 0;
  new_params.desired_terminal_state[2] = M_PI;
  new_params.desired_terminal_state[3] = 0;
  cost->setParams(new_params);

  float dt = 0.01;
  int max_iter = 1;
  float lambda = 0.25;
  float alpha = 0.01;
  const int num_timesteps = 100;

  auto sampler_params = SAMPLER_T::SAMPLING_PARAMS_T();
  for (int i = 0; i < CartpoleDynamics::CONTROL_DIM; i++)
  {
    sampler_params.std_dev[i] = 5.0;
    sampler_params.control_cost_coeff[i] = 1.0;
  }
  sampler_params.pure_noise_trajectories_percentage = 0.01f;
  auto sampler = new SAMPLER_T(sampler_params);
  NoFeedback* fb_controller = nullptr;

  auto CartpoleController = new VanillaMPPIController<CartpoleDynamics, CartpoleQuadraticCost, NoFeedback, num_timesteps, 2048>(
      model, cost, fb_controller, sampler, dt, max_iter, lambda, alpha);
  auto controller_params = CartpoleController->getParams();
  controller_params.dynamics_rollout_dim_ = dim3(64, 4, 1);
  controller_params.cost_rollout_dim_ = dim3(64, 4, 1);
  controller_params.seed_ = 42;
  controller_params.slide_control_scale_[0] = 1.0;
  CartpoleController->setParams(controller_params);

  CartpoleDynamics::state_array current_state = CartpoleDynamics::state_array::Zero();
  CartpoleDynamics::state_array next_state = CartpoleDynamics::state_array::Zero();
  CartpoleDynamics::output_array output = CartpoleDynamics::output_array::Zero();
  CartpoleDynamics::state_array xdot = CartpoleDynamics::state_array::Zero();
  int time_horizon = 1000;
  auto time_start = std::chrono::system_clock::now();
  for (int i = 0; i < time_horizon; ++i)
  {
    CartpoleController->computeControl(current_state, 1);
    CartpoleDynamics::control_array control;
    control = CartpoleController->getControlSeq().block(0, 0, CartpoleDynamics::CONTROL_DIM, 1);
    model->enforceConstraints(current_state, control);
    model->step(current_state, next_state, xdot, control, output, i, dt);
    current_state = next_state;
    if (i % 250 == 0)
    {
      printf("Current Time: %f    ", i * dt);
      printf("Current Baseline Cost: %f    ", CartpoleController->getBaselineCost());
      model->printState(current_state.data());
    }
    CartpoleController->slideControlSequence(1);
  }
  auto diff = std::chrono::duration<double, std::milli>(std::chrono::system_clock::now() - time_start);
  printf("The elapsed time is: %f milliseconds (%f solves/s)\n", diff.count(), 1000.0 * time_horizon / diff.count());
  {  // host-only helpers of the Controller base (controller.cuh:236-261,317-378)
    printf("controller: %s\n", CartpoleController->getFullName().c_str());
    auto seq = CartpoleController->getControlSeq();
    CartpoleDynamics::state_array target = CartpoleDynamics::state_array::Zero();
    CartpoleDynamics::control_array u_mid = CartpoleController->getCurrentControl(current_state, 0.5 * dt, target, seq);
    CartpoleDynamics::control_array lo = seq.col(0), hi = seq.col(1);
    model->enforceConstraints(current_state, lo);
    const float expect = 0.5f * (seq(0, 0) + seq(0, 1));
    if (fabsf(u_mid(0) - expect) > 1e-5f * (1.0f + fabsf(expect)) ||
        CartpoleController->getFullName() != "Vanilla MPPI(Cartpole, Cartpole quadratic cost, Gaussian)")
    {
      printf("getCurrentControl / getFullName mismatch: %f vs %f\n", u_mid(0), expect);
      return 7;
    }
  }
  const float pole_err = fabsf(fabsf(current_state(2)) - (float)M_PI);
  printf("final pole angle error %f, baseline %f\n", pole_err, CartpoleController->getBaselineCost());
  int rc = (CartpoleController->getBaselineCost() < 1.0f && pole_err < 0.3f) ? 0 : 2;  // EXPECT_LT(baseline, 1.0)

  // Tube-MPPI on the double integrator: just exercise the two-system path through the C++ layer
  {
    using DI = DoubleIntegratorDynamics;
    DI di_model(1.0f);
    DoubleIntegratorCircleCost di_cost;
    using DS = mppi::sampling_distributions::GaussianDistribution<DI::DYN_PARAMS_T>;
    DS di_sampler;
    TubeMPPIController<DI, DoubleIntegratorCircleCost, NoFeedback, 50, 1024> tube(&di_model, &di_cost, nullptr, &di_sampler,
                                                                                  0.02f, 1, 2.0f, 0.0f);
    DI::state_array x;
    x << 2, 0, 0, 1;
    for (int t = 0; t < 50; t++)
    {
      tube.computeControl(x, 1);
      DI::control_array u = tube.getControlSeq().col(0);
      DI::state_array xn, xd;
      DI::output_array y;
      di_model.step(x, xn, xd, u, y, t, 0.02f);
      x = xn;
      tube.slideControlSequence(1);
    }
    const float r = sqrtf(x(0) * x(0) + x(1) * x(1));
    printf("tube: radius after 50 steps %f, baselines %f / %f\n", r, tube.getBaselineCost(0), tube.getBaselineCost(1));
    if (!(r > 1.675f && r < 2.325f))
      rc = 3;
  }
  // Robust MPPI on the same system: init-eval line search + feedback in the rollout, through the reference's include path
  {
    using DI = DoubleIntegratorDynamics;
    DI di_model(1.0f);
    DoubleIntegratorCircleCost di_cost;
    using DS = mppi::sampling_distributions::GaussianDistribution<DI::DYN_PARAMS_T>;
    DS di_sampler;
    using RMPPI = RobustMPPIController<DI, DoubleIntegratorCircleCost, NoFeedback, 50, 2048>;
    RMPPI rmppi(&di_model, &di_cost, nullptr, &di_sampler, 0.02f, 1, 2.0f, 0.0f, 20.0f);
    RMPPI::feedback_gain_matrix K = RMPPI::feedback_gain_matrix::Zero();
    K(0, 0) = K(1, 1) = -4.0f;
    K(0, 2) = K(1, 3) = -2.0f;
    rmppi.setFeedbackGains(std::vector<RMPPI::feedback_gain_matrix>(50, K));
    DI::state_array x;
    x << 2, 0, 0, 1;
    for (int t = 0; t < 80; t++)
    {
      rmppi.updateImportanceSamplingControl(x, 1);
      rmppi.computeControl(x, 1);
      DI::state_array xnom = rmppi.getNominalStateSeq().col(0), e = x;
      for (int i = 0; i < 4; i++)
        e(i) = x(i) - xnom(i);
      DI::control_array u = rmppi.getNominalControlSeq().col(0);
      for (int c = 0; c < 2; c++)
        for (int i = 0; i < 4; i++)
          u(c) += K(c, i) * e(i);
      DI::state_array xn, xd;
      DI::output_array y;
      di_model.step(x, xn, xd, u, y, t, 0.02f);
      x = xn;
    }
    const float r = sqrtf(x(0) * x(0) + x(1) * x(1));
    printf("rmppi: radius after 80 steps %f, baselines nominal %f / real %f, candidate used %d\n", r,
           rmppi.getBaselineCost(0), rmppi.getBaselineCost(1), rmppi.getBestIndex());
    if (!(r > 1.675f && r < 2.325f))
      rc = 4;
  }
  // NLN (log-MPPI) sampler behind VanillaMPPI on the double integrator, through the reference's include path
  {
    using DI = DoubleIntegratorDynamics;
    DI di_model(1.0f);
    DoubleIntegratorCircleCost di_cost;
    using NS = mppi::sampling_distributions::NLNDistribution<DI::DYN_PARAMS_T>;
    auto np = NS::SAMPLING_PARAMS_T();
    np.std_dev[0] = np.std_dev[1] = 0.8f;
    NS nln_sampler(np);
    VanillaMPPIController<DI, DoubleIntegratorCircleCost, NoFeedback, 64, 2048, NS> nln(&di_model, &di_cost, nullptr,
                                                                                      &nln_sampler, 0.02f, 1, 2.0f, 0.0f);
    DI::state_array x;
    x << 2, 0, 0, 1;
    for (int t = 0; t < 100; t++)
    {
      nln.computeControl(x, 1);
      DI::control_array u = nln.getControlSeq().col(0);
      DI::state_array xn, xd;
      DI::output_array y;
      di_model.step(x, xn, xd, u, y, t, 0.02f);
      x = xn;
      nln.slideControlSequence(1);
    }
    const float r = sqrtf(x(0) * x(0) + x(1) * x(1));
    printf("nln: radius after 100 steps %f, baseline %f (log-noise mean %f)\n", r, nln.getBaselineCost(),
           nln_sampler.getLogNoiseMean()[0]);
    if (!(r > 1.675f && r < 2.325f))
      rc = 6;
  }
  delete CartpoleController;
  return rc;
}
