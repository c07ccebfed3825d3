// Host-layer check for the Quadrotor pair through the reference's own include path
// (instantiations/quadrotor_mppi/quadrotor_mppi.cuh), compiled with plain g++.
// Exit codes: 0 = the vehicle reaches the goal, 5 = no CUDA device (expected on the CPU-only box).
#include <mppi/instantiations/quadrotor_mppi/quadrotor_mppi.cuh>

#include <cmath>
#include <cstdio>

using DYN = QuadrotorDynamics;
using COST = QuadrotorQuadraticCost;
using SAMPLER_T = mppi::sampling_distributions::GaussianDistribution<DYN::DYN_PARAMS_T>;
struct NoFeedback
{
};

int main()
{
  {  // fail-loudly probe: no device => status -5 from the C-ABI, no fallback
    mppib_engine* probe = nullptr;
    mppib_desc d{};
    d.dynamics_id = MPPIB_DYN_QUADROTOR;
    d.cost_id = MPPIB_COST_QUADROTOR_QUADRATIC;
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
  DYN model;  // thrust in [0, 36], zero control = hover thrust
  std::array<float2, 4> rngs = { float2{ -3.0f, 3.0f }, float2{ -3.0f, 3.0f }, float2{ -3.0f, 3.0f },
                                 float2{ 0.0f, 36.0f } };
  model.setControlRanges(rngs);
  COST cost;
  auto cp = cost.getParams();
  cp.x_goal()[0] = 4.0f;
  cp.x_goal()[1] = 1.0f;
  cp.x_goal()[2] = 2.0f;
  cp.x_coeff = 10.0f;
  cp.w_coeff = 0.5f;
  cp.roll_coeff = cp.pitch_coeff = cp.yaw_coeff = 5.0f;
  cost.setParams(cp);

  auto sp = SAMPLER_T::SAMPLING_PARAMS_T();
  const float sd[4] = { 0.5f, 0.5f, 0.5f, 2.0f };
  const float cc[4] = { 0.1f, 0.1f, 0.1f, 0.01f };
  for (int i = 0; i < 4; i++)
  {
    sp.std_dev[i] = sd[i];
    sp.control_cost_coeff[i] = cc[i];
  }
  SAMPLER_T sampler(sp);

  const int T = 75;
  const float dt = 0.02f;
  typedef VanillaMPPIController<DYN, COST, NoFeedback, T, 4096> CTRL;
  CTRL::control_trajectory init = CTRL::control_trajectory::Zero();
  for (int t = 0; t < T; t++)
    init(3, t) = model.zero_control_[3];
  try
  {
    CTRL ctrl(&model, &cost, nullptr, &sampler, dt, 1, 1.0f, 0.0f, T, init);
    DYN::state_array x = model.getZeroState(), xn, xd;
    DYN::output_array y;
    for (int it = 0; it < 250; it++)
    {
      ctrl.computeControl(x, 1);
      DYN::control_array u = ctrl.getControlSeq().col(0);
      model.enforceConstraints(x, u);
      model.step(x, xn, xd, u, y, it, dt);
      x = xn;
      ctrl.slideControlSequence(1);
    }
    // sampled (visualisation) trajectories of the last solve: 1 % random + the top 4 by weight
    ctrl.setPercentageSampledControlTrajectories(0.01f);
    ctrl.setTopNSampledControlTrajectories(4);
    ctrl.computeControl(x, 1);
    ctrl.calculateSampledStateTrajectories();
    auto trajs = ctrl.getSampledOutputTrajectories();
    auto ctrajs = ctrl.getSampledCostTrajectories();
    auto crashes = ctrl.getSampledCrashStatusTrajectories();
    auto top = ctrl.getTopNCosts();
    auto idx = ctrl.getSampledIndices();
    auto all_costs = ctrl.getSampledCostSeq();
    bool vis_ok = (int)trajs.size() == ctrl.getTotalSampledTrajectories() && (int)trajs.size() == 40 + 4 && idx[0] == -1 &&
                  top.size() == 4 && fabsf(top[0] * ctrl.getNormalizerCost() - 1.0f) < 1e-5f;
    for (size_t i = trajs.size() - 4; i < trajs.size() && vis_ok; i++)
    {  // a stored rollout's per-step costs sum to its trajectory cost
      double sum = 0;
      for (int t = 0; t <= T; t++)
        sum += ctrajs[i](t);
      vis_ok = fabs(sum - all_costs(idx[i])) <= 1e-4 * fmax(1.0, fabs(all_costs(idx[i]))) && crashes[i][T - 1] == 0;
    }
    printf("sampled trajectories: %zu (top weight %f), consistent %d\n", trajs.size(), top[0], (int)vis_ok);
    if (!vis_ok)
      return 3;
    const float dist = sqrtf((x(0) - 4) * (x(0) - 4) + (x(1) - 1) * (x(1) - 1) + (x(2) - 2) * (x(2) - 2));
    const float qn = sqrtf(x(6) * x(6) + x(7) * x(7) + x(8) * x(8) + x(9) * x(9));
    printf("distance to goal after 250 steps %f, |q| %f, baseline %f\n", dist, qn, ctrl.getBaselineCost());
    model.printState(x.data());
    return (dist < 0.5f && fabsf(qn - 1.0f) < 1e-4f) ? 0 : 1;
  }
  catch (const std::exception& e)
  {
    printf("exception: %s\n", e.what());
    return 2;
  }
}
