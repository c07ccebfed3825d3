// A user of the reference's instantiation libraries (src/controllers/cartpole: cartpole_mppi): includes the instantiation
// header, defines MPPIB_USE_INSTANTIATION_LIBRARY (compile flag) and links libmppi_b200_controllers.so — the controller's member
// functions are NOT compiled in this translation unit. Exit codes: 0 = the closed loop ran and the cost went down,
// 5 = no CUDA device, other = failure.
#include <mppi/instantiations/cartpole_mppi/cartpole_mppi.cuh>
#include <mppi/instantiations/double_integrator_mppi/double_integrator_mppi.cuh>

#include <cstdio>

int main()
{
  typedef DDPFeedback<CartpoleDynamics, 100> FB;
  typedef VanillaMPPIController<CartpoleDynamics, CartpoleQuadraticCost, FB, 100, 2048> CONTROLLER;  // a pre-built one
  typedef mppi::sampling_distributions::GaussianDistribution<CartpoleDynamics::DYN_PARAMS_T> SAMPLER;
  CartpoleDynamics model(1.0, 1.0, 1.0);
  CartpoleQuadraticCost cost;
  FB fb(&model, 0.02f);
  SAMPLER::SAMPLING_PARAMS_T sp;
  sp.std_dev[0] = 5.0f;
  SAMPLER sampler(sp);
  CONTROLLER::control_trajectory init = CONTROLLER::control_trajectory::Zero();
  try
  {
    CONTROLLER controller(&model, &cost, &fb, &sampler, 0.02f, 1, 0.25f, 0.0f, 100, init);
    CartpoleDynamics::state_array x = CartpoleDynamics::state_array::Zero();
    float first = 0.0f, last = 0.0f;
    for (int i = 0; i < 50; i++)
    {
      controller.computeControl(x, 1);
      const float b = controller.getBaselineCost();
      if (i == 0)
        first = b;
      last = b;
      x = controller.getTargetStateSeq().col(1);
      controller.slideControlSequence(1);
    }
    printf("baseline first %f last %f\n", first, last);
    return (last < first) ? 0 : 1;
  }
  catch (const std::exception& e)
  {
    printf("exception: %s\n", e.what());
    return std::string(e.what()).find("no CUDA") != std::string::npos || std::string(e.what()).find("-5") != std::string::npos ? 5 : 2;
  }
}
