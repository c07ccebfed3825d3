// probe TU (step 1): does the unmodified reference controller instantiate with the Eigen stand-in?
#include <mppi/controllers/MPPI/mppi_controller.cuh>
#include <mppi/dynamics/cartpole/cartpole_dynamics.cuh>
#include <mppi/cost_functions/cartpole/cartpole_quadratic_cost.cuh>

// A feedback controller that does nothing (the reference's DDP needs Eigen decompositions; the hot path never calls it:
// computeControl() does not compute feedback, controller.cuh:546-549 is called by the plant only)
template <class DYN_T>
class NullGPUFeedback : public GPUFeedbackController<NullGPUFeedback<DYN_T>, DYN_T, GPUState>
{
public:
  NullGPUFeedback(cudaStream_t stream = 0) : GPUFeedbackController<NullGPUFeedback<DYN_T>, DYN_T, GPUState>(stream)
  {
  }
};
struct NullFeedbackParams
{
};
template <class DYN_T, int T>
class NullFeedback : public FeedbackController<NullGPUFeedback<DYN_T>, NullFeedbackParams, T>
{
public:
  typedef FeedbackController<NullGPUFeedback<DYN_T>, NullFeedbackParams, T> PARENT;
  using state_array = typename PARENT::state_array;
  using control_array = typename PARENT::control_array;
  using state_trajectory = typename PARENT::state_trajectory;
  using control_trajectory = typename PARENT::control_trajectory;
  using FB_STATE = typename PARENT::TEMPLATED_FEEDBACK_STATE;
  NullFeedback(DYN_T* = nullptr, float dt = 0.01f, int num_timesteps = T, cudaStream_t stream = 0) : PARENT(dt, num_timesteps, stream)
  {
  }
  void initTrackingController() override
  {
  }
  control_array k_(const Eigen::Ref<const state_array>&, const Eigen::Ref<const state_array>&, int, FB_STATE&) override
  {
    return control_array::Zero();
  }
  void computeFeedback(const Eigen::Ref<const state_array>&, const Eigen::Ref<const state_trajectory>&,
                       const Eigen::Ref<const control_trajectory>&) override
  {
  }
};

typedef VanillaMPPIController<CartpoleDynamics, CartpoleQuadraticCost, NullFeedback<CartpoleDynamics, 100>, 100, 2048> CartpoleCtl;
template class VanillaMPPIController<CartpoleDynamics, CartpoleQuadraticCost, NullFeedback<CartpoleDynamics, 100>, 100, 2048>;
int main() { return 0; }
