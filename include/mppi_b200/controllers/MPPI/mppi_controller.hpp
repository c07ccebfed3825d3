/*
 * VanillaMPPIController — host class of include/mppi/controllers/MPPI/mppi_controller.cuh:14-17 (same template
 * parameter order, same constructors). computeControl follows controllers/MPPI/mppi_controller.cu:151-241 with the
 * device pipeline collapsed into one engine solve per optimisation iteration.
 */
#pragma once
#include "../controller.hpp"
#include "../../sampling_distributions/gaussian/gaussian.hpp"

template <class DYN_T, class COST_T, class FB_T, int MAX_TIMESTEPS, int NUM_ROLLOUTS,
          class SAMPLING_T = ::mppi::sampling_distributions::GaussianDistribution<typename DYN_T::DYN_PARAMS_T>,
          class PARAMS_T = ControllerParams<DYN_T::STATE_DIM, DYN_T::CONTROL_DIM, MAX_TIMESTEPS>>
class VanillaMPPIController
  : public Controller<DYN_T, COST_T, FB_T, SAMPLING_T, MAX_TIMESTEPS, NUM_ROLLOUTS, PARAMS_T, 1>
{
public:
  typedef Controller<DYN_T, COST_T, FB_T, SAMPLING_T, MAX_TIMESTEPS, NUM_ROLLOUTS, PARAMS_T, 1> PARENT_CLASS;
  using control_array = typename PARENT_CLASS::control_array;
  using control_trajectory = typename PARENT_CLASS::control_trajectory;
  using state_trajectory = typename PARENT_CLASS::state_trajectory;
  using state_array = typename PARENT_CLASS::state_array;
  using output_array = typename PARENT_CLASS::output_array;

  VanillaMPPIController(DYN_T* model, COST_T* cost, FB_T* fb_controller, SAMPLING_T* sampler, float dt, int max_iter,
                        float lambda, float alpha, int num_timesteps = MAX_TIMESTEPS,
                        const Eigen::Ref<const control_trajectory>& init_control_traj = control_trajectory::Zero(),
                        cudaStream_t stream = nullptr)
    : PARENT_CLASS(model, cost, fb_controller, sampler, dt, max_iter, lambda, alpha, num_timesteps, init_control_traj,
                   stream)
  {
    chooseAppropriateKernel();
  }
  VanillaMPPIController(DYN_T* model, COST_T* cost, FB_T* fb_controller, SAMPLING_T* sampler, PARAMS_T& params,
                        cudaStream_t stream = nullptr)
    : PARENT_CLASS(model, cost, fb_controller, sampler, params, stream)
  {
    chooseAppropriateKernel();
  }
  std::string getControllerName() override
  {
    return "Vanilla MPPI";
  }
  void chooseAppropriateKernel() override
  {
    PARENT_CLASS::chooseAppropriateKernel();
  }

  void computeControl(const Eigen::Ref<const state_array>& state, int optimization_stride = 1) override
  {
    this->free_energy_statistics_.real_sys.previousBaseline = this->getBaselineCost();
    state_array x0 = state;
    control_trajectory u_nominal = this->control_;
    for (int opt_iter = 0; opt_iter < this->getNumIters(); opt_iter++)
    {
      control_trajectory u_out = control_trajectory::Zero();
      u_nominal = this->control_;
      this->solve(x0.data(), this->control_.data(), optimization_stride, opt_iter, u_out.data());
      this->control_ = u_out;
    }
    if (this->getTotalSampledTrajectories() > 0)  // mppi_controller.cu:232-240
      this->pickSampledControls(x0, u_nominal, this->control_);
    this->free_energy_statistics_.real_sys.normalizerPercent = this->getNormalizerCost() / NUM_ROLLOUTS;
    this->free_energy_statistics_.real_sys.increase =
        this->getBaselineCost() - this->free_energy_statistics_.real_sys.previousBaseline;
    if (this->device_side_tail_)
      this->deviceSideTail(state, this->control_, this->control_history_, this->state_, this->output_);
    else
    {
      smoothControlTrajectory();
      computeStateTrajectory(state);
    }
    state_array zero_state = this->model_->getZeroState();
    for (int i = 0; i < this->getNumTimesteps(); i++)
    {  // mppi_controller.cu:227-231
      control_array u = this->control_.col(i);
      this->model_->enforceConstraints(zero_state, u);
      this->control_.col(i) = u;
    }
  }
  void slideControlSequence(int steps) override
  {  // mppi_controller.cuh: save history then slide
    this->saveControlHistoryHelper(steps, this->control_, this->control_history_);
    this->slideControlSequenceHelper(steps, this->control_);
  }
  void computeStateTrajectory(const Eigen::Ref<const state_array>& x0)
  {
    this->computeOutputTrajectoryHelper(this->output_, this->state_, x0, this->control_);
  }
  void smoothControlTrajectory()
  {
    this->smoothControlTrajectoryHelper(this->control_, this->control_history_);
  }
};
