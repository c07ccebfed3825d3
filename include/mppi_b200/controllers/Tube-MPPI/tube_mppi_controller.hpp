/*
 * TubeMPPIController — host class of include/mppi/controllers/Tube-MPPI/tube_mppi_controller.cuh. Two systems (actual,
 * nominal) share one noise draw and run as ONE engine solve with num_distributions = 2
 * (tube_mppi_controller.cu:157-299; nominal swap rule :262-280; slide :315-325).
 */
#pragma once
#include <cstring>

#include "../controller.hpp"
#include "../../sampling_distributions/gaussian/gaussian.hpp"

template <int S_DIM, int C_DIM, int MAX_TIMESTEPS>
struct TubeMPPIParams : public ControllerParams<S_DIM, C_DIM, MAX_TIMESTEPS>
{
  float nominal_threshold_ = 20;  // tube_mppi_controller.cuh
};

template <class DYN_T, class COST_T, class FB_T, int MAX_TIMESTEPS, int NUM_ROLLOUTS,
          class SAMPLING_T = ::mppi::sampling_distributions::GaussianDistribution<typename DYN_T::DYN_PARAMS_T>,
          class PARAMS_T = TubeMPPIParams<DYN_T::STATE_DIM, DYN_T::CONTROL_DIM, MAX_TIMESTEPS>>
class TubeMPPIController : public Controller<DYN_T, COST_T, FB_T, SAMPLING_T, MAX_TIMESTEPS, NUM_ROLLOUTS, PARAMS_T, 2>
{
public:
  typedef Controller<DYN_T, COST_T, FB_T, SAMPLING_T, MAX_TIMESTEPS, NUM_ROLLOUTS, PARAMS_T, 2> PARENT_CLASS;
  using control_array = typename PARENT_CLASS::control_array;
  using control_trajectory = typename PARENT_CLASS::control_trajectory;
  using state_trajectory = typename PARENT_CLASS::state_trajectory;
  using output_trajectory = typename PARENT_CLASS::output_trajectory;
  using state_array = typename PARENT_CLASS::state_array;
  using output_array = typename PARENT_CLASS::output_array;

  TubeMPPIController(DYN_T* model, COST_T* cost, FB_T* fb_controller, SAMPLING_T* sampler, float dt, int max_iter,
                     float lambda, float alpha, int num_timesteps = MAX_TIMESTEPS,
                     const Eigen::Ref<const control_trajectory>& init_control_traj = control_trajectory::Zero(),
                     cudaStream_t stream = nullptr)
    : PARENT_CLASS(model, cost, fb_controller, sampler, dt, max_iter, lambda, alpha, num_timesteps, init_control_traj,
                   stream)
  {
    nominal_control_trajectory_ = init_control_traj;
    this->chooseAppropriateKernel();
  }
  TubeMPPIController(DYN_T* model, COST_T* cost, FB_T* fb_controller, SAMPLING_T* sampler, PARAMS_T& params,
                     cudaStream_t stream = nullptr)
    : PARENT_CLASS(model, cost, fb_controller, sampler, params, stream)
  {
    nominal_control_trajectory_ = this->params_.init_control_traj_;
    this->chooseAppropriateKernel();
  }
  std::string getControllerName() override
  {
    return "Tube MPPI";
  }
  // tube_mppi_controller.cuh:49-63: the solution is the nominal system
  control_trajectory getControlSeq() const override
  {
    return nominal_control_trajectory_;
  }
  state_trajectory getTargetStateSeq() const override
  {
    return nominal_state_trajectory_;
  }
  control_trajectory getActualControlSeq()
  {
    return this->control_;
  }
  state_trajectory getActualStateSeq()
  {
    return this->state_;
  }
  float getNominalThreshold() const
  {
    return this->params_.nominal_threshold_;
  }
  void setNominalThreshold(float threshold)
  {
    this->params_.nominal_threshold_ = threshold;
  }

  void computeControl(const Eigen::Ref<const state_array>& state, int optimization_stride = 1) override
  {
    constexpr int S = DYN_T::STATE_DIM, C = DYN_T::CONTROL_DIM;
    const int T = this->getNumTimesteps();
    if (!nominalStateInit_)
    {
      nominal_state_trajectory_.col(0) = state;
      nominalStateInit_ = true;
    }
    this->free_energy_statistics_.real_sys.previousBaseline = this->getBaselineCost(0);
    this->free_energy_statistics_.nominal_sys.previousBaseline = this->getBaselineCost(1);
    std::vector<float> x0s(2 * S), uin((size_t)2 * T * C), uout((size_t)2 * T * C);
    for (int opt_iter = 0; opt_iter < this->getNumIters(); opt_iter++)
    {
      for (int i = 0; i < S; i++)
      {
        x0s[i] = state(i);
        x0s[S + i] = nominal_state_trajectory_(i, 0);
      }
      memcpy(uin.data(), this->control_.data(), sizeof(float) * T * C);
      memcpy(uin.data() + (size_t)T * C, nominal_control_trajectory_.data(), sizeof(float) * T * C);
      this->solve(x0s.data(), uin.data(), optimization_stride, opt_iter, uout.data());
      memcpy(this->control_.data(), uout.data(), sizeof(float) * T * C);
      memcpy(nominal_control_trajectory_.data(), uout.data() + (size_t)T * C, sizeof(float) * T * C);
      computeStateTrajectory(state);
      if (this->getBaselineCost(0) < this->getBaselineCost(1) + getNominalThreshold())
      {  // tube_mppi_controller.cu:268-280
        this->free_energy_statistics_.nominal_state_used = 0;
        nominal_state_trajectory_ = this->state_;
        nominal_control_trajectory_ = this->control_;
      }
      else
      {
        this->free_energy_statistics_.nominal_state_used = 1;
      }
    }
    smoothControlTrajectory();
    computeStateTrajectory(state);
    auto& fe = this->free_energy_statistics_;
    fe.real_sys.normalizerPercent = this->getNormalizerCost(0) / NUM_ROLLOUTS;
    fe.real_sys.increase = this->getBaselineCost(0) - fe.real_sys.previousBaseline;
    fe.nominal_sys.normalizerPercent = this->getNormalizerCost(1) / NUM_ROLLOUTS;
    fe.nominal_sys.increase = this->getBaselineCost(1) - fe.nominal_sys.previousBaseline;
  }
  void updateNominalState(const Eigen::Ref<const control_array>& u)
  {  // tube_mppi_controller.cu:342-349
    state_array x = nominal_state_trajectory_.col(0), xn, xdot;
    output_array out;
    this->model_->step(x, xn, xdot, u, out, 0, this->getDt());
    nominal_state_trajectory_.col(0) = xn;
  }
  void slideControlSequence(int steps) override
  {  // tube_mppi_controller.cu:315-325
    control_array u0 = nominal_control_trajectory_.col(0);
    updateNominalState(u0);
    this->saveControlHistoryHelper(steps, nominal_control_trajectory_, this->control_history_);
    this->slideControlSequenceHelper(steps, nominal_control_trajectory_);
    this->slideControlSequenceHelper(steps, this->control_);
  }
  void smoothControlTrajectory()
  {
    this->smoothControlTrajectoryHelper(nominal_control_trajectory_, this->control_history_);
  }

private:
  void computeStateTrajectory(const Eigen::Ref<const state_array>& x0_actual)
  {  // tube_mppi_controller.cu:333-340
    state_array x0n = nominal_state_trajectory_.col(0);
    output_trajectory nominal_out = output_trajectory::Zero();
    this->computeOutputTrajectoryHelper(nominal_out, nominal_state_trajectory_, x0n, nominal_control_trajectory_);
    this->computeOutputTrajectoryHelper(this->output_, this->state_, x0_actual, this->control_);
  }
  control_trajectory nominal_control_trajectory_ = control_trajectory::Zero();
  state_trajectory nominal_state_trajectory_ = state_trajectory::Zero();
  bool nominalStateInit_ = false;
};
