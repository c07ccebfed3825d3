/*
 * ColoredMPPIController — host class of include/mppi/controllers/ColoredMPPI/colored_mppi_controller.cuh: VanillaMPPI's
 * flow with the ColoredNoise sampler by default, an optional state leash on the initial condition
 * (colored_mppi_controller.cu:150-153, Dynamics::enforceLeash dynamics.cuh:448-466), Tsallis weights when gamma and r
 * are both non-zero (:199-209 -> mppib_set_tsallis) and the clamp of control 1 after smoothing (:232-238).
 */
#pragma once
#include <cmath>

#include "../controller.hpp"
#include "../../sampling_distributions/colored_noise/colored_noise.hpp"

template <int S_DIM, int C_DIM, int MAX_TIMESTEPS>
struct ColoredMPPIParams : public ControllerParams<S_DIM, C_DIM, MAX_TIMESTEPS>
{  // colored_mppi_controller.cuh:16-22
  float r = 0;
  float gamma = 0;
  Eigen::Matrix<float, S_DIM, 1> state_leash_dist_ = Eigen::Matrix<float, S_DIM, 1>::Zero();
};

template <class DYN_T, class COST_T, class FB_T, int MAX_TIMESTEPS, int NUM_ROLLOUTS,
          class SAMPLING_T = ::mppi::sampling_distributions::ColoredNoiseDistribution<typename DYN_T::DYN_PARAMS_T>,
          class PARAMS_T = ColoredMPPIParams<DYN_T::STATE_DIM, DYN_T::CONTROL_DIM, MAX_TIMESTEPS>>
class ColoredMPPIController
  : public Controller<DYN_T, COST_T, FB_T, SAMPLING_T, MAX_TIMESTEPS, NUM_ROLLOUTS, PARAMS_T, 1>
{
public:
  typedef Controller<DYN_T, COST_T, FB_T, SAMPLING_T, MAX_TIMESTEPS, NUM_ROLLOUTS, PARAMS_T, 1> PARENT_CLASS;
  using control_array = typename PARENT_CLASS::control_array;
  using control_trajectory = typename PARENT_CLASS::control_trajectory;
  using state_trajectory = typename PARENT_CLASS::state_trajectory;
  using state_array = typename PARENT_CLASS::state_array;

  ColoredMPPIController(DYN_T* model, COST_T* cost, FB_T* fb_controller, SAMPLING_T* sampler, float dt, int max_iter,
                        float lambda, float alpha, int num_timesteps = MAX_TIMESTEPS,
                        const Eigen::Ref<const control_trajectory>& init_control_traj = control_trajectory::Zero(),
                        cudaStream_t stream = nullptr)
    : PARENT_CLASS(model, cost, fb_controller, sampler, dt, max_iter, lambda, alpha, num_timesteps, init_control_traj,
                   stream)
  {
    this->chooseAppropriateKernel();
  }
  // colored_mppi_controller.cuh: the PARAMS_T constructor (gamma, r and the leash travel in the parameter struct)
  ColoredMPPIController(DYN_T* model, COST_T* cost, FB_T* fb_controller, SAMPLING_T* sampler, PARAMS_T& params,
                        cudaStream_t stream = nullptr)
    : PARENT_CLASS(model, cost, fb_controller, sampler, params, stream)
  {
    pushWeighting();
    this->chooseAppropriateKernel();
  }
  void setParams(const PARAMS_T& p)
  {
    PARENT_CLASS::setParams(p);
    pushWeighting();
  }
  std::string getControllerName() override
  {
    return "Colored MPPI";
  }
  // colored_mppi_controller.cuh:96-184
  void setGamma(float gamma)
  {
    this->params_.gamma = gamma;
    pushWeighting();
  }
  float getGamma()
  {
    return this->params_.gamma;
  }
  void setRExp(float r)
  {
    this->params_.r = r;
    pushWeighting();
  }
  float getRExp()
  {
    return this->params_.r;
  }
  void setOffsetDecayRate(float decay_rate)
  {
    this->sampler_->setOffsetDecayRate(decay_rate);
    this->pushParams();
  }
  float getOffsetDecayRate()
  {
    return this->sampler_->getOffsetDecayRate();
  }
  void setColoredNoiseExponents(std::vector<float>& new_exponents)
  {
    auto sp = this->sampler_->getParams();
    for (size_t i = 0; i < new_exponents.size(); i++)
      sp.exponents[i] = new_exponents[i];
    this->sampler_->setParams(sp);
    this->pushParams();
  }
  float getColoredNoiseExponent(int index)
  {
    return this->sampler_->getParams().exponents[index];
  }
  void setStateLeashLength(float new_state_leash, int index = 0)
  {
    this->params_.state_leash_dist_[index] = new_state_leash;
  }
  float getStateLeashLength(int index)
  {
    return this->params_.state_leash_dist_[index];
  }
  bool getLeashActive()
  {
    return leash_active_;
  }
  void setLeashActive(bool v)
  {
    leash_active_ = v;
  }

  void computeControl(const Eigen::Ref<const state_array>& state, int optimization_stride = 1) override
  {
    this->free_energy_statistics_.real_sys.previousBaseline = this->getBaselineCost();
    state_array local_state = state;
    if (getLeashActive())
    {  // the model's own enforceLeash (dynamics.cuh:448-466, RacerDubins: racer_dubins.cu:177-230) against the planned state
       // leash_jump_ steps ahead (colored_mppi_controller.cu:150-153)
      state_array nominal = this->state_.col(leash_jump_);
      state_array leash = this->params_.state_leash_dist_;
      this->model_->enforceLeash(state, nominal, leash, local_state);
    }
    control_trajectory u_nominal = this->control_;
    for (int opt_iter = 0; opt_iter < this->getNumIters(); opt_iter++)
    {
      control_trajectory u_out = control_trajectory::Zero();
      u_nominal = this->control_;
      this->solve(local_state.data(), this->control_.data(), optimization_stride, opt_iter, u_out.data());
      this->control_ = u_out;
    }
    if (this->getTotalSampledTrajectories() > 0)  // colored_mppi_controller.cu:242-248
      this->pickSampledControls(local_state, u_nominal, this->control_);
    this->free_energy_statistics_.real_sys.normalizerPercent = this->getNormalizerCost() / NUM_ROLLOUTS;
    this->free_energy_statistics_.real_sys.increase =
        this->getBaselineCost() - this->free_energy_statistics_.real_sys.previousBaseline;
    this->smoothControlTrajectoryHelper(this->control_, this->control_history_);
    this->computeOutputTrajectoryHelper(this->output_, this->state_, local_state, this->control_);
    if (DYN_T::CONTROL_DIM > 1)
      for (int i = 0; i < this->getNumTimesteps(); i++)  // colored_mppi_controller.cu:232-238
        this->control_(1, i) =
            fminf(fmaxf(this->control_(1, i), this->model_->control_rngs_[1].x), this->model_->control_rngs_[1].y);
  }
  void slideControlSequence(int steps) override
  {  // colored_mppi_controller.cu:268-276
    leash_jump_ = steps;
    this->saveControlHistoryHelper(steps, this->control_, this->control_history_);
    this->slideControlSequenceHelper(steps, this->control_);
  }

protected:
  // a re-created engine (new horizon, stream or write-back flag) starts with the exponential weights: give it ours again
  void onEngineCreated() override
  {
    const bool tsallis = this->params_.gamma != 0 && this->params_.r != 0;
    if (tsallis && (this->extra_flags_ & MPPIB_FLAG_WRITEBACK_CONTROLS))
      MPPIB_HANDLE(mppib_set_tsallis(this->engine_, this->params_.gamma, this->params_.r));
  }

private:
  void pushWeighting()
  {
    const bool tsallis = this->params_.gamma != 0 && this->params_.r != 0;
    if (tsallis && !(this->extra_flags_ & MPPIB_FLAG_WRITEBACK_CONTROLS))
    {  // the Tsallis reduction reads the written-back controls: re-create the engine with that buffer
      this->extra_flags_ |= MPPIB_FLAG_WRITEBACK_CONTROLS;
      this->createEngine();
    }
    MPPIB_HANDLE(mppib_set_tsallis(this->engine_, tsallis ? this->params_.gamma : 0.0f, tsallis ? this->params_.r : 0.0f));
  }
  int leash_jump_ = 1;
  bool leash_active_ = false;
};
