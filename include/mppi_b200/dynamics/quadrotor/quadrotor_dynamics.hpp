/* QuadrotorDynamics — host class of include/mppi/dynamics/quadrotor/quadrotor_dynamics.cuh:9-120 (same params struct,
 * constructors, index enums). Device twin: mppi-generic_b200/csrc/plugins/dynamics.cuh (QuadrotorDynamics). */
#pragma once
#include <cmath>

#include "../dynamics.hpp"

struct QuadrotorDynamicsParams
{  // quadrotor_dynamics.cuh:9-63
  enum class StateIndex : int
  {
    POS_X = 0, POS_Y, POS_Z, VEL_X, VEL_Y, VEL_Z, QUAT_W, QUAT_X, QUAT_Y, QUAT_Z, ANG_VEL_X, ANG_VEL_Y, ANG_VEL_Z,
    NUM_STATES
  };
  enum class ControlIndex : int { ANG_RATE_X = 0, ANG_RATE_Y, ANG_RATE_Z, THRUST, NUM_CONTROLS };
  enum class OutputIndex : int
  {
    POS_X = 0, POS_Y, POS_Z, VEL_X, VEL_Y, VEL_Z, QUAT_W, QUAT_X, QUAT_Y, QUAT_Z, ANG_VEL_X, ANG_VEL_Y, ANG_VEL_Z,
    NUM_OUTPUTS
  };
  float tau_roll = 0.25;
  float tau_pitch = 0.25;
  float tau_yaw = 0.25;
  float mass = 1;  // kg
  QuadrotorDynamicsParams(float mass_in) : mass(mass_in){};
  QuadrotorDynamicsParams() = default;
};

class QuadrotorDynamics
  : public MPPI_internal::Dynamics<QuadrotorDynamics, mppib_quadrotor_dyn_params, MPPIB_DYN_QUADROTOR, 13, 4, 13>
{
public:
  typedef MPPI_internal::Dynamics<QuadrotorDynamics, mppib_quadrotor_dyn_params, MPPIB_DYN_QUADROTOR, 13, 4, 13>
      PARENT_CLASS;
  typedef QuadrotorDynamicsParams DYN_PARAMS_T;
  // quadrotor_dynamics.cu:11-19: thrust limited to [0, 36], hover thrust as the zero control
  QuadrotorDynamics(cudaStream_t stream = nullptr) : PARENT_CLASS(stream)
  {
    this->control_rngs_[3].x = 0;
    this->control_rngs_[3].y = 36;
    this->zero_control_[3] = MPPIB_GRAVITY;
  }
  // quadrotor_dynamics.cu:4-9
  QuadrotorDynamics(std::array<float2, 4> control_rngs, cudaStream_t stream = nullptr) : PARENT_CLASS(control_rngs, stream)
  {
    this->zero_control_[3] = MPPIB_GRAVITY;
  }
  void setParams(const QuadrotorDynamicsParams& p)
  {
    params_ = p;
  }
  QuadrotorDynamicsParams getParams() const
  {
    return params_;
  }
  std::string getDynamicsModelName() const override
  {
    return "Quadrotor Model";
  }
  state_array getZeroState() const
  {  // quadrotor_dynamics.cu:200-205
    state_array zero = state_array::Zero();
    zero[6] = 1.0f;
    return zero;
  }
  // quadrotor_dynamics.cu:114-122 (in-place host form): Euler step, quaternion renormalised with w >= 0
  void updateState(Eigen::Ref<state_array> state, Eigen::Ref<state_array> state_der, const float dt)
  {
    PARENT_CLASS::updateState(state, state_der, dt);
    const float n = std::sqrt(state[6] * state[6] + state[7] * state[7] + state[8] * state[8] + state[9] * state[9]);
    const float div = (float)((double)n * std::copysign(1.0, (double)state[6]));
    for (int i = 6; i < 10; i++)
      state[i] /= div;
  }
  void printState(const float* s) const
  {
    printf("Pos     x: %8.4f, y: %8.4f, z: %8.4f\n", s[0], s[1], s[2]);
    printf("Vel     x: %8.4f, y: %8.4f, z: %8.4f\n", s[3], s[4], s[5]);
    printf("Quat    w: %8.4f, x: %8.4f, y: %8.4f, z: %8.4f\n", s[6], s[7], s[8], s[9]);
    printf("Ang Vel x: %8.4f, y: %8.4f, z: %8.4f\n", s[10], s[11], s[12]);
  }
  mppib_quadrotor_dyn_params modelBlob() const
  {
    mppib_quadrotor_dyn_params b{};
    b.tau_roll = params_.tau_roll;
    b.tau_pitch = params_.tau_pitch;
    b.tau_yaw = params_.tau_yaw;
    b.mass = params_.mass;
    return b;
  }

protected:
  QuadrotorDynamicsParams params_;
};
