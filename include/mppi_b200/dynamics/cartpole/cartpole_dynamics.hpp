/* CartpoleDynamics — host class of include/mppi/dynamics/cartpole/cartpole_dynamics.cuh:44-110 (same constructor,
 * setParams/getParams, members). */
#pragma once
#include "../dynamics.hpp"

struct CartpoleDynamicsParams
{  // cartpole_dynamics.cuh:8-37
  enum class StateIndex : int { POS_X = 0, VEL_X, THETA, THETA_DOT, NUM_STATES };
  enum class ControlIndex : int { FORCE = 0, NUM_CONTROLS };
  enum class OutputIndex : int { POS_X = 0, VEL_X, THETA, THETA_DOT, NUM_OUTPUTS };
  float cart_mass = 1.0f;
  float pole_mass = 1.0f;
  float pole_length = 1.0f;
  CartpoleDynamicsParams() = default;
  CartpoleDynamicsParams(float cart_mass, float pole_mass, float pole_length)
    : cart_mass(cart_mass), pole_mass(pole_mass), pole_length(pole_length){};
};

class CartpoleDynamics
  : public MPPI_internal::Dynamics<CartpoleDynamics, mppib_cartpole_dyn_params, MPPIB_DYN_CARTPOLE, 4, 1, 4>
{
public:
  typedef CartpoleDynamicsParams DYN_PARAMS_T;
  CartpoleDynamics(float cart_mass = 1.0f, float pole_mass = 1.0f, float pole_length = 1.0f, cudaStream_t stream = 0)
  {
    params_ = CartpoleDynamicsParams(cart_mass, pole_mass, pole_length);
  }
  void setParams(const CartpoleDynamicsParams& p)
  {
    params_ = p;
  }
  CartpoleDynamicsParams getParams() const
  {
    return params_;
  }
  float getCartMass() const
  {
    return params_.cart_mass;
  }
  float getPoleMass() const
  {
    return params_.pole_mass;
  }
  float getPoleLength() const
  {
    return params_.pole_length;
  }
  float getGravity() const
  {
    return gravity_;
  }
  std::string getDynamicsModelName() const override
  {
    return "Cartpole";
  }
  void printState(const float* state) const
  {
    printf("Cart position: %f; Cart velocity: %f; Pole angle: %f; Pole rate: %f \n", state[0], state[1], state[2],
           state[3]);
  }
  mppib_cartpole_dyn_params modelBlob() const
  {
    mppib_cartpole_dyn_params b{};
    b.cart_mass = params_.cart_mass;
    b.pole_mass = params_.pole_mass;
    b.pole_length = params_.pole_length;
    b.gravity = gravity_;
    return b;
  }

protected:
  CartpoleDynamicsParams params_;
  const float gravity_ = 9.81;  // cartpole_dynamics.cuh:101
};
