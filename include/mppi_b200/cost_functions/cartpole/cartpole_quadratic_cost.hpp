/* CartpoleQuadraticCost — include/mppi/cost_functions/cartpole/cartpole_quadratic_cost.cuh:10-62. */
#pragma once
#include <cmath>

#include "../cost.hpp"

struct CartpoleQuadraticCostParams : public CostParams<1>
{
  float cart_position_coeff = 1000;
  float cart_velocity_coeff = 100;
  float pole_angle_coeff = 2000;
  float pole_angular_velocity_coeff = 100;
  float terminal_cost_coeff = 0;
  float desired_terminal_state[4] = { 0, 0, (float)M_PI, 0 };
  CartpoleQuadraticCostParams()
  {
    this->control_cost_coeff[0] = 10.0;
  }
};

class CartpoleQuadraticCost : public MPPI_internal::Cost<CartpoleQuadraticCost, CartpoleQuadraticCostParams,
                                                         mppib_cartpole_cost_params, MPPIB_COST_CARTPOLE_QUADRATIC>
{
public:
  CartpoleQuadraticCost(cudaStream_t stream = 0)
  {
  }
  std::string getCostFunctionName() const override
  {
    return "Cartpole quadratic cost";
  }
  mppib_cartpole_cost_params blob() const
  {
    mppib_cartpole_cost_params b{};
    fillBase(b);
    b.cart_position_coeff = params_.cart_position_coeff;
    b.cart_velocity_coeff = params_.cart_velocity_coeff;
    b.pole_angle_coeff = params_.pole_angle_coeff;
    b.pole_angular_velocity_coeff = params_.pole_angular_velocity_coeff;
    b.terminal_cost_coeff = params_.terminal_cost_coeff;
    for (int i = 0; i < 4; i++)
      b.desired_terminal_state[i] = params_.desired_terminal_state[i];
    return b;
  }
};
