/* QuadrotorQuadraticCost — include/mppi/cost_functions/quadrotor/quadrotor_quadratic_cost.cuh:9-115. Device twin:
 * mppi-generic_b200/csrc/plugins/costs.cuh (QuadrotorQuadraticCost). */
#pragma once
#include "../cost.hpp"

struct QuadrotorQuadraticCostParams : public CostParams<4>
{
  float s_goal[13] = { 0, 0, 0,     // x
                       0, 0, 0,     // v
                       1, 0, 0, 0,  // q
                       0, 0, 0 };   // w
  float* x_goal()
  {
    return &s_goal[0];
  }
  float* v_goal()
  {
    return &s_goal[3];
  }
  float* q_goal()
  {
    return &s_goal[6];
  }
  float* w_goal()
  {
    return &s_goal[10];
  }
  float x_coeff = 1.0;
  float v_coeff = 1.0;
  bool use_euler = true;
  float q_coeff = 1.0;
  float roll_coeff = 1.0;
  float pitch_coeff = 1.0;
  float yaw_coeff = 1.0;
  float w_coeff = 1.0;
  float terminal_cost_coeff = 0;
  QuadrotorQuadraticCostParams()
  {
    for (int i = 0; i < 4; i++)
      control_cost_coeff[i] = 2.0;
  }
  Eigen::Matrix<float, 13, 1> getDesiredState()
  {
    Eigen::Matrix<float, 13, 1> s;
    for (int i = 0; i < 13; i++)
      s[i] = s_goal[i];
    return s;
  }
};

class QuadrotorQuadraticCost
  : public MPPI_internal::Cost<QuadrotorQuadraticCost, QuadrotorQuadraticCostParams, mppib_quadrotor_cost_params,
                               MPPIB_COST_QUADROTOR_QUADRATIC>
{
public:
  static constexpr float MAX_COST_VALUE = 1e16;
  QuadrotorQuadraticCost(cudaStream_t stream = nullptr)
  {
  }
  std::string getCostFunctionName() const override
  {
    return "Quadrotor quadratic cost";
  }
  mppib_quadrotor_cost_params blob() const
  {
    mppib_quadrotor_cost_params b{};
    fillBase(b);
    for (int i = 0; i < 13; i++)
      b.s_goal[i] = params_.s_goal[i];
    b.x_coeff = params_.x_coeff;
    b.v_coeff = params_.v_coeff;
    b.use_euler = params_.use_euler ? 1 : 0;
    b.q_coeff = params_.q_coeff;
    b.roll_coeff = params_.roll_coeff;
    b.pitch_coeff = params_.pitch_coeff;
    b.yaw_coeff = params_.yaw_coeff;
    b.w_coeff = params_.w_coeff;
    b.terminal_cost_coeff = params_.terminal_cost_coeff;
    return b;
  }
};
