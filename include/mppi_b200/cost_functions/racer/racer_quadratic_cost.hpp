/* RacerQuadraticCost — quadratic tracking cost on the RACER output vector (ours: the RACER cost classes are not part of
 * the reference tree; include/mppi_b200/params.h: mppib_racer_quadratic_cost_params documents the formula). */
#pragma once
#include "../cost.hpp"

struct RacerQuadraticCostParams : public CostParams<2>
{
  float desired_speed = 5.0f;
  float speed_coeff = 4.0f;
  float desired_yaw = 0.0f;
  float yaw_coeff = 20.0f;
  float desired_y = 0.0f;
  float lateral_coeff = 2.0f;
  float steer_coeff = 1.0f;
};

class RacerQuadraticCost : public MPPI_internal::Cost<RacerQuadraticCost, RacerQuadraticCostParams,
                                                      mppib_racer_quadratic_cost_params, MPPIB_COST_RACER_QUADRATIC>
{
public:
  RacerQuadraticCost(cudaStream_t stream = 0)
  {
  }
  std::string getCostFunctionName() const override
  {
    return "RACER quadratic tracking cost";
  }
  mppib_racer_quadratic_cost_params blob() const
  {
    mppib_racer_quadratic_cost_params b{};
    fillBase(b);
    b.desired_speed = params_.desired_speed;
    b.speed_coeff = params_.speed_coeff;
    b.desired_yaw = params_.desired_yaw;
    b.yaw_coeff = params_.yaw_coeff;
    b.desired_y = params_.desired_y;
    b.lateral_coeff = params_.lateral_coeff;
    b.steer_coeff = params_.steer_coeff;
    return b;
  }
};
