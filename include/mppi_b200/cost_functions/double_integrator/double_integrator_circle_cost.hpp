/* DoubleIntegratorCircleCost — include/mppi/cost_functions/double_integrator/double_integrator_circle_cost.cuh:8-40. */
#pragma once
#include "../cost.hpp"

struct DoubleIntegratorCircleCostParams : public CostParams<2>
{
  float velocity_cost = 1;
  float crash_cost = 1000;
  float velocity_desired = 2;
  float inner_path_radius2 = 1.875 * 1.875;
  float outer_path_radius2 = 2.125 * 2.125;
  float angular_momentum_desired = 2 * velocity_desired;
  DoubleIntegratorCircleCostParams()
  {
    control_cost_coeff[0] = 0.01;
    control_cost_coeff[1] = 0.01;
    discount = 1.0;
  }
};

class DoubleIntegratorCircleCost
  : public MPPI_internal::Cost<DoubleIntegratorCircleCost, DoubleIntegratorCircleCostParams, mppib_di_circle_cost_params,
                               MPPIB_COST_DI_CIRCLE>
{
public:
  DoubleIntegratorCircleCost(cudaStream_t stream = nullptr)
  {
  }
  std::string getCostFunctionName() const override
  {
    return "Double integrator circle cost";
  }
  mppib_di_circle_cost_params blob() const
  {
    mppib_di_circle_cost_params b{};
    fillBase(b);
    b.velocity_cost = params_.velocity_cost;
    b.crash_cost = params_.crash_cost;
    b.velocity_desired = params_.velocity_desired;
    b.inner_path_radius2 = params_.inner_path_radius2;
    b.outer_path_radius2 = params_.outer_path_radius2;
    b.angular_momentum_desired = params_.angular_momentum_desired;
    return b;
  }
};
