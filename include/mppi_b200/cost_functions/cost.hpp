/*
 * mppi_b200/cost_functions/cost.hpp — host side of the Cost plugin contract (include/mppi/cost_functions/cost.cuh:15-35):
 * CostParams<C_DIM> with control_cost_coeff / discount, setParams / getParams, GPUSetup no-ops. Device twins live in
 * libmppi_b200.so (mppi-generic_b200/csrc/plugins/costs.cuh), selected by COST_ID; blob() crosses the C-ABI.
 */
#pragma once
#include <string>

#include "../utils/common.hpp"

template <int C_DIM>
struct CostParams
{  // cost.cuh:15-31
  static const int CONTROL_DIM = C_DIM;
  float control_cost_coeff[C_DIM];
  float discount = 1.0f;
  CostParams()
  {
    for (int i = 0; i < C_DIM; ++i)
      control_cost_coeff[i] = 1.0f;
  }
};

namespace MPPI_internal
{
template <class CLASS_T, class PARAMS_T, class BLOB_T, int COST_ID_V>
class Cost
{
public:
  static const int COST_ID = COST_ID_V;
  typedef PARAMS_T COST_PARAMS_T;
  typedef BLOB_T BLOB;
  Cost() = default;
  virtual ~Cost() = default;
  void setParams(const PARAMS_T& params)
  {
    params_ = params;
  }
  PARAMS_T getParams() const
  {
    return params_;
  }
  void GPUSetup()
  {
  }
  void freeCudaMem()
  {
  }
  void bindToStream(cudaStream_t)
  {
  }
  virtual std::string getCostFunctionName() const
  {
    return "cost function name not set";
  }
  template <class B>
  void fillBase(B& b) const
  {
    for (int i = 0; i < MPPIB_MAX_CONTROL_DIM; i++)
      b.control_cost_coeff[i] = i < PARAMS_T::CONTROL_DIM ? params_.control_cost_coeff[i] : 0.0f;
    b.discount = params_.discount;
  }
  const float* costmap() const
  {
    return nullptr;
  }

protected:
  PARAMS_T params_;
};
}  // namespace MPPI_internal
