/* NeuralNetModel<S_DIM, C_DIM, K_DIM> — host class of include/mppi/dynamics/autorally/ar_nn_model.cuh. Only the
 * <7,2,3> instance (6-32-32-4 network) has a device twin in libmppi_b200.so. */
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "../dynamics.hpp"

struct NNDynamicsParams
{  // ar_nn_model.cuh:18-51
  enum class StateIndex : int { POS_X = 0, POS_Y, YAW, ROLL, BODY_VEL_X, BODY_VEL_Y, YAW_RATE, NUM_STATES };
  enum class ControlIndex : int { STEERING = 0, THROTTLE, NUM_CONTROLS };
  enum class OutputIndex : int { POS_X = 0, POS_Y, YAW, ROLL, BODY_VEL_X, BODY_VEL_Y, YAW_RATE, FILLER_1, NUM_OUTPUTS };
};

template <int S_DIM, int C_DIM, int K_DIM>
class NeuralNetModel
  : public MPPI_internal::Dynamics<NeuralNetModel<S_DIM, C_DIM, K_DIM>, mppib_ar_nn_dyn_params, MPPIB_DYN_AUTORALLY_NN,
                                   S_DIM, C_DIM, 8>
{
  static_assert(S_DIM == 7 && C_DIM == 2 && K_DIM == 3, "libmppi_b200 ships NeuralNetModel<7,2,3> only");

public:
  typedef NNDynamicsParams DYN_PARAMS_T;
  using PARENT = MPPI_internal::Dynamics<NeuralNetModel<S_DIM, C_DIM, K_DIM>, mppib_ar_nn_dyn_params,
                                         MPPIB_DYN_AUTORALLY_NN, S_DIM, C_DIM, 8>;
  static const int DYNAMICS_DIM = S_DIM - K_DIM;
  NeuralNetModel(cudaStream_t stream = 0) : theta_(MPPIB_AR_NN_NUM_PARAMS, 0.0f)
  {
  }
  NeuralNetModel(std::array<float2, C_DIM> control_rngs, cudaStream_t stream = 0) : NeuralNetModel()
  {
    this->setControlRanges(control_rngs);
  }
  // ar_nn_model.cu:40-45 -> FNNHelper::updateModel (fnn_helper.cu:218-257)
  void updateModel(const std::vector<int>& description, const std::vector<float>& data)
  {
    const int expect[4] = { 6, 32, 32, 4 };
    if (description.size() != 4)
      throw std::invalid_argument("Invalid model trying to to be set for NN");
    for (int i = 0; i < 4; i++)
      if (description[i] != expect[i])
        throw std::invalid_argument("Invalid model trying to to be set for NN");
    if (data.size() != (size_t)MPPIB_AR_NN_NUM_PARAMS)
      throw std::invalid_argument("NN parameter vector must hold 1412 floats");
    for (float v : data)
      if (!std::isfinite(v))
        throw std::invalid_argument("NN parameters must be finite");
    theta_ = data;
  }
  // ar_nn_model.cu:58-61 -> FNNHelper::loadParams (fnn_helper.cu:44-127): npz arrays "dynamics_W<i>" (out x in, row-major)
  // and "dynamics_b<i>", i = 1.., float64 in the reference's files; the layer sizes are taken from the arrays
  void loadParams(const std::string& model_path)
  {
    std::vector<int> layers;
    std::vector<float> theta;
    for (int i = 1;; i++)
    {
      const std::string wn = "dynamics_W" + std::to_string(i), bn = "dynamics_b" + std::to_string(i);
      size_t nw = 0, nb = 0;
      if (mppib_host_npz_read(model_path.c_str(), bn.c_str(), nullptr, 0, &nb, nullptr, nullptr) != MPPIB_OK)
      {
        if (i == 1)
          throw std::runtime_error(std::string("Could not load neural net model: ") + mppib_last_error());
        break;
      }
      if (mppib_host_npz_read(model_path.c_str(), wn.c_str(), nullptr, 0, &nw, nullptr, nullptr) != MPPIB_OK || nb == 0 ||
          nw % nb != 0)
        throw std::runtime_error(std::string("Could not load neural net model: ") + mppib_last_error());
      if (i == 1)
        layers.push_back((int)(nw / nb));
      if ((size_t)layers.back() * nb != nw)
        throw std::runtime_error("Could not load neural net model: " + wn + " does not match the previous layer");
      layers.push_back((int)nb);
      const size_t at = theta.size();
      theta.resize(at + nw + nb);
      MPPIB_HANDLE(mppib_host_npz_read(model_path.c_str(), wn.c_str(), theta.data() + at, nw, nullptr, nullptr, nullptr));
      MPPIB_HANDLE(mppib_host_npz_read(model_path.c_str(), bn.c_str(), theta.data() + at + nw, nb, nullptr, nullptr, nullptr));
    }
    updateModel(layers, theta);
  }
  const float* nnWeights() const
  {
    return theta_.data();
  }
  const std::vector<float>& getTheta() const
  {
    return theta_;
  }
  std::string getDynamicsModelName() const override
  {
    return "FCN Autorally Model";
  }
  mppib_ar_nn_dyn_params modelBlob() const
  {
    return mppib_ar_nn_dyn_params{};
  }

private:
  std::vector<float> theta_;
};
