// Source-compatibility forwarder for <mppi/feedback_controllers/DDP/ddp.cuh>.
// The DDP solve itself is OUT OF SCOPE (its product, the feedback-gain trajectory, is an INPUT of the RMPPI rollout —
// RobustMPPIController::setFeedbackGains / mppib_set_rmppi). DDPFeedback<DYN_T, NUM_TIMESTEPS> exists here so that user code
// and the explicit instantiations of include/mppi/instantiations/ can keep NAMING the reference's feedback type
// (feedback_controllers/DDP/ddp.cuh:97-141); it carries a gain trajectory the user fills in and computes nothing.
#pragma once
#include <mppi_b200/eigen_shim.hpp>

#include <vector>

template <class DYN_T, int NUM_TIMESTEPS>
class DDPFeedback
{
public:
  typedef Eigen::Matrix<float, DYN_T::CONTROL_DIM, DYN_T::STATE_DIM> feedback_gain_matrix;
  static const int FB_TIMESTEPS = NUM_TIMESTEPS;
  DDPFeedback(DYN_T* model = nullptr, float dt = 0.01f) : model_(model), dt_(dt), fb_gain_traj_(NUM_TIMESTEPS)
  {
    for (auto& k : fb_gain_traj_)
      k = feedback_gain_matrix::Zero();
  }
  // the gain trajectory (C x S, column-major per step: the layout mppib_set_rmppi takes); filled in by the user
  std::vector<feedback_gain_matrix>& getFeedbackGainTrajectory()
  {
    return fb_gain_traj_;
  }
  float getDt() const
  {
    return dt_;
  }
  DYN_T* model_;

private:
  float dt_;
  std::vector<feedback_gain_matrix> fb_gain_traj_;
};
