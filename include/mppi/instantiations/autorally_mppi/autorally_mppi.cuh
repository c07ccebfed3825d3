// Source-compatibility forwarder: the reference's include path, served by the B200 host layer.
// With -DMPPIB_USE_INSTANTIATION_LIBRARY the controller classes the reference pre-builds in src/controllers/autorally/ are
// declared `extern template` here and come from libmppi_b200_controllers.so (src/controllers/build.sh) instead of being
// instantiated in the including translation unit — the role of the reference's instantiation libraries
// (src/controllers/autorally/autorally_mppi.cu:10-11). Without the macro everything stays header-only (g++ compiles a controller in ~2 s).
#pragma once
#include <mppi/feedback_controllers/DDP/ddp.cuh>
#include <mppi_b200/controllers/MPPI/mppi_controller.hpp>
#include <mppi/cost_functions/autorally/ar_standard_cost.cuh>
#include <mppi/dynamics/autorally/ar_nn_model.cuh>

// instantiations/autorally_mppi/autorally_mppi.cuh:9-19 (BLOCKSIZE_X / BLOCKSIZE_Y have no meaning here: the engine picks its
// own launch geometry)
const int MPPI_NUM_ROLLOUTS__ = 1920;
const int NUM_TIMESTEPS = 150;
typedef NeuralNetModel<7, 2, 3> DynamicsModel;
typedef ARStandardCost CostFunctionClass;
typedef DDPFeedback<DynamicsModel, NUM_TIMESTEPS> FEEDBACK_T;
typedef mppi::sampling_distributions::GaussianDistribution<DynamicsModel::DYN_PARAMS_T> Sampler;

#if defined(MPPIB_USE_INSTANTIATION_LIBRARY) && !defined(MPPIB_INSTANTIATIONS_BUILD)
#define MPPIB_INST extern template class
#elif defined(MPPIB_INSTANTIATIONS_BUILD)
#define MPPIB_INST template class
#endif
#ifdef MPPIB_INST
MPPIB_INST VanillaMPPIController<DynamicsModel, CostFunctionClass, FEEDBACK_T, NUM_TIMESTEPS, MPPI_NUM_ROLLOUTS__, Sampler>;
#undef MPPIB_INST
#endif
