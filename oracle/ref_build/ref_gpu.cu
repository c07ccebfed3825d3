// oracle/ref_build/ref_gpu.cu — TEST / BENCH INFRASTRUCTURE. A C-ABI harness around the UNMODIFIED reference
// (ACDSLab/MPPI-Generic, read where it lies under /root/reference): VanillaMPPIController instantiated from the reference's own
// headers — controllers/MPPI/mppi_controller.cu:151-241, core/mppi_common.cu (rolloutKernel, normExpKernel,
// weightedReductionKernel), sampling_distributions/gaussian/gaussian.cu, the Cartpole and Autorally plugins — for the two
// configurations BASELINE.json's metric is quoted on (C2: Cartpole 8192 x 100, C4: Autorally NN + map cost 32768 x 100).
// Built by oracle/ref_build/build.sh into oracle/_ref/libmppi_ref_gpu.so with an Eigen stand-in (shim/Eigen, the image has
// no Eigen) and a no-op feedback controller instead of DDP (whose Eigen decompositions the stand-in does not provide; the
// hot path never calls it). Label in every report: "reference kernels, shimmed host".
// Used by bench.py (`reference_gpu` block: the reference GPU build's computeControl Hz on the same box) and by
// tests/test_gpu_vs_reference.py (GPU == GPU costs on the same cuRAND seed). Nothing under mppi-generic_b200/ links it.
#include <mppi/controllers/MPPI/mppi_controller.cuh>
#include <mppi/cost_functions/autorally/ar_standard_cost.cuh>
#include <mppi/cost_functions/cartpole/cartpole_quadratic_cost.cuh>
#include <mppi/dynamics/autorally/ar_nn_model.cuh>
#include <mppi/dynamics/cartpole/cartpole_dynamics.cuh>

#include <chrono>
#include <cstring>
#include <string>
#include <vector>

// ---- a feedback controller that does nothing (feedback_controllers/feedback.cuh interface) ---------------------------
template <class DYN_T>
class NullGPUFeedback : public GPUFeedbackController<NullGPUFeedback<DYN_T>, DYN_T, GPUState>
{
public:
  NullGPUFeedback(cudaStream_t stream = 0) : GPUFeedbackController<NullGPUFeedback<DYN_T>, DYN_T, GPUState>(stream)
  {
  }
};
struct NullFeedbackParams
{
};
template <class DYN_T, int T>
class NullFeedback : public FeedbackController<NullGPUFeedback<DYN_T>, NullFeedbackParams, T>
{
public:
  typedef FeedbackController<NullGPUFeedback<DYN_T>, NullFeedbackParams, T> PARENT;
  using state_array = typename PARENT::state_array;
  using control_array = typename PARENT::control_array;
  using state_trajectory = typename PARENT::state_trajectory;
  using control_trajectory = typename PARENT::control_trajectory;
  using FB_STATE = typename PARENT::TEMPLATED_FEEDBACK_STATE;
  NullFeedback(DYN_T* = nullptr, float dt = 0.01f, int num_timesteps = T, cudaStream_t stream = 0)
    : PARENT(dt, num_timesteps, stream)
  {
  }
  void initTrackingController() override
  {
  }
  control_array k_(const Eigen::Ref<const state_array>&, const Eigen::Ref<const state_array>&, int, FB_STATE&) override
  {
    return control_array::Zero();
  }
  void computeFeedback(const Eigen::Ref<const state_array>&, const Eigen::Ref<const state_trajectory>&,
                       const Eigen::Ref<const control_trajectory>&) override
  {
  }
};

// ---- the controller with the first half of computeControl exposed (costs before they are turned into weights) ---------
template <class DYN_T, class COST_T, int T, int N>
class RefController : public VanillaMPPIController<DYN_T, COST_T, NullFeedback<DYN_T, T>, T, N>
{
public:
  typedef VanillaMPPIController<DYN_T, COST_T, NullFeedback<DYN_T, T>, T, N> BASE;
  using BASE::BASE;
  using state_array = typename BASE::state_array;
  // mppi_controller.cu:155-186: state + nominal control to the device, generateSamples, rollout kernel, costs to the host
  void rolloutCosts(const Eigen::Ref<const state_array>& state, int optimization_stride, float* costs_out)
  {
    HANDLE_ERROR(cudaMemcpyAsync(this->initial_state_d_, state.data(), DYN_T::STATE_DIM * sizeof(float),
                                 cudaMemcpyHostToDevice, this->stream_));
    this->copyNominalControlToDevice(false);
    this->sampler_->generateSamples(optimization_stride, 0, this->gen_, false);
    if (this->getKernelChoiceAsEnum() == kernelType::USE_SPLIT_KERNELS)
      mppi::kernels::launchSplitRolloutKernel<DYN_T, COST_T, typename BASE::TEMPLATED_SAMPLING>(
          this->model_, this->cost_, this->sampler_, this->getDt(), this->getNumTimesteps(), N, this->getLambda(),
          this->getAlpha(), this->initial_state_d_, this->output_d_, this->trajectory_costs_d_,
          this->params_.dynamics_rollout_dim_, this->params_.cost_rollout_dim_, this->stream_, false);
    else
      mppi::kernels::launchRolloutKernel<DYN_T, COST_T, typename BASE::TEMPLATED_SAMPLING>(
          this->model_, this->cost_, this->sampler_, this->getDt(), this->getNumTimesteps(), N, this->getLambda(),
          this->getAlpha(), this->initial_state_d_, this->trajectory_costs_d_, this->params_.dynamics_rollout_dim_,
          this->stream_, false);
    HANDLE_ERROR(cudaMemcpyAsync(costs_out, this->trajectory_costs_d_, N * sizeof(float), cudaMemcpyDeviceToHost, this->stream_));
    HANDLE_ERROR(cudaStreamSynchronize(this->stream_));
  }
  int kernelChoice()
  {
    return this->getKernelChoiceAsEnum() == kernelType::USE_SPLIT_KERNELS ? 1 : 0;
  }
  void forceKernel(int split)
  {
    this->setKernelChoice(split ? kernelType::USE_SPLIT_KERNELS : kernelType::USE_SINGLE_KERNEL);
  }
};

struct RefHandleBase
{
  virtual ~RefHandleBase()
  {
  }
  virtual int N() const = 0;
  virtual int T() const = 0;
  virtual int S() const = 0;
  virtual int C() const = 0;
  virtual void rolloutCosts(const float* x0, int stride, float* costs) = 0;
  virtual void computeControl(const float* x0, int stride) = 0;
  virtual void getControl(float* U) = 0;  // [T][C]
  virtual void setControl(const float* U) = 0;
  virtual void stats(float* out) = 0;  // baseline, normalizer
  virtual int kernelChoice() = 0;
  virtual void forceKernel(int split) = 0;
  virtual void slide(int steps) = 0;
};

template <class DYN_T, class COST_T, int TT, int NN>
struct RefHandle : public RefHandleBase
{
  typedef RefController<DYN_T, COST_T, TT, NN> CTL;
  typedef mppi::sampling_distributions::GaussianDistribution<typename DYN_T::DYN_PARAMS_T> SAMPLER;
  DYN_T* model = nullptr;
  COST_T* cost = nullptr;
  NullFeedback<DYN_T, TT>* fb = nullptr;
  SAMPLER* sampler = nullptr;
  CTL* ctl = nullptr;
  ~RefHandle() override
  {
    delete ctl;
    delete sampler;
    delete fb;
    delete cost;
    delete model;
  }
  int N() const override
  {
    return NN;
  }
  int T() const override
  {
    return TT;
  }
  int S() const override
  {
    return DYN_T::STATE_DIM;
  }
  int C() const override
  {
    return DYN_T::CONTROL_DIM;
  }
  typename DYN_T::state_array state(const float* x0)
  {
    typename DYN_T::state_array s;
    for (int i = 0; i < DYN_T::STATE_DIM; i++)
      s(i) = x0[i];
    return s;
  }
  void rolloutCosts(const float* x0, int stride, float* costs) override
  {
    ctl->rolloutCosts(state(x0), stride, costs);
  }
  void computeControl(const float* x0, int stride) override
  {
    ctl->computeControl(state(x0), stride);
  }
  void getControl(float* U) override
  {
    auto u = ctl->getControlSeq();  // C x T, column-major == [t][c]
    memcpy(U, u.data(), sizeof(float) * TT * DYN_T::CONTROL_DIM);
  }
  void setControl(const float* U) override
  {
    typename CTL::control_trajectory u;
    memcpy(u.data(), U, sizeof(float) * TT * DYN_T::CONTROL_DIM);
    ctl->updateImportanceSampler(u);
  }
  void stats(float* out) override
  {
    out[0] = ctl->getBaselineCost();
    out[1] = ctl->getNormalizerCost();
  }
  int kernelChoice() override
  {
    return ctl->kernelChoice();
  }
  void forceKernel(int split) override
  {
    ctl->forceKernel(split);
  }
  void slide(int steps) override
  {
    ctl->slideControlSequence(steps);
  }
  template <class SP>
  void finish(const SP& sp, float dt, float lambda, float alpha, unsigned seed, const int* blk)
  {
    sampler = new SAMPLER(sp);
    fb = new NullFeedback<DYN_T, TT>(model, dt);
    // params constructor: the seed is fixed before the generator is created (controller.cuh:59,117,165), like our engine's
    // lockstep mode assumes
    typename CTL::TEMPLATED_PARAMS params;
    params.dt_ = dt;
    params.lambda_ = lambda;
    params.alpha_ = alpha;
    params.num_iters_ = 1;
    params.num_timesteps_ = TT;
    params.seed_ = seed;
    params.dynamics_rollout_dim_ = dim3(blk[0], blk[1], 1);
    params.cost_rollout_dim_ = dim3(blk[2], blk[3], 1);
    ctl = new CTL(model, cost, fb, sampler, params);
  }
};

typedef NeuralNetModel<7, 2, 3> ARModel;
typedef RefHandle<CartpoleDynamics, CartpoleQuadraticCost, 100, 8192> CartpoleC2;
typedef RefHandle<ARModel, ARStandardCost, 100, 32768> AutorallyC4;
// the same pairs at the sizes the GPU == GPU parity tests use (the oracle finishes these in seconds)
typedef RefHandle<CartpoleDynamics, CartpoleQuadraticCost, 100, 2048> CartpoleSmall;
typedef RefHandle<ARModel, ARStandardCost, 100, 4096> AutorallySmall;

static thread_local std::string g_err;

template <class H>
static RefHandleBase* make_cartpole(const float* p, unsigned seed, const int* blk)
{
  // p: cart_mass pole_mass pole_length | u_lo u_hi | q_pos q_vel q_ang q_angvel control_cost terminal | goal[4] |
  //    std_dev sampler_control_cost pure_noise_pct | dt lambda alpha
  H* h = new H();
  h->model = new CartpoleDynamics(p[0], p[1], p[2]);
  h->model->control_rngs_[0].x = p[3];
  h->model->control_rngs_[0].y = p[4];
  h->cost = new CartpoleQuadraticCost();
  CartpoleQuadraticCostParams cp;
  cp.cart_position_coeff = p[5];
  cp.cart_velocity_coeff = p[6];
  cp.pole_angle_coeff = p[7];
  cp.pole_angular_velocity_coeff = p[8];
  cp.control_cost_coeff[0] = p[9];
  cp.terminal_cost_coeff = p[10];
  for (int i = 0; i < 4; i++)
    cp.desired_terminal_state[i] = p[11 + i];
  h->cost->setParams(cp);
  auto sp = typename H::SAMPLER::SAMPLING_PARAMS_T();
  sp.std_dev[0] = p[15];
  sp.control_cost_coeff[0] = p[16];
  sp.pure_noise_trajectories_percentage = p[17];
  h->finish(sp, p[18], p[19], p[20], seed, blk);
  return h;
}

template <class H>
static RefHandleBase* make_autorally(const float* theta, int ntheta, const char* map_path, const float* p, unsigned seed,
                                     const int* blk)
{
  // p: steer_lo steer_hi throttle_lo throttle_hi | std_dev[2] sampler_control_cost[2] pure_noise_pct | dt lambda alpha
  //    (cost parameters: the reference's defaults, ar_standard_cost.cuh:14-41, like workloads.autorally)
  H* h = new H();
  std::array<float2, 2> rng;
  rng[0] = make_float2(p[0], p[1]);
  rng[1] = make_float2(p[2], p[3]);
  h->model = new ARModel(rng);
  h->model->updateModel({ 6, 32, 32, 4 }, std::vector<float>(theta, theta + ntheta));
  h->cost = new ARStandardCost();
  h->cost->GPUSetup();  // costmapToTexture writes the texture handle into the device copy (ar_standard_cost.cu:179-183)
  if (h->cost->loadTrackData(map_path).empty())
  {
    g_err = std::string("loadTrackData failed for ") + map_path;
    delete h;
    return nullptr;
  }
  auto sp = typename H::SAMPLER::SAMPLING_PARAMS_T();
  sp.std_dev[0] = p[4];
  sp.std_dev[1] = p[5];
  sp.control_cost_coeff[0] = p[6];
  sp.control_cost_coeff[1] = p[7];
  sp.pure_noise_trajectories_percentage = p[8];
  h->finish(sp, p[9], p[10], p[11], seed, blk);
  return h;
}

extern "C" {
const char* refgpu_last_error()
{
  return g_err.c_str();
}
// config: 0 = Cartpole 8192 x 100 (C2), 1 = Cartpole 2048 x 100. blk = the reference's rollout block shapes
// {dynamics_rollout_dim_.x, .y, cost_rollout_dim_.x, .y} (examples/cartpole_example.cu: 64 x 4 for both); the controller
// itself then picks its single or its split rollout kernel by timing both (mppi_controller.cu:45-140)
void* refgpu_create_cartpole(int config, const float* p, unsigned seed, const int* blk)
{
  try
  {
    return config == 0 ? make_cartpole<CartpoleC2>(p, seed, blk) : make_cartpole<CartpoleSmall>(p, seed, blk);
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return nullptr;
  }
}
// config: 0 = Autorally 32768 x 100 (C4), 1 = Autorally 4096 x 100
void* refgpu_create_autorally(int config, const float* theta, int ntheta, const char* map_path, const float* p,
                              unsigned seed, const int* blk)
{
  try
  {
    return config == 0 ? make_autorally<AutorallyC4>(theta, ntheta, map_path, p, seed, blk) :
                         make_autorally<AutorallySmall>(theta, ntheta, map_path, p, seed, blk);
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return nullptr;
  }
}
void refgpu_destroy(void* h)
{
  delete static_cast<RefHandleBase*>(h);
}
void refgpu_dims(void* h, int* N, int* T, int* S, int* C)
{
  RefHandleBase* r = static_cast<RefHandleBase*>(h);
  *N = r->N(), *T = r->T(), *S = r->S(), *C = r->C();
}
int refgpu_kernel_choice(void* h)
{
  return static_cast<RefHandleBase*>(h)->kernelChoice();
}
void refgpu_force_kernel(void* h, int split)
{
  static_cast<RefHandleBase*>(h)->forceKernel(split);
}
int refgpu_rollout_costs(void* h, const float* x0, int stride, float* costs)
{
  try
  {
    static_cast<RefHandleBase*>(h)->rolloutCosts(x0, stride, costs);
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return 1;
  }
}
int refgpu_compute_control(void* h, const float* x0, int stride, float* U, float* stats2)
{
  try
  {
    RefHandleBase* r = static_cast<RefHandleBase*>(h);
    r->computeControl(x0, stride);
    if (U)
      r->getControl(U);
    if (stats2)
      r->stats(stats2);
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return 1;
  }
}
void refgpu_set_control(void* h, const float* U)
{
  static_cast<RefHandleBase*>(h)->setControl(U);
}
void refgpu_slide(void* h, int steps)
{
  static_cast<RefHandleBase*>(h)->slide(steps);
}
// `iters` back-to-back computeControl calls timed with steady_clock around the host call — what the reference's own
// timing test does (tests/controllers/vanilla_mppi_test.cu:290-292). Returns seconds per call.
double refgpu_time_compute_control(void* h, const float* x0, int stride, int warmup, int iters)
{
  RefHandleBase* r = static_cast<RefHandleBase*>(h);
  for (int i = 0; i < warmup; i++)
    r->computeControl(x0, stride);
  cudaDeviceSynchronize();
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < iters; i++)
    r->computeControl(x0, stride);
  cudaDeviceSynchronize();
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count() / iters;
}
}
