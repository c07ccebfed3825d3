/*
 * mppi_oracle.cpp — CPU restatement of the reference's MPPI rollout-and-reduce path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product (mppi-generic_b200/, include/) may include, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference/). The reference's host
 * twins are written against Eigen, which is not installed in this image, so the identical FP32 arithmetic is restated
 * on plain float arrays, in the same operation order as the cited lines.
 *
 * Parity pinning: tests/test_oracle_golden.py checks this file against the known-answer values held by the
 * reference's own tests: FNN all-ones => 33, ARStandardCost speed/slip/track/crash values on the
 * generated track map, Savitzky-Golay smoothing values, slide semantics, normExp/min/sum identities, weighted
 * reduction against the serial triple loop. The cuRAND / cuFFT numeric streams are third-party (libcurand 10.3.10,
 * libcufft 11.4.1 from CUDA 12.9); no reference test pins them ("parity unpinned" at that boundary) — we call the
 * same library entry points with the same generator type / seed / offset / count instead.
 *
 * Build: make -C oracle   (g++ -O2, no -ffast-math, -ffp-contract=off so products are rounded like the Eigen host
 * code compiled without FMA contraction).
 */
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <curand.h>

#include "../include/mppi_b200/params.h"

#ifndef M_PIf32
#define M_PIf32 3.14159265358979323846f
#endif

namespace orc
{
// ---------------------------------------------------------------------------------------------------------------
// utils/math_utils.h:744-747 (the float overload wins overload resolution for float arguments)
static inline float sign(float v)
{
  return v >= 0 ? 1 : -1;
}

// utils/angle_utils.cuh:20-26
static inline float normalizeAngle(float angle)
{
  const float result = fmodf(angle + M_PIf32, 2.0f * M_PIf32);
  if (result <= 0.0f)
    return result + M_PIf32;
  return result - M_PIf32;
}

#define SQ(a) ((a) * (a))

// dynamics/dynamics.cuh:250-264 (host enforceConstraints)
template <int C>
static inline void enforceConstraints(const mppib_control_limits& lim, float* control)
{
  for (int i = 0; i < C; i++)
  {
    if (fabsf(control[i]) < lim.deadband[i])
    {
      control[i] = lim.zero_control[i];
    }
    else
    {
      control[i] += lim.deadband[i] * -sign(control[i]);
    }
    control[i] = fminf(fmaxf(lim.rng_lo[i], control[i]), lim.rng_hi[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Dynamics host twins. Contract (dynamics/dynamics.cuh:277-300): step = computeStateDeriv -> updateState (Euler)
// -> stateToOutput (identity on the first min(S,O) entries).
struct Aux
{
  const float* nn_theta = nullptr;   // packed W,b per layer (fnn_helper.cu:176-183)
  const float* costmap = nullptr;    // float4 per texel, row-major [h][w]
};

struct Cartpole
{
  static constexpr int S = 4, C = 1, O = 4;
  typedef mppib_cartpole_dyn_params P;
  // dynamics/cartpole/cartpole_dynamics.cu:48-69 (host computeDynamics); kinematics empty
  static void computeStateDeriv(const P& p, const Aux&, const float* state, const float* control, float* state_der)
  {
    const float theta = state[2];
    const float sin_theta = sinf(theta);
    const float cos_theta = cosf(theta);
    float theta_dot = state[3];
    float force = control[0];
    float m_c = p.cart_mass;
    float m_p = p.pole_mass;
    float l_p = p.pole_length;
    const float gravity_ = p.gravity;
    state_der[0] = state[1];
    state_der[1] =
        1.0f / (m_c + m_p * SQ(sin_theta)) * (force + m_p * sin_theta * (l_p * SQ(theta_dot) + gravity_ * cos_theta));
    state_der[2] = theta_dot;
    state_der[3] =
        1.0f / (l_p * (m_c + m_p * SQ(sin_theta))) *
        (-force * cos_theta - m_p * l_p * SQ(theta_dot) * cos_theta * sin_theta - (m_c + m_p) * gravity_ * sin_theta);
  }
};

struct DoubleIntegrator
{
  static constexpr int S = 4, C = 2, O = 4;
  typedef mppib_di_dyn_params P;
  // dynamics/double_integrator/di_dynamics.cu:14-22
  static void computeStateDeriv(const P&, const Aux&, const float* state, const float* control, float* state_der)
  {
    state_der[0] = state[2];
    state_der[1] = state[3];
    state_der[2] = control[0];
    state_der[3] = control[1];
  }
};

// utils/nn_helpers/fnn_helper.cu:354-382 (host forward), layout :176-183, tanh = tanhf (activation_functions.cuh:15-26)
static void fnn_forward(const float* theta, const int* layers, int num_layers, const float* input, float* output)
{
  float acts[2][64];
  int cur = 0;
  for (int i = 0; i < layers[0]; i++)
    acts[0][i] = input[i];
  int stride = 0;
  for (int i = 0; i < num_layers - 1; i++)
  {
    const float* W = theta + stride;
    stride += layers[i + 1] * layers[i];
    const float* b = theta + stride;
    stride += layers[i + 1];
    for (int j = 0; j < layers[i + 1]; j++)
    {
      // Eigen (cur_weights * acts + cur_bias): row-vector dot in index order, then + bias
      float tmp = 0;
      for (int k = 0; k < layers[i]; k++)
      {
        tmp += W[j * layers[i] + k] * acts[cur][k];
      }
      tmp += b[j];
      if (i < num_layers - 2)
      {
        tmp = tanhf(tmp);
      }
      acts[1 - cur][j] = tmp;
    }
    cur = 1 - cur;
  }
  for (int j = 0; j < layers[num_layers - 1]; j++)
    output[j] = acts[cur][j];
}

struct AutorallyNN
{
  static constexpr int S = 7, C = 2, O = 8;
  static constexpr int DYNAMICS_DIM = 4;  // S_DIM - K_DIM (ar_nn_model.cuh)
  typedef mppib_ar_nn_dyn_params P;
  // dynamics/autorally/ar_nn_model.cu:90-119 (host computeKinematics + computeDynamics)
  static void computeStateDeriv(const P&, const Aux& aux, const float* state, const float* control, float* state_der)
  {
    state_der[0] = cosf(state[2]) * state[4] - sinf(state[2]) * state[5];
    state_der[1] = sinf(state[2]) * state[4] + cosf(state[2]) * state[5];
    state_der[2] = -state[6];
    float input[6], output[4];
    for (int i = 0; i < DYNAMICS_DIM; i++)
      input[i] = state[i + (S - DYNAMICS_DIM)];
    for (int i = 0; i < C; i++)
      input[DYNAMICS_DIM + i] = control[i];
    static const int layers[4] = { 6, 32, 32, 4 };
    fnn_forward(aux.nn_theta, layers, 4, input, output);
    for (int i = 0; i < DYNAMICS_DIM; i++)
      state_der[i + (S - DYNAMICS_DIM)] = output[i];
  }
};

// dynamics/dynamics.cuh:277-300
template <class DYN>
static inline void dyn_step(const typename DYN::P& p, const Aux& aux, const float* state, float* next_state,
                            float* state_der, const float* control, float* output, float dt)
{
  for (int i = 0; i < DYN::S; i++)
    state_der[i] = 0.0f;  // Eigen state_array locals are written fully by every model used here
  DYN::computeStateDeriv(p, aux, state, control, state_der);
  for (int i = 0; i < DYN::S; i++)
    next_state[i] = state[i] + state_der[i] * dt;
  for (int i = 0; i < DYN::O && i < DYN::S; i++)
    output[i] = next_state[i];
}

// ---------------------------------------------------------------------------------------------------------------
// Cost host twins. computeRunningCost = computeStateCost + computeControlCost(=0) (cost.cuh:136-139,212-219)
struct CartpoleQuadraticCost
{
  typedef mppib_cartpole_cost_params P;
  // cost_functions/cartpole/cartpole_quadratic_cost.cu:8-18
  static float computeStateCost(const P& params_, const Aux&, const float* s, int, int*)
  {
    return (s[0] - params_.desired_terminal_state[0]) * (s[0] - params_.desired_terminal_state[0]) *
               params_.cart_position_coeff +
           (s[1] - params_.desired_terminal_state[1]) * (s[1] - params_.desired_terminal_state[1]) *
               params_.cart_velocity_coeff +
           (s[2] - params_.desired_terminal_state[2]) * (s[2] - params_.desired_terminal_state[2]) *
               params_.pole_angle_coeff +
           (s[3] - params_.desired_terminal_state[3]) * (s[3] - params_.desired_terminal_state[3]) *
               params_.pole_angular_velocity_coeff;
  }
  // cartpole_quadratic_cost.cu:44-55
  static float terminalCost(const P& params_, const Aux& a, const float* s)
  {
    return computeStateCost(params_, a, s, 0, nullptr) * params_.terminal_cost_coeff;
  }
};

struct DICircleCost
{
  typedef mppib_di_circle_cost_params P;
  // cost_functions/double_integrator/double_integrator_circle_cost.cu:34-59
  static float computeStateCost(const P& params_, const Aux&, const float* s, int timestep, int*)
  {
    float radial_position = s[0] * s[0] + s[1] * s[1];
    float current_velocity = sqrtf(s[2] * s[2] + s[3] * s[3]);
    float current_angular_momentum = s[0] * s[3] - s[1] * s[2];
    float cost = 0;
    if ((radial_position < params_.inner_path_radius2) || (radial_position > params_.outer_path_radius2))
    {
      cost += powf(params_.discount, timestep) * params_.crash_cost;
    }
    cost += params_.velocity_cost * std::abs(current_velocity - params_.velocity_desired);
    cost += params_.velocity_cost * std::abs(current_angular_momentum - params_.angular_momentum_desired);
    return cost;
  }
  static float terminalCost(const P&, const Aux&, const float*)
  {
    return 0;
  }
};

struct ARStandardCost
{
  typedef mppib_ar_standard_cost_params P;
  static constexpr float MAX_COST_VALUE = 1e16;  // ar_standard_cost.cuh
  // cost_functions/autorally/ar_standard_cost.cu:225-243 (host branch: -0.5, clamp, round => point sampling)
  static float queryTextureTransformedX(const P& p, const Aux& aux, float x, float y)
  {
    float u = p.r_c1[0] * x + p.r_c2[0] * y + p.trs[0];
    float v = p.r_c1[1] * x + p.r_c2[1] * y + p.trs[1];
    float w = p.r_c1[2] * x + p.r_c2[2] * y + p.trs[2];
    float qx = u / w * p.map_width;
    float qy = v / w * p.map_height;
    qx = qx - 0.5f;
    qy = qy - 0.5f;
    qx = fmaxf(0.0f, fminf(p.map_width - 1, qx));
    qy = fmaxf(0.0f, fminf(p.map_height - 1, qy));
    return aux.costmap[4 * ((size_t)std::round(qy) * p.map_width + (size_t)std::round(qx)) + 0];
  }
  // ar_standard_cost.cu:284-297
  static float getSpeedCost(const P& p, const float* s)
  {
    float cost = 0;
    float error = s[4] - p.desired_speed;
    if (p.l1_cost)
      cost = fabs(error);
    else
      cost = error * error;
    return (p.speed_coeff * cost);
  }
  // ar_standard_cost.cu:300-321
  static float getStabilizingCost(const P& p, const float* s, int* crash_status)
  {
    float stabilizing_cost = 0;
    if (fabs(s[4]) > 0.001)
    {
      float slip = -atan(s[5] / fabs(s[4]));
      stabilizing_cost = p.slip_coeff * powf(slip, 2);
      if (fabs(-atan(s[5] / fabs(s[4]))) > p.max_slip_ang)
      {
        stabilizing_cost += p.crash_coeff;
      }
    }
    if (fabs(s[3]) > M_PI_2)
    {
      crash_status[0] = 1;
    }
    return stabilizing_cost;
  }
  // ar_standard_cost.cu:324-334
  static float getCrashCost(const P& p, const int* crash)
  {
    float crash_cost = 0;
    if (crash[0] > 0)
      crash_cost = p.crash_coeff;
    return crash_cost;
  }
  // ar_standard_cost.cu:337-378 (host branch)
  static float getTrackCost(const P& p, const Aux& aux, const float* s, int* crash)
  {
    float track_cost = 0;
    float x_front = s[0] + p.front_d * cosf(s[2]);
    float y_front = s[1] + p.front_d * sinf(s[2]);
    float x_back = s[0] + p.back_d * cosf(s[2]);
    float y_back = s[1] + p.back_d * sinf(s[2]);
    float track_cost_front = queryTextureTransformedX(p, aux, x_front, y_front);
    float track_cost_back = queryTextureTransformedX(p, aux, x_back, y_back);
    track_cost = (fabs(track_cost_front) + fabs(track_cost_back)) / 2.0;
    if (fabs(track_cost) < p.track_slop)
      track_cost = 0;
    else
      track_cost = p.track_coeff * track_cost;
    if (track_cost_front >= p.boundary_threshold || track_cost_back >= p.boundary_threshold)
      crash[0] = 1;
    return track_cost;
  }
  // ar_standard_cost.cu:381-413
  static float computeStateCost(const P& p, const Aux& aux, const float* s, int timestep, int* crash_status)
  {
    float track_cost = getTrackCost(p, aux, s, crash_status);
    float speed_cost = getSpeedCost(p, s);
    float stabilizing_cost = getStabilizingCost(p, s, crash_status);
    float crash_cost = powf(p.discount, timestep) * getCrashCost(p, crash_status);
    float cost = speed_cost + crash_cost + track_cost + stabilizing_cost;
    if (cost > MAX_COST_VALUE || std::isnan(cost))
      cost = MAX_COST_VALUE;
    return cost;
  }
  static float terminalCost(const P&, const Aux&, const float*)
  {
    return 0.0;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Sampler semantics.
// sampling_distributions/gaussian/gaussian.cu:17-277 (setGaussianControls), applied in place on raw eps.
static void setGaussianControls(const float* mean /*[D][T][C]*/, const mppib_gaussian_params& sp, float* samples
                                /*[D][N][T][C]*/,
                                int C, int T, int N, int D, int optimization_stride, float std_dev_decay)
{
  for (int d = 0; d < D; d++)
    for (int n = 0; n < N; n++)
      for (int t = 0; t < T; t++)
        for (int c = 0; c < C; c++)
        {
          float* v = &samples[(((size_t)d * N + n) * T + t) * C + c];
          const float m = mean[((size_t)d * T + t) * C + c];
          const float sd = std_dev_decay * sp.std_dev[d * C + c];  // gaussian.cu:86-90
          if (n == 0 || t < optimization_stride)
            *v = m;  // :101-107
          else if (n >= (1.0f - sp.pure_noise_trajectories_percentage) * N)
            *v = sd * (*v);  // :108-114
          else
            *v = m + sd * (*v);  // :115-121
        }
}

// gaussian.cu:481-569 — the DEVICE formula (the host overload :632-651 disagrees; SURVEY §8c says follow the device)
static inline float likelihoodRatioCost(const mppib_gaussian_params& sp, const float* mean_t /*[C] of distribution d*/,
                                        const float* u, int C, int d, int sample_index, int N, float lambda,
                                        float alpha)
{
  float cost = 0.0f;
  const bool pure = sample_index >= (1.0f - sp.pure_noise_trajectories_percentage) * N;
  // Device code accumulates float4/float2 lanes then adds them; for C<=2 this equals index-order summation.
  for (int i = 0; i < C; i++)
  {
    float mean_i = pure ? 0.0f : mean_t[i];
    float sd = sp.std_dev[d * C + i];
    cost += sp.control_cost_coeff[i] * mean_i * (mean_i - 2.0f * u[i]) / (sd * sd);
  }
  return 0.5f * lambda * (1.0f - alpha) * cost;
}

// ---------------------------------------------------------------------------------------------------------------
// tests/include/kernel_tests/core/rollout_kernel_test.cu:504-541 (launchCPURolloutKernel) — with the write-back of the
// constrained control that the GPU kernel performs (core/mppi_common.cu:110-117) so that the weighted average sees it.
template <class DYN, class COST>
static void rollout_range(const typename DYN::P& dp, const typename COST::P& cp, const mppib_gaussian_params& sp,
                          const Aux& aux, int N, int T, int D, float dt, float lambda, float alpha, const float* x0,
                          const float* means, float* samples, float* costs, int n_begin, int n_end)
{
  constexpr int S = DYN::S, C = DYN::C, O = DYN::O;
  for (int d = 0; d < D; d++)
  {
    for (int n = n_begin; n < n_end; n++)
    {
      float curr_x[S], next_x[S], x_der[S], u[C], y[O];
      for (int i = 0; i < S; i++)
        curr_x[i] = x0[d * S + i];
      for (int i = 0; i < O; i++)
        y[i] = 0.0f;
      int crash_status = 0;
      float running_cost = 0.0f;
      for (int t = 0; t < T; t++)
      {
        float* us = &samples[(((size_t)d * N + n) * T + t) * C];
        for (int i = 0; i < C; i++)
          u[i] = us[i];
        enforceConstraints<C>(dp.lim, u);
        for (int i = 0; i < C; i++)
          us[i] = u[i];
        dyn_step<DYN>(dp, aux, curr_x, next_x, x_der, u, y, dt);
        running_cost += COST::computeStateCost(cp, aux, y, t, &crash_status);
        running_cost += likelihoodRatioCost(sp, &means[((size_t)d * T + t) * C], u, C, d, n, N, lambda, alpha);
        for (int i = 0; i < S; i++)
          curr_x[i] = next_x[i];
      }
      running_cost += COST::terminalCost(cp, aux, y);
      running_cost /= T;
      costs[(size_t)d * N + n] = running_cost;
    }
  }
}

template <class DYN, class COST>
static void rollout(const void* dp, const void* cp, const mppib_gaussian_params& sp, const Aux& aux, int N, int T,
                    int D, float dt, float lambda, float alpha, const float* x0, const float* means, float* samples,
                    float* costs, int nthreads)
{
  const auto& d = *(const typename DYN::P*)dp;
  const auto& c = *(const typename COST::P*)cp;
  if (nthreads <= 1)
  {
    rollout_range<DYN, COST>(d, c, sp, aux, N, T, D, dt, lambda, alpha, x0, means, samples, costs, 0, N);
    return;
  }
  std::vector<std::thread> th;
  for (int i = 0; i < nthreads; i++)
  {
    int b = (int)((long long)N * i / nthreads), e = (int)((long long)N * (i + 1) / nthreads);
    th.emplace_back([=, &d, &c, &sp, &aux]() {
      rollout_range<DYN, COST>(d, c, sp, aux, N, T, D, dt, lambda, alpha, x0, means, samples, costs, b, e);
    });
  }
  for (auto& t : th)
    t.join();
}

typedef void (*rollout_fn)(const void*, const void*, const mppib_gaussian_params&, const Aux&, int, int, int, float,
                           float, float, const float*, const float*, float*, float*, int);

static rollout_fn pick_rollout(int dyn_id, int cost_id)
{
  if (dyn_id == MPPIB_DYN_CARTPOLE && cost_id == MPPIB_COST_CARTPOLE_QUADRATIC)
    return &rollout<Cartpole, CartpoleQuadraticCost>;
  if (dyn_id == MPPIB_DYN_DOUBLE_INTEGRATOR && cost_id == MPPIB_COST_DI_CIRCLE)
    return &rollout<DoubleIntegrator, DICircleCost>;
  if (dyn_id == MPPIB_DYN_AUTORALLY_NN && cost_id == MPPIB_COST_AR_STANDARD)
    return &rollout<AutorallyNN, ARStandardCost>;
  return nullptr;
}

static void dims(int dyn_id, int* S, int* C, int* O)
{
  switch (dyn_id)
  {
    case MPPIB_DYN_CARTPOLE:
      *S = 4, *C = 1, *O = 4;
      break;
    case MPPIB_DYN_DOUBLE_INTEGRATOR:
      *S = 4, *C = 2, *O = 4;
      break;
    case MPPIB_DYN_AUTORALLY_NN:
      *S = 7, *C = 2, *O = 8;
      break;
    default:
      *S = *C = *O = 0;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Weights. core/mppi_common.cu:858-900 (first minimum wins), :958-966 (expf), :1055-1063 (double sum),
// :1065-1081 (free energy)
static float computeBaselineCost(const float* c, int n)
{
  float best_cost = c[0];
  for (int i = 1; i < n; i++)
    if (c[i] < best_cost)
      best_cost = c[i];
  return best_cost;
}
static void normExpTransform(float* c, int n, float lambda_inv, float baseline)
{
  for (int i = 0; i < n; i++)
  {
    float cost_dif = c[i] - baseline;
    c[i] = expf(-lambda_inv * cost_dif);
  }
}
static float computeNormalizer(const float* w, int n)
{
  double normalizer = 0.0;
  for (int i = 0; i < n; ++i)
    normalizer += w[i];
  return normalizer;
}
static void computeFreeEnergy(float& free_energy, float& free_energy_var, float& free_energy_modified, const float* w,
                              int num_rollouts, float baseline, float lambda)
{
  float var = 0;
  float norm = 0;
  for (int i = 0; i < num_rollouts; i++)
  {
    norm += w[i];
    var += SQ(w[i]);
  }
  norm /= num_rollouts;
  free_energy = -lambda * logf(norm) + baseline;
  free_energy_var = lambda * (var / num_rollouts - SQ(norm));
  float weird_term = free_energy_var / (norm * sqrtf(1.0 * num_rollouts));
  free_energy_modified = lambda * (weird_term + 0.5 * SQ(weird_term));
}

// core/mppi_common.cu:710-737,1115-1160 (weightedReductionKernel with its thread-partial order)
static void weightedReduction(const float* w, const float* du /*[N][T][C]*/, float* out /*[T][C]*/, float normalizer,
                              int T, int N, int C, int sum_stride)
{
  const int nthreads = (N - 1) / sum_stride + 1;
  std::vector<float> inter((size_t)nthreads * C);
  for (int t = 0; t < T; t++)
  {
    std::fill(inter.begin(), inter.end(), 0.0f);
    for (int th = 0; th < nthreads; th++)
      for (int i = 0; i < sum_stride; ++i)
      {
        int n = th * sum_stride + i;
        if (n < N)
        {
          float weight = w[n] / normalizer;
          for (int j = 0; j < C; ++j)
            inter[(size_t)th * C + j] += weight * du[((size_t)n * T + t) * C + j];
        }
      }
    for (int j = 0; j < C; j++)
    {
      float u = 0;
      for (int i = 0; i < nthreads; ++i)
        u += inter[(size_t)i * C + j];
      out[t * C + j] = u;
    }
  }
}

// controllers/controller.cuh:557-586 (5-tap Savitzky-Golay with 2-sample history). u is [T][C] (== Eigen C x T
// column-major), history [2][C] (== Eigen C x 2 column-major).
static void smoothControlTrajectory(float* u, const float* history, int T, int C)
{
  const float coef[5] = { -3.0f / 35.0f, 12.0f / 35.0f, 17.0f / 35.0f, 12.0f / 35.0f, -3.0f / 35.0f };
  std::vector<float> buf((size_t)(T + 4) * C);
  for (int c = 0; c < C; c++)
  {
    buf[0 * C + c] = history[0 * C + c];
    buf[1 * C + c] = history[1 * C + c];
    for (int t = 0; t < T; t++)
      buf[(size_t)(t + 2) * C + c] = u[(size_t)t * C + c];
    buf[(size_t)(T + 2) * C + c] = u[(size_t)(T - 1) * C + c];
    buf[(size_t)(T + 3) * C + c] = u[(size_t)(T - 1) * C + c];
  }
  for (int t = 0; t < T; t++)
    for (int c = 0; c < C; c++)
    {
      float acc = 0.0f;  // Eigen row-vector * matrix product: index-order dot
      for (int k = 0; k < 5; k++)
        acc += coef[k] * buf[(size_t)(t + k) * C + c];
      u[(size_t)t * C + c] = acc;
    }
}

// controllers/controller.cuh:588-600
static void slideControlSequence(float* u, int steps, int T, int C, const float* zero_control,
                                 const float* slide_control_scale)
{
  for (int i = 0; i < T; ++i)
  {
    int ind = std::min(i + steps, T - 1);
    for (int c = 0; c < C; c++)
      u[(size_t)i * C + c] = u[(size_t)ind * C + c];
    if (i + steps > T - 1)
      for (int c = 0; c < C; c++)
        u[(size_t)i * C + c] = (u[(size_t)ind * C + c] - zero_control[c]) * slide_control_scale[c] + zero_control[c];
  }
}

// controllers/controller.cuh:643-663 (computeOutputTrajectoryHelper)
template <class DYN>
static void outputTrajectory(const void* dpv, const Aux& aux, const float* x0, const float* u /*[T][C]*/, int T,
                             float dt, float* states /*[T][S]*/, float* outputs /*[T][O]*/)
{
  const auto& dp = *(const typename DYN::P*)dpv;
  constexpr int S = DYN::S, C = DYN::C, O = DYN::O;
  float state[S], next_state[S], xdot[S], output[O], ui[C];
  for (int i = 0; i < S; i++)
    states[i] = x0[i];
  for (int i = 0; i < O; i++)
    output[i] = 0.0f;
  // initializeDynamics (dynamics.cuh host): stateToOutput-free default => output_result.col(0) = output as initialised.
  // The reference default initializeDynamics copies state to output (dynamics.cuh:416-423).
  for (int i = 0; i < O && i < S; i++)
    output[i] = x0[i];
  for (int i = 0; i < O; i++)
    outputs[i] = output[i];
  for (int t = 0; t < T - 1; ++t)
  {
    for (int i = 0; i < S; i++)
      state[i] = states[(size_t)t * S + i];
    for (int i = 0; i < C; i++)
      ui[i] = u[(size_t)t * C + i];
    enforceConstraints<C>(dp.lim, ui);
    dyn_step<DYN>(dp, aux, state, next_state, xdot, ui, output, dt);
    for (int i = 0; i < S; i++)
      states[(size_t)(t + 1) * S + i] = next_state[i];
    for (int i = 0; i < O; i++)
      outputs[(size_t)(t + 1) * O + i] = output[i];
  }
}
}  // namespace orc

// =================================================================================================================
// C interface for ctypes (tests/, bench.py cpu_baseline)
// =================================================================================================================
extern "C" {

int orc_dims(int dyn_id, int* S, int* C, int* O)
{
  orc::dims(dyn_id, S, C, O);
  return *S ? 0 : -1;
}

// Host cuRAND XORWOW stream, same generator type/seed/offset semantics as controllers/controller.cu:192-207 +
// gaussian.cu:380-381. Returns 0 on success.
int orc_curand_normal(unsigned long long seed, unsigned long long offset, size_t n, float* out)
{
  curandGenerator_t g;
  if (curandCreateGeneratorHost(&g, CURAND_RNG_PSEUDO_DEFAULT))
    return -1;
  int rc = 0;
  if (curandSetPseudoRandomGeneratorSeed(g, seed))
    rc = -2;
  if (!rc && curandSetGeneratorOffset(g, offset))
    rc = -3;
  if (!rc && curandGenerateNormal(g, out, n, 0.0f, 1.0f))
    rc = -4;
  curandDestroyGenerator(g);
  return rc;
}

void orc_set_gaussian_controls(const float* means, const mppib_gaussian_params* sp, float* samples, int C, int T,
                               int N, int D, int optimization_stride, int iteration_num)
{
  // gaussian.cu:423  powf(std_dev_decay, iteration_num)
  orc::setGaussianControls(means, *sp, samples, C, T, N, D, optimization_stride,
                           powf(sp->std_dev_decay, iteration_num));
}

// samples: [D][N][T][C] already holding the sampled controls (after orc_set_gaussian_controls); constrained in place.
int orc_rollout(int dyn_id, int cost_id, const void* dyn_params, const void* cost_params,
                const mppib_gaussian_params* sp, const float* nn_theta, const float* costmap, int N, int T, int D,
                float dt, float lambda, float alpha, const float* x0, const float* means, float* samples, float* costs,
                int nthreads)
{
  orc::rollout_fn f = orc::pick_rollout(dyn_id, cost_id);
  if (!f)
    return -1;
  orc::Aux aux;
  aux.nn_theta = nn_theta;
  aux.costmap = costmap;
  f(dyn_params, cost_params, *sp, aux, N, T, D, dt, lambda, alpha, x0, means, samples, costs, nthreads);
  return 0;
}

float orc_baseline(const float* costs, int n)
{
  return orc::computeBaselineCost(costs, n);
}
void orc_norm_exp(float* costs, int n, float lambda_inv, float baseline)
{
  orc::normExpTransform(costs, n, lambda_inv, baseline);
}
float orc_normalizer(const float* w, int n)
{
  return orc::computeNormalizer(w, n);
}
void orc_free_energy(const float* w, int n, float baseline, float lambda, float* out3)
{
  orc::computeFreeEnergy(out3[0], out3[1], out3[2], w, n, baseline, lambda);
}
void orc_weighted_reduction(const float* w, const float* du, float* out, float normalizer, int T, int N, int C,
                            int sum_stride)
{
  orc::weightedReduction(w, du, out, normalizer, T, N, C, sum_stride);
}
void orc_smooth(float* u, const float* history, int T, int C)
{
  orc::smoothControlTrajectory(u, history, T, C);
}
void orc_slide(float* u, int steps, int T, int C, const float* zero_control, const float* scale)
{
  orc::slideControlSequence(u, steps, T, C, zero_control, scale);
}

int orc_enforce_constraints(const mppib_control_limits* lim, float* u, int C)
{
  switch (C)
  {
    case 1:
      orc::enforceConstraints<1>(*lim, u);
      return 0;
    case 2:
      orc::enforceConstraints<2>(*lim, u);
      return 0;
    case 3:
      orc::enforceConstraints<3>(*lim, u);
      return 0;
    case 4:
      orc::enforceConstraints<4>(*lim, u);
      return 0;
  }
  return -1;
}

int orc_dyn_step(int dyn_id, const void* dyn_params, const float* nn_theta, const float* x, const float* u, float dt,
                 float* x_next, float* xdot, float* y)
{
  orc::Aux aux;
  aux.nn_theta = nn_theta;
  switch (dyn_id)
  {
    case MPPIB_DYN_CARTPOLE:
      orc::dyn_step<orc::Cartpole>(*(const mppib_cartpole_dyn_params*)dyn_params, aux, x, x_next, xdot, u, y, dt);
      return 0;
    case MPPIB_DYN_DOUBLE_INTEGRATOR:
      orc::dyn_step<orc::DoubleIntegrator>(*(const mppib_di_dyn_params*)dyn_params, aux, x, x_next, xdot, u, y, dt);
      return 0;
    case MPPIB_DYN_AUTORALLY_NN:
      orc::dyn_step<orc::AutorallyNN>(*(const mppib_ar_nn_dyn_params*)dyn_params, aux, x, x_next, xdot, u, y, dt);
      return 0;
  }
  return -1;
}

int orc_state_cost(int cost_id, const void* cost_params, const float* costmap, const float* y, int t, int* crash,
                   float* cost_out, float* terminal_out)
{
  orc::Aux aux;
  aux.costmap = costmap;
  switch (cost_id)
  {
    case MPPIB_COST_CARTPOLE_QUADRATIC:
      *cost_out = orc::CartpoleQuadraticCost::computeStateCost(*(const mppib_cartpole_cost_params*)cost_params, aux,
                                                               y, t, crash);
      *terminal_out = orc::CartpoleQuadraticCost::terminalCost(*(const mppib_cartpole_cost_params*)cost_params, aux, y);
      return 0;
    case MPPIB_COST_DI_CIRCLE:
      *cost_out =
          orc::DICircleCost::computeStateCost(*(const mppib_di_circle_cost_params*)cost_params, aux, y, t, crash);
      *terminal_out = 0;
      return 0;
    case MPPIB_COST_AR_STANDARD:
      *cost_out =
          orc::ARStandardCost::computeStateCost(*(const mppib_ar_standard_cost_params*)cost_params, aux, y, t, crash);
      *terminal_out = 0;
      return 0;
  }
  return -1;
}

// individual ARStandardCost terms, for the reference's known-answer test
// (tests/cost_functions/autorally_standard_cost_test.cu:897-982)
void orc_ar_cost_terms(const mppib_ar_standard_cost_params* p, const float* costmap, const float* s, int* crash,
                       float* speed, float* stabilizing, float* track, float* crash_cost)
{
  orc::Aux aux;
  aux.costmap = costmap;
  *track = orc::ARStandardCost::getTrackCost(*p, aux, s, crash);
  *speed = orc::ARStandardCost::getSpeedCost(*p, s);
  *stabilizing = orc::ARStandardCost::getStabilizingCost(*p, s, crash);
  *crash_cost = orc::ARStandardCost::getCrashCost(*p, crash);
}

float orc_ar_query_texture(const mppib_ar_standard_cost_params* p, const float* costmap, float x, float y)
{
  orc::Aux aux;
  aux.costmap = costmap;
  return orc::ARStandardCost::queryTextureTransformedX(*p, aux, x, y);
}

void orc_fnn_forward(const float* theta, const int* layers, int num_layers, const float* input, float* output)
{
  orc::fnn_forward(theta, layers, num_layers, input, output);
}

int orc_output_trajectory(int dyn_id, const void* dyn_params, const float* nn_theta, const float* x0, const float* u,
                          int T, float dt, float* states, float* outputs)
{
  orc::Aux aux;
  aux.nn_theta = nn_theta;
  switch (dyn_id)
  {
    case MPPIB_DYN_CARTPOLE:
      orc::outputTrajectory<orc::Cartpole>(dyn_params, aux, x0, u, T, dt, states, outputs);
      return 0;
    case MPPIB_DYN_DOUBLE_INTEGRATOR:
      orc::outputTrajectory<orc::DoubleIntegrator>(dyn_params, aux, x0, u, T, dt, states, outputs);
      return 0;
    case MPPIB_DYN_AUTORALLY_NN:
      orc::outputTrajectory<orc::AutorallyNN>(dyn_params, aux, x0, u, T, dt, states, outputs);
      return 0;
  }
  return -1;
}

/*
 * One optimisation iteration of VanillaMPPIController::computeControl / TubeMPPIController::computeControl up to and
 * including the new mean (controllers/MPPI/mppi_controller.cu:162-218, Tube-MPPI/tube_mppi_controller.cu:176-258):
 *   eps (raw N(0,1), [N][T][C], shared by both distributions gaussian.cu:378-389) -> setGaussianControls -> rollout
 *   -> baseline -> normExp -> normalizer -> free energy -> weighted reduction.
 * Outputs per distribution d: U_out[d][T][C], baseline[d], normalizer[d], free_energy[d][3]; costs_out [D][N] are the
 * raw trajectory costs; samples_out [D][N][T][C] the constrained sampled controls.
 */
int orc_solve(int dyn_id, int cost_id, const void* dyn_params, const void* cost_params,
              const mppib_gaussian_params* sp, const float* nn_theta, const float* costmap, int N, int T, int D,
              float dt, float lambda, float alpha, const float* x0, const float* U_in, const float* eps,
              int optimization_stride, int iteration_num, int sum_stride, int nthreads, float* U_out,
              float* baseline, float* normalizer, float* free_energy, float* costs_out, float* samples_out)
{
  int S, C, O;
  orc::dims(dyn_id, &S, &C, &O);
  if (!S)
    return -1;
  const size_t per = (size_t)N * T * C;
  std::vector<float> local_samples;
  float* samples = samples_out;
  if (!samples)
  {
    local_samples.resize(per * D);
    samples = local_samples.data();
  }
  for (int d = 0; d < D; d++)
    memcpy(samples + per * d, eps, per * sizeof(float));
  orc_set_gaussian_controls(U_in, sp, samples, C, T, N, D, optimization_stride, iteration_num);
  std::vector<float> costs((size_t)N * D);
  int rc = orc_rollout(dyn_id, cost_id, dyn_params, cost_params, sp, nn_theta, costmap, N, T, D, dt, lambda, alpha, x0,
                       U_in, samples, costs.data(), nthreads);
  if (rc)
    return rc;
  if (costs_out)
    memcpy(costs_out, costs.data(), costs.size() * sizeof(float));
  for (int d = 0; d < D; d++)
  {
    float* c = costs.data() + (size_t)d * N;
    baseline[d] = orc::computeBaselineCost(c, N);
    // mppi_controller.cu:201-202: launchNormExpKernel(..., 1.0 / lambda, baseline) — double division narrowed to float
    orc::normExpTransform(c, N, (float)(1.0 / lambda), baseline[d]);
    normalizer[d] = orc::computeNormalizer(c, N);
    if (free_energy)
      orc::computeFreeEnergy(free_energy[3 * d + 0], free_energy[3 * d + 1], free_energy[3 * d + 2], c, N, baseline[d],
                             lambda);
    orc::weightedReduction(c, samples + per * d, U_out + (size_t)d * T * C, normalizer[d], T, N, C, sum_stride);
  }
  return 0;
}

}  // extern "C"
