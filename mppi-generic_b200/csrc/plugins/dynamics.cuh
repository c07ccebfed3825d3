/*
 * plugins/dynamics.cuh — device twins of the reference's Dynamics plugins, re-bodied for one-thread-per-sample,
 * register-resident rollouts on sm_100a.
 *
 * The reference's plugin contract (include/mppi/dynamics/dynamics.cuh:67-76,250-300; device bodies
 * include/mppi/dynamics/dynamics.cu:83-155) is kept by NAME and MEANING — STATE_DIM / CONTROL_DIM / OUTPUT_DIM,
 * initializeDynamics, enforceConstraints, computeKinematics, computeDynamics, computeStateDeriv, updateState,
 * stateToOutput, step — but the bodies are static functions over a POD Params blob (include/mppi_b200/params.h)
 * instead of methods of a device-resident object, and x / xdot / y / u are thread-private register arrays instead of
 * per-sample shared-memory slices: there is no blockDim.y lane cooperation and therefore none of the
 * __syncthreads() the reference needs inside step() (dynamics.cu:137-141).
 *
 * `theta_s` keeps its reference meaning: per-block shared scratch the plugin requests (SHARED_FLOATS ==
 * SHARED_MEM_REQUEST_GRD_BYTES/4, managed.cuh:109-116), filled cooperatively in initializeDynamics.
 */
#pragma once
#include "../device_utils.cuh"
#include "../../../include/mppi_b200/params.h"

namespace mppib
{
namespace plugins
{
// dynamics.cu:97-116 — deadband then clamp. State-independent for every in-tree model.
template <int C>
__device__ __forceinline__ void enforceConstraintsDefault(const mppib_control_limits& lim, float* control)
{
#pragma unroll
  for (int i = 0; i < C; i++)
  {
    if (fabsf(control[i]) < lim.deadband[i])
    {
      control[i] = lim.zero_control[i];
    }
    else
    {
      control[i] += lim.deadband[i] * -signf_ref(control[i]);
    }
    control[i] = fminf(fmaxf(lim.rng_lo[i], control[i]), lim.rng_hi[i]);
  }
}

// CRTP base: the parts of Dynamics<CLASS_T, PARAMS_T> (dynamics.cu) every model shares.
template <class CLASS_T, class PARAMS_T, int S, int C, int O>
struct Dynamics
{
  using Params = PARAMS_T;
  static constexpr int STATE_DIM = S;
  static constexpr int CONTROL_DIM = C;
  static constexpr int OUTPUT_DIM = O;
  static constexpr int SHARED_FLOATS = 0;  // SHARED_MEM_REQUEST_GRD_BYTES / 4
  struct Aux
  {
  };

  // dynamics.cuh:429-435 — y <- x on the first min(S,O) entries. theta_s untouched by default.
  __device__ static __forceinline__ void initializeDynamics(const Params&, const Aux&, float* /*theta_s*/, const float* x,
                                                            float* y)
  {
#pragma unroll
    for (int i = 0; i < O && i < S; i++)
      y[i] = x[i];
  }
  __device__ static __forceinline__ void enforceConstraints(const Params& p, const float* /*x*/, float* u)
  {
    enforceConstraintsDefault<C>(p.lim, u);
  }
  __device__ static __forceinline__ void computeKinematics(const Params&, const float*, float*)
  {
  }
  // dynamics.cu:83-95
  __device__ static __forceinline__ void computeStateDeriv(const Params& p, const float* theta_s, const float* x,
                                                           const float* u, float* xdot)
  {
    CLASS_T::computeKinematics(p, x, xdot);
    CLASS_T::computeDynamics(p, theta_s, x, u, xdot);
  }
  // dynamics.cu:118-129 — explicit Euler
  __device__ static __forceinline__ void updateState(const float* x, float* x_next, const float* xdot, float dt)
  {
#pragma unroll
    for (int i = 0; i < S; i++)
      x_next[i] = x[i] + xdot[i] * dt;
  }
  // dynamics.cu:144-155
  __device__ static __forceinline__ void stateToOutput(const float* x, float* y)
  {
#pragma unroll
    for (int i = 0; i < O && i < S; i++)
      y[i] = x[i];
  }
  // dynamics.cu:131-142
  __device__ static __forceinline__ void step(const Params& p, const float* theta_s, const float* x, float* x_next,
                                              float* xdot, const float* u, float* y, int /*t*/, float dt)
  {
    CLASS_T::computeStateDeriv(p, theta_s, x, u, xdot);
    CLASS_T::updateState(x, x_next, xdot, dt);
    CLASS_T::stateToOutput(x_next, y);
  }
};

// ---- Cartpole: dynamics/cartpole/cartpole_dynamics.cu:89-107 (device body) ----------------------------------------
struct CartpoleDynamics : public Dynamics<CartpoleDynamics, mppib_cartpole_dyn_params, 4, 1, 4>
{
  __device__ static __forceinline__ void computeDynamics(const Params& p, const float*, const float* state,
                                                         const float* control, float* state_der)
  {
    float theta = normalizeAngle(state[2]);
    const float sin_theta = __sinf(theta);
    const float cos_theta = __cosf(theta);
    float theta_dot = state[3];
    float force = control[0];
    float m_c = p.cart_mass;
    float m_p = p.pole_mass;
    float l_p = p.pole_length;
    const float gravity_ = p.gravity;

    state_der[0] = state[1];
    state_der[1] = 1.0f / (m_c + m_p * MPPIB_SQ(sin_theta)) *
                   (force + m_p * sin_theta * (l_p * MPPIB_SQ(theta_dot) + gravity_ * cos_theta));
    state_der[2] = theta_dot;
    state_der[3] = 1.0f / (l_p * (m_c + m_p * MPPIB_SQ(sin_theta))) *
                   (-force * cos_theta - m_p * l_p * MPPIB_SQ(theta_dot) * cos_theta * sin_theta -
                    (m_c + m_p) * gravity_ * sin_theta);
  }
};

// ---- Double integrator: dynamics/double_integrator/di_dynamics.cu:46-53 -------------------------------------------
struct DoubleIntegratorDynamics : public Dynamics<DoubleIntegratorDynamics, mppib_di_dyn_params, 4, 2, 4>
{
  __device__ static __forceinline__ void computeDynamics(const Params&, const float*, const float* state,
                                                         const float* control, float* state_der)
  {
    state_der[0] = state[2];
    state_der[1] = state[3];
    state_der[2] = control[0];
    state_der[3] = control[1];
  }
};

// ---- Autorally NeuralNetModel<7,2,3>: dynamics/autorally/ar_nn_model.cu:123-160 + FNNHelper::forward
//      (utils/nn_helpers/fnn_helper.cu:419-484) -------------------------------------------------------------------
// theta_s layout (ours): each layer's W rows padded to a multiple of 4 inputs so a row is read with broadcast LDS.128
// (all 32 lanes of a warp read the same weight at the same time — one wavefront), then the layer's biases:
//   L1: W[32][8] (cols 6,7 zero) | b[32]   L2: W[32][32] | b[32]   L3: W[4][32] | b[4]
struct AutorallyNNDynamics : public Dynamics<AutorallyNNDynamics, mppib_ar_nn_dyn_params, 7, 2, 8>
{
  static constexpr int DYNAMICS_DIM = 4;  // S_DIM - K_DIM
  static constexpr int L1_W = 0, L1_B = L1_W + 32 * 8, L2_W = L1_B + 32, L2_B = L2_W + 32 * 32, L3_W = L2_B + 32,
                       L3_B = L3_W + 4 * 32;
  static constexpr int SHARED_FLOATS = L3_B + 4;  // 1476
  struct Aux
  {
    const float* theta_d;  // reference packed layout, MPPIB_AR_NN_NUM_PARAMS floats (fnn_helper.cu:176-183)
  };

  // FNNHelper::initialize (fnn_helper.cu:385-416): block-cooperative global -> shared copy of the weights.
  __device__ static __forceinline__ void initializeDynamics(const Params&, const Aux& aux, float* theta_s,
                                                            const float* x, float* y)
  {
    const float* g = aux.theta_d;
    for (int i = threadIdx.x; i < SHARED_FLOATS; i += blockDim.x)
    {
      float v;
      if (i < L1_B)
      {  // W1[j][k], k padded 6 -> 8
        const int j = i >> 3, k = i & 7;
        v = k < 6 ? g[j * 6 + k] : 0.0f;
      }
      else if (i < L2_W)
        v = g[192 + (i - L1_B)];
      else if (i < L2_B)
        v = g[224 + (i - L2_W)];
      else if (i < L3_W)
        v = g[224 + 1024 + (i - L2_B)];
      else if (i < L3_B)
        v = g[224 + 1024 + 32 + (i - L3_W)];
      else
        v = g[224 + 1024 + 32 + 128 + (i - L3_B)];
      theta_s[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 7; i++)
      y[i] = x[i];
  }

  // ar_nn_model.cu:123-128
  __device__ static __forceinline__ void computeKinematics(const Params&, const float* state, float* state_der)
  {
    state_der[0] = cosf(state[2]) * state[4] - sinf(state[2]) * state[5];
    state_der[1] = sinf(state[2]) * state[4] + cosf(state[2]) * state[5];
    state_der[2] = -state[6];
  }

  template <int IN4 /*inputs/4*/, int OUT, bool TANH>
  __device__ static __forceinline__ void layer(const float* __restrict__ W, const float* __restrict__ b,
                                               const float* in, float* out)
  {
#pragma unroll
    for (int j = 0; j < OUT; j++)
    {
      float tmp = 0.0f;
#pragma unroll
      for (int k4 = 0; k4 < IN4; k4++)
      {
        const float4 w = *reinterpret_cast<const float4*>(W + j * IN4 * 4 + k4 * 4);
        // same k order as the reference's inner loop (fnn_helper.cu:466-470)
        tmp += w.x * in[k4 * 4 + 0];
        tmp += w.y * in[k4 * 4 + 1];
        tmp += w.z * in[k4 * 4 + 2];
        tmp += w.w * in[k4 * 4 + 3];
      }
      tmp += b[j];
      out[j] = TANH ? tanhf(tmp) : tmp;
    }
  }

  __device__ static __forceinline__ void computeDynamics(const Params&, const float* theta_s, const float* state,
                                                         const float* control, float* state_der)
  {
    float a0[8], a1[32], a2[32], a3[4];
#pragma unroll
    for (int i = 0; i < DYNAMICS_DIM; i++)
      a0[i] = state[i + (7 - DYNAMICS_DIM)];
    a0[4] = control[0];
    a0[5] = control[1];
    a0[6] = 0.0f;
    a0[7] = 0.0f;
    layer<2, 32, true>(theta_s + L1_W, theta_s + L1_B, a0, a1);
    layer<8, 32, true>(theta_s + L2_W, theta_s + L2_B, a1, a2);
    layer<8, 4, false>(theta_s + L3_W, theta_s + L3_B, a2, a3);
#pragma unroll
    for (int i = 0; i < DYNAMICS_DIM; i++)
      state_der[i + (7 - DYNAMICS_DIM)] = a3[i];
  }
};

}  // namespace plugins
}  // namespace mppib
