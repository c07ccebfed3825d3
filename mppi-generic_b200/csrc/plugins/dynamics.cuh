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
  static constexpr int MAX_BLOCK_THREADS = 256;  // __launch_bounds__ of the rollout kernel for this model
  static constexpr bool UNROLL_STEPS = true;     // unroll the 4/C steps that share one 16-byte noise group
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

    // cartpole_dynamics.cu:100-106 with the two reciprocals taken by rcp_nr (denominators >= m_c > 0)
    const float denom = m_c + m_p * MPPIB_SQ(sin_theta);
    state_der[0] = state[1];
    state_der[1] = rcp_nr(denom) * (force + m_p * sin_theta * (l_p * MPPIB_SQ(theta_dot) + gravity_ * cos_theta));
    state_der[2] = theta_dot;
    state_der[3] = rcp_nr(l_p * denom) * (-force * cos_theta - m_p * l_p * MPPIB_SQ(theta_dot) * cos_theta * sin_theta -
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
// One thread = one sample; the 6-32-32-4 forward pass is 1344 FMAs per step, issued as packed FP32x2 FMAs (FFMA2,
// sm_100's full-rate FP32 path): the two halves of every FFMA2 are two adjacent OUTPUT neurons, so the accumulation over
// the inputs k runs in the reference's order (k ascending, bias added last, fnn_helper.cu:463-472) and each neuron's sum
// is bit-identical to a scalar FFMA chain. Weights sit in shared memory TRANSPOSED ([in][out]) so one broadcast LDS.128
// (all lanes read the same address: one wavefront) feeds two FFMA2.
//   theta_s: WT1[6][32] | b1[32] | WT2[32][32] | b2[32] | WT3[32][4] | b3[4]      (1412 floats, like the reference)
__device__ __forceinline__ float tanh_fast(float x)
{
  // tanh(x) = 1 - 2 / (exp(2x) + 1) with ex2.approx / rcp.approx: |abs error| < 2e-7 over the whole range (the
  // reference's tanhf, activation_functions.cuh:15-26, is ~1 ulp; its own FNN test bound is 1e-4, fnn_helper_test.cu:546)
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}

struct AutorallyNNDynamics : public Dynamics<AutorallyNNDynamics, mppib_ar_nn_dyn_params, 7, 2, 8>
{
  static constexpr int DYNAMICS_DIM = 4;  // S_DIM - K_DIM
  static constexpr int L1_W = 0, L1_B = L1_W + 6 * 32, L2_W = L1_B + 32, L2_B = L2_W + 32 * 32, L3_W = L2_B + 32,
                       L3_B = L3_W + 32 * 4;
  static constexpr int SHARED_FLOATS = L3_B + 4;  // 1412
  static constexpr int MAX_BLOCK_THREADS = 256;   // 86 registers/thread: up to 7 warps of samples share one SM's tile
  static constexpr bool UNROLL_STEPS = false;     // one copy of the 1344-FMA step body
  struct Aux
  {
    const float* theta_d;  // reference packed layout, MPPIB_AR_NN_NUM_PARAMS floats (fnn_helper.cu:176-183)
  };

  // FNNHelper::initialize (fnn_helper.cu:385-416): block-cooperative global -> shared copy, transposing W on the way.
  __device__ static __forceinline__ void initializeDynamics(const Params&, const Aux& aux, float* theta_s,
                                                            const float* x, float* y)
  {
    const float* g = aux.theta_d;
    for (int i = threadIdx.x; i < SHARED_FLOATS; i += blockDim.x)
    {
      float v;
      if (i < L1_B)
      {  // WT1[k][j] = W1[j][k]
        const int k = i >> 5, j = i & 31;
        v = g[j * 6 + k];
      }
      else if (i < L2_W)
        v = g[192 + (i - L1_B)];
      else if (i < L2_B)
      {
        const int q = i - L2_W, k = q >> 5, j = q & 31;
        v = g[224 + j * 32 + k];
      }
      else if (i < L3_W)
        v = g[1248 + (i - L2_B)];
      else if (i < L3_B)
      {
        const int q = i - L3_W, k = q >> 2, j = q & 3;
        v = g[1280 + j * 32 + k];
      }
      else
        v = g[1408 + (i - L3_B)];
      theta_s[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 7; i++)
      y[i] = x[i];
  }

  // ar_nn_model.cu:123-128 (cosf / sinf, not the fast intrinsics, in the reference's device code)
  __device__ static __forceinline__ void computeKinematics(const Params&, const float* state, float* state_der)
  {
    float sn, cs;
    sincosf(state[2], &sn, &cs);
    state_der[0] = cs * state[4] - sn * state[5];
    state_der[1] = sn * state[4] + cs * state[5];
    state_der[2] = -state[6];
  }

  // One weight row (all OUT outputs of input k) as OUT/4 broadcast LDS.128. `volatile` keeps the loads in program
  // order, which is how the software pipeline below is expressed: row k+1 is requested BEFORE the FFMA2s of row k, so the
  // ~30-cycle shared-memory latency is covered by arithmetic even when a scheduler holds a single warp (ptxas otherwise
  // funnels every load through one register quad and stalls on each: profiles/r01_autorally_k1_notes.md).
  template <int OUT>
  __device__ static __forceinline__ void load_row(float4 (&w)[OUT / 4], const float* row)
  {
    const uint32_t a = smem_u32(row);
#pragma unroll
    for (int j4 = 0; j4 < OUT / 4; j4++)
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(w[j4].x), "=f"(w[j4].y), "=f"(w[j4].z), "=f"(w[j4].w)
                   : "r"(a + 16u * j4));
  }

  template <int IN, int OUT, bool TANH>
  __device__ static __forceinline__ void layer(const float* __restrict__ WT, const float* __restrict__ b,
                                               const float* in, float* out)
  {
    float2 acc[OUT / 2];
#pragma unroll
    for (int j = 0; j < OUT / 2; j++)
      acc[j] = make_float2(0.0f, 0.0f);
    float4 w[OUT / 4], wn[OUT / 4];
    load_row<OUT>(w, WT);
#pragma unroll
    for (int k = 0; k < IN; k++)
    {
      if (k + 1 < IN)
        load_row<OUT>(wn, WT + (k + 1) * OUT);
      else
        load_row<OUT>(wn, b);  // the bias row rides the same pipeline
      const float2 xk = make_float2(in[k], in[k]);
#pragma unroll
      for (int j4 = 0; j4 < OUT / 4; j4++)
      {
        acc[2 * j4] = __ffma2_rn(make_float2(w[j4].x, w[j4].y), xk, acc[2 * j4]);
        acc[2 * j4 + 1] = __ffma2_rn(make_float2(w[j4].z, w[j4].w), xk, acc[2 * j4 + 1]);
      }
#pragma unroll
      for (int j4 = 0; j4 < OUT / 4; j4++)
        w[j4] = wn[j4];
    }
    // w now holds the bias row
#pragma unroll
    for (int j4 = 0; j4 < OUT / 4; j4++)
    {
      const float t0 = acc[2 * j4].x + w[j4].x, t1 = acc[2 * j4].y + w[j4].y, t2 = acc[2 * j4 + 1].x + w[j4].z,
                  t3 = acc[2 * j4 + 1].y + w[j4].w;
      out[4 * j4 + 0] = TANH ? tanh_fast(t0) : t0;
      out[4 * j4 + 1] = TANH ? tanh_fast(t1) : t1;
      out[4 * j4 + 2] = TANH ? tanh_fast(t2) : t2;
      out[4 * j4 + 3] = TANH ? tanh_fast(t3) : t3;
    }
  }

  __device__ static __forceinline__ void computeDynamics(const Params&, const float* theta_s, const float* state,
                                                         const float* control, float* state_der)
  {
    float a0[6], a1[32], a2[32], a3[4];
#pragma unroll
    for (int i = 0; i < DYNAMICS_DIM; i++)
      a0[i] = state[i + (7 - DYNAMICS_DIM)];
    a0[4] = control[0];
    a0[5] = control[1];
    layer<6, 32, true>(theta_s + L1_W, theta_s + L1_B, a0, a1);
    layer<32, 32, true>(theta_s + L2_W, theta_s + L2_B, a1, a2);
    layer<32, 4, false>(theta_s + L3_W, theta_s + L3_B, a2, a3);
#pragma unroll
    for (int i = 0; i < DYNAMICS_DIM; i++)
      state_der[i + (7 - DYNAMICS_DIM)] = a3[i];
  }
};

}  // namespace plugins
}  // namespace mppib
