/*
 * plugins/costs.cuh — device twins of the reference's Cost plugins (include/mppi/cost_functions/cost.cuh:34-35 contract:
 * initializeCosts, computeStateCost, computeControlCost (== 0, cost.cuh:205-208), computeRunningCost, terminalCost).
 * One thread owns one sample, so `crash_status` is a thread-private int that is sticky across the horizon exactly as
 * in the reference's single-kernel path (mppi_common.cu:77-78).
 */
#pragma once
#include "../device_utils.cuh"
#include "../../../include/mppi_b200/params.h"

namespace mppib
{
namespace plugins
{
template <class CLASS_T, class PARAMS_T>
struct Cost
{
  using Params = PARAMS_T;
  struct Aux
  {
  };
  // per-block shared scratch ("theta_c" in the reference, cost.cuh:186-189) as a function of the horizon
  __host__ __device__ static constexpr int sharedFloats(int /*T*/)
  {
    return 0;
  }
  // block-cooperative setup of theta_c (Cost::initializeCosts, cost.cuh:186-189)
  __device__ static __forceinline__ void initializeCosts(const Params&, const Aux&, float* /*theta_c*/, int /*T*/)
  {
  }
  __device__ static __forceinline__ float computeControlCost(const Params&, const float* /*u*/, int /*t*/)
  {
    return 0.0f;  // cost.cuh:205-208
  }
  // cost.cu:40-53
  template <class AUX>
  __device__ static __forceinline__ float computeRunningCost(const Params& p, const AUX& aux, const float* theta_c,
                                                             const float* y, const float* u, int t, int* crash)
  {
    return CLASS_T::computeStateCost(p, aux, theta_c, y, t, crash) + CLASS_T::computeControlCost(p, u, t);
  }
};

// powf(discount, t) is the same number for every sample of a block: it is evaluated once per time step into theta_c
// (identical function, identical arguments => identical bits to the reference's per-sample powf call).
template <class PARAMS_T>
__device__ __forceinline__ void fill_discount_table(const PARAMS_T& p, float* theta_c, int T)
{
  for (int t = threadIdx.x; t < T; t += blockDim.x)
    theta_c[t] = powf(p.discount, t);
}

// cost_functions/cartpole/cartpole_quadratic_cost.cu:20-43
struct CartpoleQuadraticCost : public Cost<CartpoleQuadraticCost, mppib_cartpole_cost_params>
{
  __device__ static __forceinline__ float computeStateCost(const Params& params_, const Aux&, const float*,
                                                           const float* state, int, int*)
  {
    return (state[0] - params_.desired_terminal_state[0]) * (state[0] - params_.desired_terminal_state[0]) *
               params_.cart_position_coeff +
           (state[1] - params_.desired_terminal_state[1]) * (state[1] - params_.desired_terminal_state[1]) *
               params_.cart_velocity_coeff +
           (state[2] - params_.desired_terminal_state[2]) * (state[2] - params_.desired_terminal_state[2]) *
               params_.pole_angle_coeff +
           (state[3] - params_.desired_terminal_state[3]) * (state[3] - params_.desired_terminal_state[3]) *
               params_.pole_angular_velocity_coeff;
  }
  __device__ static __forceinline__ float terminalCost(const Params& params_, const Aux& a, const float* state)
  {
    return computeStateCost(params_, a, nullptr, state, 0, nullptr) * params_.terminal_cost_coeff;
  }
};

// cost_functions/double_integrator/double_integrator_circle_cost.cu:8-32
struct DoubleIntegratorCircleCost : public Cost<DoubleIntegratorCircleCost, mppib_di_circle_cost_params>
{
  __host__ __device__ static constexpr int sharedFloats(int T)
  {
    return T;
  }
  __device__ static __forceinline__ void initializeCosts(const Params& p, const Aux&, float* theta_c, int T)
  {
    fill_discount_table(p, theta_c, T);
  }
  __device__ static __forceinline__ float computeStateCost(const Params& params_, const Aux&, const float* theta_c,
                                                           const float* s, int timestep, int*)
  {
    float radial_position = s[0] * s[0] + s[1] * s[1];
    float current_velocity = sqrtf(s[2] * s[2] + s[3] * s[3]);
    float current_angular_momentum = s[0] * s[3] - s[1] * s[2];
    float cost = 0;
    const float crash = theta_c[timestep] * params_.crash_cost;  // powf(discount, timestep); loaded unconditionally, then selected
    if ((radial_position < params_.inner_path_radius2) || (radial_position > params_.outer_path_radius2))
    {
      cost += crash;
    }
    cost += params_.velocity_cost * fabsf(current_velocity - params_.velocity_desired);
    cost += params_.velocity_cost * fabsf(current_angular_momentum - params_.angular_momentum_desired);
    return cost;
  }
  __device__ static __forceinline__ float terminalCost(const Params&, const Aux&, const float*)
  {
    return 0.0f;
  }
};

// cost_functions/autorally/ar_standard_cost.cu:284-413 (device branches)
struct ARStandardCost : public Cost<ARStandardCost, mppib_ar_standard_cost_params>
{
  static constexpr float MAX_COST_VALUE = 1e16f;
  struct Aux
  {
    cudaTextureObject_t costmap_tex;  // float4 texels, point filter, clamp, normalised coords (ar_standard_cost.cu:160-171)
  };
  __host__ __device__ static constexpr int sharedFloats(int T)
  {
    return T;
  }
  __device__ static __forceinline__ void initializeCosts(const Params& p, const Aux&, float* theta_c, int T)
  {
    fill_discount_table(p, theta_c, T);
  }
  // ar_standard_cost.cu:206-243 (device branch)
  __device__ static __forceinline__ float4 queryTextureTransformed(const Params& p, const Aux& aux, float x, float y)
  {
    float u = p.r_c1[0] * x + p.r_c2[0] * y + p.trs[0];
    float v = p.r_c1[1] * x + p.r_c2[1] * y + p.trs[1];
    // An affine map transform (third row 0 0 1: every track map the reference ships) has w == 1 exactly and u / 1 == u, so
    // the four IEEE divisions of a step (66 SASS instructions) are skipped on a block-uniform test of the parameters;
    // projective transforms keep them.
    if (p.r_c1[2] == 0.0f && p.r_c2[2] == 0.0f && p.trs[2] == 1.0f)
      return tex2D<float4>(aux.costmap_tex, u, v);
    float w = p.r_c1[2] * x + p.r_c2[2] * y + p.trs[2];
    return tex2D<float4>(aux.costmap_tex, u / w, v / w);
  }
  __device__ static __forceinline__ float getSpeedCost(const Params& p, const float* s)
  {
    float cost = 0;
    float error = s[4] - p.desired_speed;
    if (p.l1_cost)
      cost = fabsf(error);
    else
      cost = error * error;
    return (p.speed_coeff * cost);
  }
  __device__ static __forceinline__ float getStabilizingCost(const Params& p, const float* s, int* crash_status)
  {
    float stabilizing_cost = 0;
    // reference compares against the double literal 0.001 (ar_standard_cost.cu:304); float(0.001) is the smallest float
    // above it, so `>=` on floats is the identical predicate
    if (fabsf(s[4]) >= 0.001f)
    {
      float slip = -atanf(s[5] / fabsf(s[4]));
      stabilizing_cost = p.slip_coeff * (slip * slip);  // powf(slip, 2) in the reference: <= 2 ulp apart
      if (fabsf(slip) > p.max_slip_ang)
      {
        stabilizing_cost += p.crash_coeff;
      }
    }
    // fabs(s[3]) > M_PI_2 in double (ar_standard_cost.cu:315); float(pi/2) is the smallest float above pi/2
    if (fabsf(s[3]) >= 1.57079632679489661923f)
    {
      crash_status[0] = 1;
    }
    return stabilizing_cost;
  }
  __device__ static __forceinline__ float getCrashCost(const Params& p, const int* crash)
  {
    return crash[0] > 0 ? p.crash_coeff : 0.0f;
  }
  __device__ static __forceinline__ float getTrackCost(const Params& p, const Aux& aux, const float* s, int* crash)
  {
    float track_cost = 0;
    float sn, cs;
    __sincosf(s[2], &sn, &cs);  // __cosf / __sinf in the reference (ar_standard_cost.cu:342-346)
    float x_front = s[0] + p.front_d * cs;
    float y_front = s[1] + p.front_d * sn;
    float x_back = s[0] + p.back_d * cs;
    float y_back = s[1] + p.back_d * sn;
    float track_cost_front = queryTextureTransformed(p, aux, x_front, y_front).x;
    float track_cost_back = queryTextureTransformed(p, aux, x_back, y_back).x;
    track_cost = (fabsf(track_cost_front) + fabsf(track_cost_back)) / 2.0f;
    if (fabsf(track_cost) < p.track_slop)
      track_cost = 0;
    else
      track_cost = p.track_coeff * track_cost;
    if (track_cost_front >= p.boundary_threshold || track_cost_back >= p.boundary_threshold)
      crash[0] = 1;
    return track_cost;
  }
  __device__ static __forceinline__ float computeStateCost(const Params& p, const Aux& aux, const float* theta_c,
                                                           const float* s, int timestep, int* crash_status)
  {
#ifdef MPPIB_EXP_NO_COST
    return s[4] * s[4];
#endif
#ifdef MPPIB_EXP_NO_TEX
    float track_cost = s[0] + s[1];
#else
    float track_cost = getTrackCost(p, aux, s, crash_status);
#endif
    float speed_cost = getSpeedCost(p, s);
    float stabilizing_cost = getStabilizingCost(p, s, crash_status);
    float crash_cost = theta_c[timestep] * getCrashCost(p, crash_status);  // powf(discount, timestep)
    float cost = speed_cost + crash_cost + track_cost + stabilizing_cost;
    if (cost > MAX_COST_VALUE || isnan(cost))
      cost = MAX_COST_VALUE;
    return cost;
  }
  __device__ static __forceinline__ float terminalCost(const Params&, const Aux&, const float*)
  {
    return 0.0f;
  }
};

// Quadratic tracking cost on the RACER output vector (ours, params.h: mppib_racer_quadratic_cost_params; the RACER cost
// classes are not in the reference tree). Output indices: racer_dubins.cuh:35-76.
struct RacerQuadraticCost : public Cost<RacerQuadraticCost, mppib_racer_quadratic_cost_params>
{
  __host__ __device__ static constexpr int sharedFloats(int T)
  {
    return T;
  }
  __device__ static __forceinline__ void initializeCosts(const Params& p, const Aux&, float* theta_c, int T)
  {
    fill_discount_table(p, theta_c, T);
  }
  __device__ static __forceinline__ float computeStateCost(const Params& p, const Aux&, const float* theta_c,
                                                           const float* y, int t, int*)
  {
    const float dv = y[0] - p.desired_speed;                   // BASELINK_VEL_B_X
    const float dyaw = normalizeAngle(y[5] - p.desired_yaw);   // YAW
    const float dy = y[3] - p.desired_y;                       // BASELINK_POS_I_Y
    const float st = y[8];                                     // STEER_ANGLE
    const float cost =
        p.speed_coeff * dv * dv + p.yaw_coeff * dyaw * dyaw + p.lateral_coeff * dy * dy + p.steer_coeff * st * st;
    return cost * theta_c[t];
  }
  __device__ static __forceinline__ float terminalCost(const Params&, const Aux&, const float*)
  {
    return 0.0f;
  }
};

// cost_functions/quadrotor/quadrotor_quadratic_cost.cu:70-132 (device body) with the float-array quaternion helpers of
// utils/math_utils.h:166-211 (QuatInv / QuatMultiply normalised with rsqrtf) and :263-270 (Quat2EulerNWU).
// The reference's NaN guard `sum * (1 - isnan(sum)) + isnan(sum) * MAX_COST_VALUE` evaluates to NaN for a NaN sum
// (NaN * 0), exactly like the unguarded host body, so it is not restated: a NaN cost stays NaN on both sides.
struct QuadrotorQuadraticCost : public Cost<QuadrotorQuadraticCost, mppib_quadrotor_cost_params>
{
  __device__ static __forceinline__ float computeStateCost(const Params& p, const Aux&, const float*, const float* s, int,
                                                           int*)
  {
    float s_diff[13];
#pragma unroll
    for (int i = 0; i < 13; i++)
    {
      const float d = s[i] - p.s_goal[i];
      s_diff[i] = d * d;  // powf(x, 2)
    }
    // QuatSubtract(s + 6, s_goal + 6): q_goal * inverse(q), normalised
    const float* q = s + 6;
    const float* g = p.s_goal + 6;
    const float inv_norm = rsqrtf(MPPIB_SQ(q[0]) + MPPIB_SQ(q[1]) + MPPIB_SQ(q[2]) + MPPIB_SQ(q[3]));
    const float a0 = q[0] * inv_norm, a1 = -q[1] * inv_norm, a2 = -q[2] * inv_norm, a3 = -q[3] * inv_norm;
    float d[4];
    d[0] = g[0] * a0 - g[1] * a1 - g[2] * a2 - g[3] * a3;
    d[1] = g[1] * a0 + g[0] * a1 - g[3] * a2 + g[2] * a3;
    d[2] = g[2] * a0 + g[3] * a1 + g[0] * a2 - g[1] * a3;
    d[3] = g[3] * a0 - g[2] * a1 + g[1] * a2 + g[0] * a3;
    const float dn = rsqrtf(MPPIB_SQ(d[0]) + MPPIB_SQ(d[1]) + MPPIB_SQ(d[2]) + MPPIB_SQ(d[3]));
#pragma unroll
    for (int i = 0; i < 4; i++)
      d[i] *= dn;
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; i++)
      s_diff[i] *= p.x_coeff;
#pragma unroll
    for (int i = 3; i < 6; i++)
      s_diff[i] *= p.v_coeff;
    if (!p.use_euler)
    {
#pragma unroll
      for (int i = 6; i < 10; i++)
        s_diff[i] = p.q_coeff * d[i - 6];
    }
    else
    {
#pragma unroll
      for (int i = 6; i < 10; i++)
        s_diff[i] = 0.0f;
      const float r_diff = atan2f(2.0f * d[3] * d[2] + 2.0f * d[0] * d[1],
                                  d[0] * d[0] + d[3] * d[3] - d[2] * d[2] - d[1] * d[1]);
      const float temp = -2.0f * d[0] * d[2] + 2.0f * d[1] * d[3];
      const float p_diff = -asinf(fmaxf(fminf(1.0f, temp), -1.0f));
      const float y_diff = atan2f(2.0f * d[2] * d[1] + 2.0f * d[3] * d[0],
                                  d[0] * d[0] + d[1] * d[1] - d[2] * d[2] - d[3] * d[3]);
      sum += p.roll_coeff * MPPIB_SQ(r_diff);
      sum += p.pitch_coeff * MPPIB_SQ(p_diff);
      sum += p.yaw_coeff * MPPIB_SQ(y_diff);
    }
#pragma unroll
    for (int i = 10; i < 13; i++)
      s_diff[i] *= p.w_coeff;
#pragma unroll
    for (int i = 0; i < 13; i++)
      sum += s_diff[i];
    return sum;
  }
  __device__ static __forceinline__ float terminalCost(const Params& p, const Aux& a, const float* s)
  {
    return p.terminal_cost_coeff * computeStateCost(p, a, nullptr, s, 0, nullptr);
  }
};

}  // namespace plugins
}  // namespace mppib
