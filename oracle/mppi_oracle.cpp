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
#include <condition_variable>
#include <functional>
#include <mutex>
#include <pthread.h>
#include <sched.h>
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <type_traits>
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
  const float* lstm_theta = nullptr; // LSTM weights then head weights (params.h: MPPIB_BLOB_LSTM_WEIGHTS)
  int lstm_hidden = 0;               // H
  int lstm_head = 0;                 // L1 (head layers {H+4, L1, 1})
  const mppib_elevation_map_header* elev = nullptr;  // header + [height][width] floats (params.h), or none = flat ground
};

// TwoDTextureHelper<float> on the host for the TextureParams defaults (texture_helper.cuh:31-55: clamp, linear, normalised):
// worldPoseToMapPose (texture_helper.cu:94-103), mapPoseToTexCoord (:106-123), queryTextureCPU (two_d_texture_helper.cu:151-243)
static float elevationAtWorldPose(const mppib_elevation_map_header* t, float wx, float wy, float wz)
{
  const float* values = reinterpret_cast<const float*>(t + 1);
  // world -> map
  const float diff[3] = { wx - t->origin[0], wy - t->origin[1], wz - t->origin[2] };
  float map[3];
  for (int r = 0; r < 3; r++)
    map[r] = t->rotations[3 * r] * diff[0] + t->rotations[3 * r + 1] * diff[1] + t->rotations[3 * r + 2] * diff[2];
  // map -> pixels -> normalised
  float tx = map[0] / t->resolution[0];
  float ty = map[1] / t->resolution[1];
  tx /= t->width;
  ty /= t->height;
  // queryTextureCPU
  float qx = tx * t->width, qy = ty * t->height;
  qx = qx - 0.5f;
  qy = qy - 0.5f;
  if (qx > t->width - 1)
    qx = t->width - 1;
  else if (qx <= 0.0)
    qx = 0.0;
  if (qy > t->height - 1)
    qy = t->height - 1;
  else if (qy <= 0.0)
    qy = 0.0;
  if (std::isnan(qx) || std::isnan(qy))
    return NAN;
  const int w = t->width;
  const int x_min = std::min((int)std::floor(qx), w - 2), x_max = x_min + 1;
  const int y_min = std::min((int)std::floor(qy), t->height - 2), y_max = y_min + 1;
  const float Q_11 = values[y_min * w + x_min], Q_12 = values[y_min * w + x_max];
  const float Q_21 = values[y_max * w + x_min], Q_22 = values[y_max * w + x_max];
  const float y_min_interp = Q_11 * ((x_max - qx) / (x_max - x_min)) + Q_12 * ((qx - x_min) / (x_max - x_min));
  const float y_max_interp = Q_21 * ((x_max - qx) / (x_max - x_min)) + Q_22 * ((qx - x_min) / (x_max - x_min));
  return y_min_interp * ((y_max - qy) / (y_max - y_min)) + y_max_interp * ((qy - y_min) / (y_max - y_min));
}
// RACER::computeStaticSettling, racer_dubins.cu:359-434, with math::bodyOffsetToWorldPoseEuler (math_utils.h:585-597) and the
// host branch of Euler2DCM_NWU (:457-482)
static float staticSettling(const mppib_elevation_map_header* t, float yaw, float x, float y, float& roll, float& pitch)
{
  float height = 0.0f;
  if (t && t->use)
  {
    float sin_phi, cos_phi, sin_theta, cos_theta, sin_psi, cos_psi;
    sincosf(roll, &sin_phi, &cos_phi);
    sincosf(pitch, &sin_theta, &cos_theta);
    sincosf(yaw, &sin_psi, &cos_psi);
    float M[3][3];
    M[0][0] = cos_theta * cos_psi;
    M[0][1] = sin_phi * sin_theta * cos_psi - cos_phi * sin_psi;
    M[0][2] = cos_phi * sin_theta * cos_psi + sin_phi * sin_psi;
    M[1][0] = cos_theta * sin_psi;
    M[1][1] = sin_phi * sin_theta * sin_psi + cos_phi * cos_psi;
    M[1][2] = cos_phi * sin_theta * sin_psi - sin_phi * cos_psi;
    M[2][0] = -sin_theta;
    M[2][1] = sin_phi * cos_theta;
    M[2][2] = cos_phi * cos_theta;
    const float offsets[4][3] = { { 2.981f, 0.737f, 0.0f }, { 2.981f, -0.737f, 0.0f }, { 0.0f, 0.737f, 0.0f },
                                  { 0.0f, -0.737f, 0.0f } };  // front left, front right, rear left, rear right
    float h[4];
    for (int k = 0; k < 4; k++)
    {
      float w[3];
      for (int r = 0; r < 3; r++)
        w[r] = M[r][0] * offsets[k][0] + M[r][1] * offsets[k][1] + M[r][2] * offsets[k][2];
      h[k] = elevationAtWorldPose(t, w[0] + x, w[1] + y, w[2] + 0.0f);
    }
    float front_diff = h[0] - h[1];
    front_diff = fmaxf(fminf(front_diff, 0.736f * 2.0f), -0.736f * 2.0f);
    float rear_diff = h[2] - h[3];
    rear_diff = fmaxf(fminf(rear_diff, 0.736f * 2.0f), -0.736f * 2.0f);
    const float front_roll = asinf(front_diff / (0.737f * 2.0f));
    const float rear_roll = asinf(rear_diff / (0.737f * 2.0f));
    roll = (front_roll + rear_roll) / 2.0f;
    float left_diff = h[2] - h[0];
    left_diff = fmaxf(fminf(left_diff, 2.98f), -2.98f);
    float right_diff = h[3] - h[1];
    right_diff = fmaxf(fminf(right_diff, 2.98f), -2.98f);
    const float left_pitch = asinf(left_diff / 2.981f);
    const float right_pitch = asinf(right_diff / 2.981f);
    pitch = (left_pitch + right_pitch) / 2.0f;
    height = (h[2] + h[3]) / 2.0f;
    if (!std::isfinite(roll) || fabsf(roll) > (float)M_PI)
      roll = 2.0f * (float)M_PI;
    if (!std::isfinite(pitch) || fabsf(pitch) > (float)M_PI)
      pitch = 2.0f * (float)M_PI;
    if (!std::isfinite(height))
      height = 0.0f;
  }
  else
  {
    roll = 0.0f;
    pitch = 0.0f;
    height = 0.0f;
  }
  return height;
}

// models without recurrent state carry nothing from step to step
struct NoCarry
{
};

struct Cartpole
{
  typedef NoCarry Carry;
  static void initCarry(const Aux&, Carry&)
  {
  }
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
  typedef NoCarry Carry;
  static void initCarry(const Aux&, Carry&)
  {
  }
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
  float acts[2][512];  // widest layer in tree: the init network's 100 (lstm_lstm_helper_test.cu:29)
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
  typedef NoCarry Carry;
  static void initCarry(const Aux&, Carry&)
  {
  }
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

// utils/nn_helpers/lstm_helper.cu:267-323 (host forward: Eigen W_*m * h + W_*i * x + b, sigmoid = 1/(1+expf(-x)),
// activation_functions.cuh:49-59 host branch) followed by the FNN head on [h_next; x] (:325-339).
// Weight layout: lstm_helper.cu:72-88.
static void lstm_forward(const float* w, int I, int H, const float* head_theta, const int* head_layers,
                         int head_num_layers, const float* input, float* h, float* c, float* output)
{
  const int HH = H * H, IH = H * I;
  const float *W_im = w, *W_fm = w + HH, *W_om = w + 2 * HH, *W_cm = w + 3 * HH;
  const float *W_ii = w + 4 * HH, *W_fi = W_ii + IH, *W_oi = W_ii + 2 * IH, *W_ci = W_ii + 3 * IH;
  const float *b_i = w + 4 * HH + 4 * IH, *b_f = b_i + H, *b_o = b_i + 2 * H, *b_c = b_i + 3 * H;
  float c_next[128], h_next[128];
  auto gate = [&](const float* Wm, const float* Wi, const float* b, int i) {
    float hm = 0.0f, im = 0.0f;
    for (int j = 0; j < H; j++)
      hm += Wm[i * H + j] * h[j];
    for (int j = 0; j < I; j++)
      im += Wi[i * I + j] * input[j];
    return (hm + im) + b[i];
  };
  for (int i = 0; i < H; i++)
  {
    const float g_i = 1.0f / (1.0f + expf(-gate(W_im, W_ii, b_i, i)));
    const float g_f = 1.0f / (1.0f + expf(-gate(W_fm, W_fi, b_f, i)));
    const float g_o = 1.0f / (1.0f + expf(-gate(W_om, W_oi, b_o, i)));
    const float g_c = tanhf(gate(W_cm, W_ci, b_c, i));
    c_next[i] = g_i * g_c + g_f * c[i];
    h_next[i] = g_o * tanhf(c_next[i]);
  }
  for (int i = 0; i < H; i++)
  {
    h[i] = h_next[i];
    c[i] = c_next[i];
  }
  float nn_input[192];
  for (int i = 0; i < H; i++)
    nn_input[i] = h[i];
  for (int i = 0; i < I; i++)
    nn_input[H + i] = input[i];
  fnn_forward(head_theta, head_layers, head_num_layers, nn_input, output);
}

// RacerDubinsElevationLSTMSteering, host path: dynamics/racer_dubins/racer_dubins_elevation_lstm_steering.cu:66-118
// (computeLSTMSteering + step), racer_dubins.cu:306-319 (computeParametricDelayDeriv), racer_dubins_elevation.cu:32-67
// (computeParametricAccelDeriv, host), :336-515 (uncertainty Jacobian and Q), :662-741 (propagation),
// lstm_steering.cu:267-285 (updateState, host), racer_dubins.cu:359-434 (static settling, no texture => flat),
// racer_dubins_elevation.cu:69-227 (setOutputs).
struct RacerLSTM
{
  static constexpr int S = 19, C = 2, O = 28;
  static constexpr bool CUSTOM_STEP = true;
  typedef mppib_racer_lstm_dyn_params P;
  struct Carry
  {
    float h[128], c[128];
  };
  enum
  {
    VEL_X = 0, YAW, POS_X, POS_Y, STEER_ANGLE, BRAKE_STATE, ROLL, PITCH, STEER_ANGLE_RATE, UNC_POS_X, UNC_POS_Y, UNC_YAW,
    UNC_VEL_X, UNC_POS_X_Y, UNC_POS_X_YAW, UNC_POS_X_VEL_X, UNC_POS_Y_YAW, UNC_POS_Y_VEL_X, UNC_YAW_VEL_X
  };
  enum
  {
    O_VEL_B_X = 0, O_VEL_B_Y, O_POS_I_X, O_POS_I_Y, O_POS_I_Z, O_YAW, O_ROLL, O_PITCH, O_STEER_ANGLE, O_STEER_ANGLE_RATE,
    O_WF_UP, O_WF_FWD, O_WF_SIDE, O_ACCEL_X, O_ACCEL_Y, O_OMEGA_Z, O_TOTAL_VELOCITY, O_UNC_POS_X, O_UNC_POS_Y, O_UNC_YAW,
    O_UNC_VEL_X, O_UNC_POS_X_Y, O_UNC_POS_X_YAW, O_UNC_POS_X_VEL_X, O_UNC_POS_Y_YAW, O_UNC_POS_Y_VEL_X, O_UNC_YAW_VEL_X
  };
  enum
  {
    U_VEL_X = 0, U_YAW, U_POS_X, U_POS_Y
  };
  static constexpr int cm(int row, int col)
  {
    return col * 4 + row;  // mm::columnMajorIndex
  }
  // LSTMHelper::resetHiddenCellCPU (lstm_helper.cu:465-473): h, c <- initial_hidden_, initial_cell_
  static void initCarry(const Aux& aux, Carry& k)
  {
    const int H = aux.lstm_hidden, I = MPPIB_RACER_LSTM_INPUT_DIM;
    const float* init = aux.lstm_theta + 4 * H * H + 4 * H * I + 4 * H;
    for (int i = 0; i < H; i++)
    {
      k.h[i] = init[i];
      k.c[i] = init[H + i];
    }
  }
  static int velIndex(float vx)
  {
    const float linear_brake_slope = 0.2f;
    return (fabsf(vx) > linear_brake_slope && fabsf(vx) <= 3.0f) + (fabsf(vx) > 3.0f) * 2;
  }
  static void step(const P& p, const Aux& aux, Carry& k, const float* state, float* next_state, float* state_der,
                   const float* control, float* output, float dt)
  {
    // --- computeParametricDelayDeriv, racer_dubins.cu:306-319
    const bool enable_brake = control[0] < 0.0f;
    const float brake_error = (enable_brake * -control[0] - state[BRAKE_STATE]);
    state_der[BRAKE_STATE] = fminf(fmaxf((brake_error > 0) * brake_error * p.brake_delay_constant +
                                             (brake_error < 0) * brake_error * p.brake_delay_constant_neg,
                                         -p.max_brake_rate_neg),
                                   p.max_brake_rate_pos);
    // --- computeParametricAccelDeriv (host), racer_dubins_elevation.cu:32-67
    {
      const float linear_brake_slope = 0.2f;
      const int index = velIndex(state[VEL_X]);
      const float brake_state = fminf(fmaxf(state[BRAKE_STATE], 0.0f), 0.25f);
      float throttle = p.c_t[index] * control[0];
      float brake = p.c_b[index] * brake_state * (state[VEL_X] >= 0.0f ? -1.0f : 1.0f);
      if (fabsf(state[VEL_X]) <= linear_brake_slope)
      {
        throttle = p.c_t[index] * fmaxf(control[0] - p.low_min_throttle, 0.0f);
        brake = p.c_b[index] * brake_state * -state[VEL_X];
      }
      state_der[VEL_X] = (!enable_brake) * throttle * p.gear_sign + brake - p.c_v[index] * state[VEL_X] + p.c_0;
      state_der[VEL_X] = fminf(fmaxf(state_der[VEL_X], -p.clamp_ax), p.clamp_ax);
      if (fabsf(state[PITCH]) < 1.57079632679489661923f)
        state_der[VEL_X] -= p.gravity * sinf(state[PITCH]);
      state_der[YAW] = (state[VEL_X] / p.wheel_base) * tanf(state[STEER_ANGLE] / p.steer_angle_scale);
      float sin_yaw, cos_yaw;
      sincosf(state[YAW], &sin_yaw, &cos_yaw);
      state_der[POS_X] = state[VEL_X] * cos_yaw;
      state_der[POS_Y] = state[VEL_X] * sin_yaw;
    }
    // --- computeLSTMSteering (host), lstm_steering.cu:66-88
    {
      const float parametric_accel =
          (control[1] * p.steer_command_angle_scale - state[STEER_ANGLE]) * p.steering_constant;
      state_der[STEER_ANGLE_RATE] =
          fmaxf(fminf((parametric_accel - state[STEER_ANGLE_RATE]) * p.steer_accel_constant -
                          state[STEER_ANGLE_RATE] * p.steer_accel_drag_constant,
                      p.max_steer_rate),
                -p.max_steer_rate);
      float input[4], nn_output[1];
      input[0] = state[STEER_ANGLE] * 0.2f;
      input[1] = state[STEER_ANGLE_RATE] * 0.2f;
      input[2] = control[1];
      input[3] = state_der[STEER_ANGLE_RATE] * 0.2f;
      const int H = aux.lstm_hidden, I = MPPIB_RACER_LSTM_INPUT_DIM;
      const int head_layers[3] = { H + I, aux.lstm_head, 1 };
      lstm_forward(aux.lstm_theta, I, H, aux.lstm_theta + 4 * H * H + 4 * H * I + 6 * H, head_layers, 3, input, k.h,
                   k.c, nn_output);
      state_der[STEER_ANGLE_RATE] += nn_output[0] * 5.0f;
      state_der[STEER_ANGLE] = state[STEER_ANGLE_RATE];
    }
    // --- updateState (host), lstm_steering.cu:267-285
    for (int i = 0; i < 6; i++)
      next_state[i] = state[i] + state_der[i] * dt;
    next_state[YAW] = normalizeAngle(next_state[YAW]);
    next_state[STEER_ANGLE] = fmaxf(fminf(next_state[STEER_ANGLE], p.max_steer_angle), -p.max_steer_angle);
    next_state[STEER_ANGLE_RATE] = state[STEER_ANGLE_RATE] + state_der[STEER_ANGLE_RATE] * dt;
    next_state[BRAKE_STATE] = fminf(fmaxf(next_state[BRAKE_STATE], 0.0f), -p.lim.rng_lo[0]);
    // --- computeUncertaintyPropagation, racer_dubins_elevation.cu:662-741
    {
      float A[16], Sa[16], Sb[16];
      // computeUncertaintyJacobian (host trig), :336-425
      float sin_yaw, cos_yaw;
      sincosf(state[YAW], &sin_yaw, &cos_yaw);
      const float delta = state[STEER_ANGLE] / p.steer_angle_scale;
      const float tan_steer_angle = tanf(delta);
      const float cos_2_delta = SQ(cosf(delta));
      const int index = velIndex(state[VEL_X]);
      const float brake_state = fminf(fmaxf(state[BRAKE_STATE], 0.0f), 0.25f);
      A[cm(U_VEL_X, U_VEL_X)] = -p.c_v[index] - p.K_vel_x - (index == 0 ? 1.0f : 0.0f) * p.c_b[0] * brake_state;
      A[cm(U_VEL_X, U_YAW)] = 0.0f;
      A[cm(U_VEL_X, U_POS_X)] = -p.K_x * cos_yaw;
      A[cm(U_VEL_X, U_POS_Y)] = -p.K_x * sin_yaw;
      A[cm(U_YAW, U_VEL_X)] = tan_steer_angle / (p.wheel_base);
      A[cm(U_YAW, U_YAW)] = -fabsf(state[VEL_X]) * p.K_yaw / (p.wheel_base * cos_2_delta);
      A[cm(U_YAW, U_POS_X)] = state[VEL_X] * p.K_y * sin_yaw / (p.wheel_base * cos_2_delta);
      A[cm(U_YAW, U_POS_Y)] = -state[VEL_X] * p.K_y * cos_yaw / (p.wheel_base * cos_2_delta);
      A[cm(U_POS_X, U_VEL_X)] = cos_yaw;
      A[cm(U_POS_X, U_YAW)] = -sin_yaw * state[VEL_X];
      A[cm(U_POS_X, U_POS_X)] = 0.0f;
      A[cm(U_POS_X, U_POS_Y)] = 0.0f;
      A[cm(U_POS_Y, U_VEL_X)] = sin_yaw;
      A[cm(U_POS_Y, U_YAW)] = cos_yaw * state[VEL_X];
      A[cm(U_POS_Y, U_POS_Y)] = 0.0f;
      A[cm(U_POS_Y, U_POS_X)] = 0.0f;
      // uncertaintyStateToMatrix, :517-577
      Sa[cm(U_VEL_X, U_VEL_X)] = state[UNC_VEL_X];
      Sa[cm(U_YAW, U_VEL_X)] = Sa[cm(U_VEL_X, U_YAW)] = state[UNC_YAW_VEL_X];
      Sa[cm(U_POS_X, U_VEL_X)] = Sa[cm(U_VEL_X, U_POS_X)] = state[UNC_POS_X_VEL_X];
      Sa[cm(U_POS_Y, U_VEL_X)] = Sa[cm(U_VEL_X, U_POS_Y)] = state[UNC_POS_Y_VEL_X];
      Sa[cm(U_YAW, U_YAW)] = state[UNC_YAW];
      Sa[cm(U_POS_X, U_YAW)] = Sa[cm(U_YAW, U_POS_X)] = state[UNC_POS_X_YAW];
      Sa[cm(U_POS_Y, U_YAW)] = Sa[cm(U_YAW, U_POS_Y)] = state[UNC_POS_Y_YAW];
      Sa[cm(U_POS_X, U_POS_X)] = state[UNC_POS_X];
      Sa[cm(U_POS_Y, U_POS_X)] = Sa[cm(U_POS_X, U_POS_Y)] = state[UNC_POS_X_Y];
      Sa[cm(U_POS_Y, U_POS_Y)] = state[UNC_POS_Y];
      // A <- I + A dt
      for (int i = 0; i < 16; i++)
        A[i] = (i % 5 == 0) + A[i] * dt;
      // Sigma_a <- A Sigma_a A^T (Eigen: (A * Sigma) * A^T, inner index ascending)
      for (int col = 0; col < 4; col++)
        for (int row = 0; row < 4; row++)
        {
          float acc = 0.0f;
          for (int q = 0; q < 4; q++)
            acc += A[cm(row, q)] * Sa[cm(q, col)];
          Sb[cm(row, col)] = acc;
        }
      for (int col = 0; col < 4; col++)
        for (int row = 0; row < 4; row++)
        {
          float acc = 0.0f;
          for (int q = 0; q < 4; q++)
            acc += Sb[cm(row, q)] * A[cm(col, q)];
          Sa[cm(row, col)] = acc;
        }
      // computeQ, :427-515. The host branch leaves sin_roll unset (only the device branch assigns it, :441); the
      // device value sin(roll) is used here.
      const float abs_vx = fabsf(state[VEL_X]);
      const float abs_acc_x = fabsf(state_der[VEL_X]);
      const float sin_roll = sinf(state[ROLL]);
      const float side_force = SQ(abs_vx) * tan_steer_angle / p.wheel_base + p.gravity * sin_roll;
      const float Q_11 = fabsf(p.Q_y_f * fabsf(side_force) * fmaxf(abs_vx - 2, 0.0f));
      for (int i = 0; i < 16; i++)
        Sb[i] = 0.0f;
      Sb[cm(U_VEL_X, U_VEL_X)] = p.Q_x_acc * abs_acc_x + p.Q_x_v[index] * abs_vx;
      Sb[cm(U_YAW, U_YAW)] = abs_vx * (p.Q_omega_steering * fabsf(delta) + p.Q_omega_v);
      Sb[cm(U_POS_X, U_POS_X)] = Q_11 * sin_yaw * sin_yaw;
      Sb[cm(U_POS_X, U_POS_Y)] = -Q_11 * sin_yaw * cos_yaw;
      Sb[cm(U_POS_Y, U_POS_Y)] = Q_11 * cos_yaw * cos_yaw;
      Sb[cm(U_POS_Y, U_POS_X)] = -Q_11 * sin_yaw * cos_yaw;
      for (int i = 0; i < 16; i++)
        Sa[i] += Sb[i] * dt;
      // uncertaintyMatrixToState, :579-621
      next_state[UNC_VEL_X] = Sa[cm(U_VEL_X, U_VEL_X)];
      next_state[UNC_YAW_VEL_X] = Sa[cm(U_YAW, U_VEL_X)];
      next_state[UNC_POS_X_VEL_X] = Sa[cm(U_POS_X, U_VEL_X)];
      next_state[UNC_POS_Y_VEL_X] = Sa[cm(U_POS_Y, U_VEL_X)];
      next_state[UNC_YAW] = Sa[cm(U_YAW, U_YAW)];
      next_state[UNC_POS_X_YAW] = Sa[cm(U_POS_X, U_YAW)];
      next_state[UNC_POS_Y_YAW] = Sa[cm(U_POS_Y, U_YAW)];
      next_state[UNC_POS_X] = Sa[cm(U_POS_X, U_POS_X)];
      next_state[UNC_POS_X_Y] = Sa[cm(U_POS_Y, U_POS_X)];
      next_state[UNC_POS_Y] = Sa[cm(U_POS_Y, U_POS_Y)];
    }
    // --- static settling (lstm_steering.cu:105-112 -> racer_dubins.cu:359-434); without a map roll = pitch = height = 0
    float roll = state[ROLL], pitch = state[PITCH];
    output[O_POS_I_Z] = staticSettling(aux.elev, next_state[YAW], next_state[POS_X], next_state[POS_Y], roll, pitch);
    next_state[PITCH] = pitch;
    next_state[ROLL] = roll;
    // --- setOutputs, racer_dubins_elevation.cu:69-227
    output[O_VEL_B_X] = next_state[VEL_X];
    output[O_VEL_B_Y] = 0.0f;
    output[O_POS_I_X] = next_state[POS_X];
    output[O_POS_I_Y] = next_state[POS_Y];
    output[O_PITCH] = next_state[PITCH];
    output[O_ROLL] = next_state[ROLL];
    output[O_YAW] = next_state[YAW];
    output[O_STEER_ANGLE] = next_state[STEER_ANGLE];
    output[O_STEER_ANGLE_RATE] = next_state[STEER_ANGLE_RATE];
    output[O_WF_UP] = NAN;
    output[O_WF_FWD] = NAN;
    output[O_WF_SIDE] = NAN;
    output[O_ACCEL_X] = state_der[VEL_X];
    output[O_ACCEL_Y] = 0.0f;
    output[O_OMEGA_Z] = state_der[YAW];
    output[O_UNC_VEL_X] = next_state[UNC_VEL_X];
    output[O_UNC_YAW_VEL_X] = next_state[UNC_YAW_VEL_X];
    output[O_UNC_POS_X_VEL_X] = next_state[UNC_POS_X_VEL_X];
    output[O_UNC_POS_Y_VEL_X] = next_state[UNC_POS_Y_VEL_X];
    output[O_UNC_YAW] = next_state[UNC_YAW];
    output[O_UNC_POS_X_YAW] = next_state[UNC_POS_X_YAW];
    output[O_UNC_POS_Y_YAW] = next_state[UNC_POS_Y_YAW];
    output[O_UNC_POS_X] = next_state[UNC_POS_X];
    output[O_UNC_POS_X_Y] = next_state[UNC_POS_X_Y];
    output[O_UNC_POS_Y] = next_state[UNC_POS_Y];
    output[O_TOTAL_VELOCITY] = fabsf(next_state[VEL_X]);
  }
};

// QuadrotorDynamics, host path: dynamics/quadrotor/quadrotor_dynamics.cu:70-112 (computeDynamics, Eigen) and :114-122
// (updateState: Euler step, then the quaternion divided by norm * copysign(1, q.w)). Quat2DCM(Eigen) is
// Eigen::Quaternionf::toRotationMatrix (utils/math_utils.h:529-532); Eigen is a system dependency of the reference
// (CMakeLists.txt find_package(Eigen3)), absent from /root/reference, so its published column-2 formulas
// (Eigen/src/Geometry/Quaternion.h, QuaternionBase::toRotationMatrix) are restated here.
struct Quadrotor
{
  static constexpr int S = 13, C = 4, O = 13;
  static constexpr bool CUSTOM_STEP = true;
  typedef mppib_quadrotor_dyn_params P;
  typedef NoCarry Carry;
  static void initCarry(const Aux&, Carry&)
  {
  }
  static void computeStateDeriv(const P& p, const Aux&, const float* state, const float* control, float* state_der)
  {
    const float u_thrust = control[3];
    const float qw = state[6], qx = state[7], qy = state[8], qz = state[9];
    // toRotationMatrix, column 2
    const float tx = 2.0f * qx, ty = 2.0f * qy, tz = 2.0f * qz;
    const float twx = tx * qw, twy = ty * qw;
    const float txx = tx * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy;
    const float col2[3] = { txz + twy, tyz - twx, 1.0f - (txx + tyy) };
    const float tau_inv[3] = { 1 / p.tau_roll, 1 / p.tau_pitch, 1 / p.tau_yaw };
    for (int i = 0; i < 3; i++)
      state_der[i] = state[3 + i];
    for (int i = 0; i < 3; i++)
      state_der[3 + i] = (u_thrust / p.mass) * col2[i];
    state_der[5] -= MPPIB_GRAVITY;
    // omega2edot (utils/math_utils.h:543-549)
    const float pp = state[10], qq = state[11], rr = state[12];
    state_der[6] = 0.5f * (-pp * qx - qq * qy - rr * qz);
    state_der[7] = 0.5f * (pp * qw - qq * qz + rr * qy);
    state_der[8] = 0.5f * (pp * qz + qq * qw - rr * qx);
    state_der[9] = 0.5f * (-pp * qy + qq * qx + rr * qw);
    for (int i = 0; i < 3; i++)
      state_der[10 + i] = tau_inv[i] * (control[i] - state[10 + i]);
  }
  static void step(const P& p, const Aux& aux, Carry&, const float* state, float* next_state, float* state_der,
                   const float* control, float* output, float dt)
  {
    computeStateDeriv(p, aux, state, control, state_der);
    for (int i = 0; i < S; i++)
      next_state[i] = state[i] + state_der[i] * dt;
    const float* q = next_state + 6;
    const float norm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const float div = (float)((double)norm * copysign(1.0, (double)q[0]));  // Eigen `/=` takes a float Scalar
    for (int i = 6; i < 10; i++)
      next_state[i] /= div;
    for (int i = 0; i < O; i++)
      output[i] = next_state[i];
  }
};

template <class T, class = void>
struct has_custom_step : std::false_type
{
};
template <class T>
struct has_custom_step<T, std::enable_if_t<T::CUSTOM_STEP>> : std::true_type
{
};

// dynamics/dynamics.cuh:277-300
template <class DYN>
static inline void dyn_step(const typename DYN::P& p, const Aux& aux, const float* state, float* next_state,
                            float* state_der, const float* control, float* output, float dt,
                            typename DYN::Carry* carry = nullptr)
{
  for (int i = 0; i < DYN::S; i++)
    state_der[i] = 0.0f;  // Eigen state_array locals are written fully by every model used here
  if constexpr (has_custom_step<DYN>::value)
  {
    DYN::step(p, aux, *carry, state, next_state, state_der, control, output, dt);
  }
  else
  {
    DYN::computeStateDeriv(p, aux, state, control, state_der);
    for (int i = 0; i < DYN::S; i++)
      next_state[i] = state[i] + state_der[i] * dt;
    for (int i = 0; i < DYN::O && i < DYN::S; i++)
      output[i] = next_state[i];
  }
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

// Ours (params.h: mppib_racer_quadratic_cost_params) — the RACER cost classes are not in the reference tree; this is
// the specification the GPU twin is checked against, not a restatement.
struct RacerQuadraticCost
{
  typedef mppib_racer_quadratic_cost_params P;
  static float computeStateCost(const P& p, const Aux&, const float* y, int t, int*)
  {
    const float dv = y[RacerLSTM::O_VEL_B_X] - p.desired_speed;
    const float dyaw = normalizeAngle(y[RacerLSTM::O_YAW] - p.desired_yaw);
    const float dy = y[RacerLSTM::O_POS_I_Y] - p.desired_y;
    const float st = y[RacerLSTM::O_STEER_ANGLE];
    const float cost = p.speed_coeff * dv * dv + p.yaw_coeff * dyaw * dyaw + p.lateral_coeff * dy * dy +
                       p.steer_coeff * st * st;
    return cost * powf(p.discount, (float)t);
  }
  static float terminalCost(const P&, const Aux&, const float*)
  {
    return 0.0f;
  }
};

// QuadrotorQuadraticCost, host path: cost_functions/quadrotor/quadrotor_quadratic_cost.cu:11-65. QuatSubtract(Eigen) is
// q_2 * q_1.inverse() (utils/math_utils.h:285-289): Eigen's inverse = conjugate / squaredNorm and its Hamilton product,
// not normalised. The !use_euler branch of the host code assigns a 4-vector to a Vector3f (ill-formed at run time), so for
// that branch this oracle follows the DEVICE body (:70-126: q_coeff * q_diff[i], unsquared, with the device QuatSubtract of
// math_utils.h:166-211); tests use the default use_euler = true.
struct QuadrotorQuadraticCost
{
  typedef mppib_quadrotor_cost_params P;
  static float computeStateCost(const P& p, const Aux&, const float* s, int, int*)
  {
    float x_cost = 0, v_cost = 0, w_cost = 0, q_cost = 0;
    for (int i = 0; i < 3; i++)
    {
      x_cost += p.x_coeff * ((s[i] - p.s_goal[i]) * (s[i] - p.s_goal[i]));
      v_cost += p.v_coeff * ((s[3 + i] - p.s_goal[3 + i]) * (s[3 + i] - p.s_goal[3 + i]));
      w_cost += p.w_coeff * ((s[10 + i] - p.s_goal[10 + i]) * (s[10 + i] - p.s_goal[10 + i]));
    }
    const float* q = s + 6;
    const float* g = p.s_goal + 6;
    if (p.use_euler)
    {
      const float n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
      const float iw = q[0] / n2, ix = -q[1] / n2, iy = -q[2] / n2, iz = -q[3] / n2;
      // q_diff = q_g * q^-1
      const float dw = g[0] * iw - g[1] * ix - g[2] * iy - g[3] * iz;
      const float dx = g[0] * ix + g[1] * iw + g[2] * iz - g[3] * iy;
      const float dy = g[0] * iy + g[2] * iw + g[3] * ix - g[1] * iz;
      const float dz = g[0] * iz + g[3] * iw + g[1] * iy - g[2] * ix;
      // Quat2EulerNWU (utils/math_utils.h:519-527)
      const float r = atan2f(2.0f * dz * dy + 2.0f * dw * dx, dw * dw + dz * dz - dy * dy - dx * dx);
      const float temp = -2.0f * dw * dy + 2.0f * dx * dz;
      const float pi = -asinf(fmaxf(-1.0f, fminf(temp, 1.0f)));
      const float y = atan2f(2.0f * dy * dx + 2.0f * dz * dw, dw * dw + dx * dx - dy * dy - dz * dz);
      q_cost = p.roll_coeff * (r * r) + p.pitch_coeff * (pi * pi) + p.yaw_coeff * (y * y);
    }
    else
    {
      const float inv_norm = 1.0f / sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      const float a[4] = { q[0] * inv_norm, -q[1] * inv_norm, -q[2] * inv_norm, -q[3] * inv_norm };
      float d[4];
      d[0] = g[0] * a[0] - g[1] * a[1] - g[2] * a[2] - g[3] * a[3];
      d[1] = g[1] * a[0] + g[0] * a[1] - g[3] * a[2] + g[2] * a[3];
      d[2] = g[2] * a[0] + g[3] * a[1] + g[0] * a[2] - g[1] * a[3];
      d[3] = g[3] * a[0] - g[2] * a[1] + g[1] * a[2] + g[0] * a[3];
      const float dn = 1.0f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3]);
      for (int i = 0; i < 4; i++)
        q_cost += p.q_coeff * (d[i] * dn);
    }
    return x_cost + v_cost + q_cost + w_cost;
  }
  static float terminalCost(const P& p, const Aux& a, const float* s)
  {
    return p.terminal_cost_coeff * computeStateCost(p, a, s, 0, nullptr);
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
      typename DYN::Carry carry;
      DYN::initCarry(aux, carry);
      for (int t = 0; t < T; t++)
      {
        float* us = &samples[(((size_t)d * N + n) * T + t) * C];
        for (int i = 0; i < C; i++)
          u[i] = us[i];
        enforceConstraints<C>(dp.lim, u);
        for (int i = 0; i < C; i++)
          us[i] = u[i];
        dyn_step<DYN>(dp, aux, curr_x, next_x, x_der, u, y, dt, &carry);
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

// One rollout with everything dumped: the per-step semantics of launchCPURolloutKernel above (and of visualizeKernel,
// core/mppi_common.cu:364-520: step, then running cost + likelihood-ratio cost on the new output, crash flag sticky),
// outputs [T][O] = y after step t, costs [T + 1] = per-step cost / T with the terminal cost / T last, crash [T].
template <class DYN, class COST>
static void sampledTrajectory(const void* dpv, const void* cpv, const mppib_gaussian_params& sp, const Aux& aux, int N, int T,
                              int d, int sample_index, int apply_constraints, float dt, float lambda, float alpha,
                              const float* x0, const float* means /*[T][C] of d*/, const float* controls /*[T][C]*/,
                              float* outputs, float* costs, int* crash)
{
  constexpr int S = DYN::S, C = DYN::C, O = DYN::O;
  const auto& dp = *(const typename DYN::P*)dpv;
  const auto& cp = *(const typename COST::P*)cpv;
  float curr_x[S], next_x[S], x_der[S], u[C], y[O];
  for (int i = 0; i < S; i++)
    curr_x[i] = x0[i];
  for (int i = 0; i < O; i++)
    y[i] = 0.0f;
  int crash_status = 0;
  typename DYN::Carry carry;
  DYN::initCarry(aux, carry);
  for (int t = 0; t < T; t++)
  {
    for (int i = 0; i < C; i++)
      u[i] = controls[(size_t)t * C + i];
    if (apply_constraints)
      enforceConstraints<C>(dp.lim, u);
    dyn_step<DYN>(dp, aux, curr_x, next_x, x_der, u, y, dt, &carry);
    float c = COST::computeStateCost(cp, aux, y, t, &crash_status);
    c += likelihoodRatioCost(sp, &means[(size_t)t * C], u, C, d, sample_index, N, lambda, alpha);
    costs[t] = c / T;
    crash[t] = crash_status;
    for (int i = 0; i < O; i++)
      outputs[(size_t)t * O + i] = y[i];
    for (int i = 0; i < S; i++)
      curr_x[i] = next_x[i];
  }
  costs[T] = COST::terminalCost(cp, aux, y) / T;
}

// ---- persistent worker pool -----------------------------------------------------------------------------------------
// The CPU arm of the bench times back-to-back solves; spawning and joining ~128 std::threads per solve, unpinned, made that
// number swing 4x between boxes (round-1 verdict). The workers below are created once, pinned one per core
// (pthread_setaffinity_np), and woken per parallel region; chunk i of a region always runs on worker i.
namespace
{
class WorkerPool
{
public:
  explicit WorkerPool(int n) : n_(n), gen_(0), pending_(0), stop_(false)
  {
    const int ncpu = (int)std::max(1u, std::thread::hardware_concurrency());
    for (int i = 0; i < n_; i++)
    {
      th_.emplace_back([this, i]() { loop(i); });
      cpu_set_t set;
      CPU_ZERO(&set);
      CPU_SET(i % ncpu, &set);
      pthread_setaffinity_np(th_.back().native_handle(), sizeof(set), &set);
    }
  }
  ~WorkerPool()
  {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto& t : th_)
      t.join();
  }
  int size() const
  {
    return n_;
  }
  void run(const std::function<void(int)>& fn)
  {
    std::unique_lock<std::mutex> lk(m_);
    fn_ = &fn;
    pending_ = n_;
    gen_++;
    cv_.notify_all();
    done_.wait(lk, [this]() { return pending_ == 0; });
    fn_ = nullptr;
  }

private:
  void loop(int i)
  {
    unsigned long seen = 0;
    for (;;)
    {
      const std::function<void(int)>* fn;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&]() { return gen_ != seen; });
        seen = gen_;
        if (stop_)
          return;
        fn = fn_;
      }
      (*fn)(i);
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_ == 0)
          done_.notify_one();
      }
    }
  }
  int n_;
  unsigned long gen_;
  int pending_;
  bool stop_;
  const std::function<void(int)>* fn_ = nullptr;
  std::mutex m_;
  std::condition_variable cv_, done_;
  std::vector<std::thread> th_;
};
// fn(chunk, nchunks) for chunk = 0 .. nthreads-1 on the pool (re-created when a different width is asked for)
static void parallel_chunks(int nthreads, const std::function<void(int)>& fn)
{
  static std::mutex guard;
  static WorkerPool* pool = nullptr;
  std::lock_guard<std::mutex> lk(guard);
  if (pool == nullptr || pool->size() != nthreads)
  {
    delete pool;
    pool = new WorkerPool(nthreads);
  }
  pool->run(fn);
}
}  // namespace

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
  parallel_chunks(nthreads, [=, &d, &c, &sp, &aux](int i) {
    int b = (int)((long long)N * i / nthreads), e = (int)((long long)N * (i + 1) / nthreads);
    rollout_range<DYN, COST>(d, c, sp, aux, N, T, D, dt, lambda, alpha, x0, means, samples, costs, b, e);
  });
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
  if (dyn_id == MPPIB_DYN_RACER_LSTM && cost_id == MPPIB_COST_RACER_QUADRATIC)
    return &rollout<RacerLSTM, RacerQuadraticCost>;
  if (dyn_id == MPPIB_DYN_QUADROTOR && cost_id == MPPIB_COST_QUADROTOR_QUADRATIC)
    return &rollout<Quadrotor, QuadrotorQuadraticCost>;
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
    case MPPIB_DYN_RACER_LSTM:
      *S = 19, *C = 2, *O = 28;
      break;
    case MPPIB_DYN_QUADROTOR:
      *S = 13, *C = 4, *O = 13;
      break;
    default:
      *S = *C = *O = 0;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// ColoredNoiseDistribution::generateSamples up to (not including) setGaussianControls,
// sampling_distributions/colored_noise/colored_noise.cu:286-372 with its kernels :12-56; the algorithm is
// scripts/colored_noise.py:12-104 (Timmer & Koenig). `normals` is what curandGenerateNormal wrote into
// samples_in_freq_complex_d_: [n][c][freq] complex = 2 * N * C * (T + 1) floats. cuFFT's C2R (unnormalised inverse real
// DFT of length 2T, libcufft — third party) is restated as the defining sum in double precision.
static void coloredNoiseTables(const mppib_gaussian_params& sp, int C, int T, std::vector<float>& sample_freqs /*[c][f]*/,
                               float* sigma)
{
  const int sample_num_timesteps = 2 * T;
  const int freq_size = sample_num_timesteps / 2 + 1;  // fftfreq, colored_noise.cuh:24-34
  std::vector<float> sample_freq(freq_size);
  for (int i = 0; i < freq_size; i++)
    sample_freq[i] = i / (1.0f * sample_num_timesteps);
  const float cutoff_freq = fmaxf(sp.fmin, 1.0f / sample_num_timesteps);
  int smaller_index = 0;
  sample_freqs.assign((size_t)freq_size * C, 0.0f);  // Eigen MatrixXf(freq_size, C), column-major => [c][f]
  auto at = [&](int f, int c) -> float& { return sample_freqs[(size_t)c * freq_size + f]; };
  for (int i = 0; i < freq_size; i++)  // colored_noise.cu:305-326
  {
    if (sample_freq[i] < cutoff_freq)
    {
      smaller_index++;
    }
    else if (smaller_index < freq_size)
    {
      for (int j = 0; j < smaller_index; j++)
      {
        sample_freq[j] = sample_freq[smaller_index];
        for (int k = 0; k < C; k++)
          at(j, k) = powf(sample_freq[smaller_index], -sp.exponents[k] / 2.0f);
      }
    }
    for (int j = 0; j < C; j++)
      at(i, j) = powf(sample_freq[i], -sp.exponents[j] / 2.0f);
  }
  for (int i = 0; i < C; i++)  // :329-338
  {
    sigma[i] = 0.0f;
    for (int j = 1; j < freq_size - 1; j++)
      sigma[i] += SQ(at(j, i));
    sigma[i] += SQ(at(freq_size - 1, i) * ((1.0f + (sample_num_timesteps % 2)) / 2.0f));
    sigma[i] = 2.0f * sqrtf(sigma[i]) / sample_num_timesteps;
  }
}

static void coloredNoise(const float* normals, const mppib_gaussian_params& sp, int N, int C, int T, int offset_t,
                         float* eps /*[N][T][C]*/, int n_begin, int n_end)
{
  const int n2 = 2 * T, F = T + 1;
  std::vector<float> coeff;
  float sigma[MPPIB_MAX_CONTROL_DIM];
  coloredNoiseTables(sp, C, T, coeff, sigma);
  std::vector<double> cs(n2), sn(n2);
  for (int i = 0; i < n2; i++)
  {
    cs[i] = cos(2.0 * M_PI * i / n2);
    sn[i] = sin(2.0 * M_PI * i / n2);
  }
  std::vector<float> re(F), im(F), time(T);
  for (int n = n_begin; n < n_end; n++)
    for (int c = 0; c < C; c++)
    {
      const float* row = normals + ((size_t)n * C + c) * F * 2;
      for (int f = 0; f < F; f++)  // configureFrequencyNoise, :12-37
      {
        const float v = coeff[(size_t)c * F + f];
        re[f] = row[2 * f] * v;
        if (f == 0)
          im[f] = 0.0f;
        else if (F % 2 == 1 && f == F - 1)
          im[f] = 0.0f;
        else
          im[f] = row[2 * f + 1] * v;
      }
      for (int t = 0; t < T; t++)  // cufftExecC2R, first T of the 2T outputs; imaginary parts of DC / Nyquist ignored
      {
        double acc = (double)re[0] + ((t & 1) ? -(double)re[T] : (double)re[T]);
        for (int k = 1; k < T; k++)
        {
          const int a = (int)(((long long)k * t) % n2);
          acc += 2.0 * ((double)re[k] * cs[a] - (double)im[k] * sn[a]);
        }
        time[t] = (float)acc;
      }
      for (int t = 0; t < T; t++)  // rearrangeNoise, :39-56
      {
        const float decayed_offset = sp.offset_decay_rate == 0 ? 0 : powf(sp.offset_decay_rate, t);
        eps[((size_t)n * T + t) * C + c] = (time[t] - time[offset_t] * decayed_offset) / (sigma[c] * 2 * T);
      }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// RMPPI. tests/include/kernel_tests/core/rmppi_kernel_test.cu:7-77 (launchCPURMPPIRolloutKernel) and :79-127
// (launchCPUInitEvalKernel); computeFeedbackCost gaussian.cu:572-629; DDP feedback k(x, x*, t) = K_t (x - x*) with K_t
// column-major C x S (feedback_controllers/DDP/ddp.cu:11-45 — the HOST form: the device code overwrites instead of
// accumulating when CONTROL_DIM is even, which keeps only the last state's column; the intended product is restated).
static inline float feedbackCost(const mppib_gaussian_params& sp, const float* u_fb, int C, int d, float lambda,
                                 float alpha)
{
  float cost = 0.0f;
  for (int i = 0; i < C; i++)
  {
    const float sd = sp.std_dev[d * C + i];
    cost += sp.control_cost_coeff[i] * (u_fb[i] * u_fb[i]) / (sd * sd);
  }
  return 0.5f * lambda * (1.0f - alpha) * cost;
}

// samples [2][N][T][C]: sampled controls of both distributions (index nominal_idx / real_idx), constrained in place.
template <class DYN, class COST>
static void rmppiRollout(const void* dpv, const void* cpv, const mppib_gaussian_params& sp, const Aux& aux, int N, int T,
                         float dt, float lambda, float alpha, float value_func_threshold, int nominal_idx,
                         const float* x0 /*[2][S]*/, const float* means /*[2][T][C]*/,
                         const float* gains /*[T][S][C] == column-major C x S per t, may be null*/, float* samples,
                         float* costs /*[2][N]*/, int n_begin, int n_end)
{
  const auto& dp = *(const typename DYN::P*)dpv;
  const auto& cp = *(const typename COST::P*)cpv;
  constexpr int S = DYN::S, C = DYN::C, O = DYN::O;
  const int real_idx = 1 - nominal_idx;
  for (int n = n_begin; n < n_end; n++)
  {
    float xr[S], xn[S], xr_next[S], xn_next[S], dr[S], dn[S], yr[O], yn[O], ur[C], un[C], ufb[C];
    for (int i = 0; i < S; i++)
    {
      xr[i] = x0[real_idx * S + i];
      xn[i] = x0[nominal_idx * S + i];
    }
    for (int i = 0; i < O; i++)
      yr[i] = yn[i] = 0.0f;
    int crash_r = 0, crash_n = 0;
    float running_real = 0, running_nom = 0, tracking_nom = 0, tracking_real = 0;
    typename DYN::Carry kr, kn;
    DYN::initCarry(aux, kr);
    DYN::initCarry(aux, kn);
    for (int t = 0; t < T; t++)
    {
      float* sr = &samples[(((size_t)real_idx * N + n) * T + t) * C];
      float* sn = &samples[(((size_t)nominal_idx * N + n) * T + t) * C];
      for (int i = 0; i < C; i++)
      {
        ur[i] = sr[i];
        un[i] = sn[i];
        ufb[i] = 0.0f;
      }
      if (gains)
        for (int i = 0; i < S; i++)
        {
          const float e = xr[i] - xn[i];
          for (int j = 0; j < C; j++)
            ufb[j] += gains[((size_t)t * S + i) * C + j] * e;
        }
      for (int i = 0; i < C; i++)
        ur[i] += ufb[i];
      enforceConstraints<C>(dp.lim, ur);
      enforceConstraints<C>(dp.lim, un);
      for (int i = 0; i < C; i++)
      {
        sr[i] = ur[i];
        sn[i] = un[i];
      }
      dyn_step<DYN>(dp, aux, xr, xr_next, dr, ur, yr, dt, &kr);
      dyn_step<DYN>(dp, aux, xn, xn_next, dn, un, yn, dt, &kn);
      const float real_cost = COST::computeStateCost(cp, aux, yr, t, &crash_r);
      const float nom_cost = COST::computeStateCost(cp, aux, yn, t, &crash_n);
      tracking_real += real_cost + feedbackCost(sp, ufb, C, real_idx, lambda, alpha);
      running_real += real_cost + likelihoodRatioCost(sp, &means[((size_t)real_idx * T + t) * C], ur, C, real_idx, n, N,
                                                      lambda, alpha);
      running_nom += nom_cost;
      tracking_nom += likelihoodRatioCost(sp, &means[((size_t)nominal_idx * T + t) * C], un, C, nominal_idx, n, N,
                                          lambda, alpha);
      for (int i = 0; i < S; i++)
      {
        xr[i] = xr_next[i];
        xn[i] = xn_next[i];
      }
    }
    running_real += COST::terminalCost(cp, aux, yr);
    tracking_real += COST::terminalCost(cp, aux, yr);
    running_nom += COST::terminalCost(cp, aux, yn);
    tracking_nom /= T;
    tracking_real /= T;
    running_nom /= T;
    running_real /= T;
    running_nom = 0.5f * running_nom + 0.5f * fmaxf(fminf(tracking_real, value_func_threshold), running_nom);
    running_nom += tracking_nom;
    costs[(size_t)nominal_idx * N + n] = running_nom;
    costs[(size_t)real_idx * N + n] = running_real;
  }
}

// controls [num_samples][T][C]: the sampler's buffer after setGaussianControls (distribution 0); constrained copies are
// not written back (the reference's CPU oracle does not either).
template <class DYN, class COST>
static void initEval(const void* dpv, const void* cpv, const mppib_gaussian_params& sp, const Aux& aux, int N_sampler,
                     int T, float dt, float lambda, float alpha, int num_candidates, int num_samples,
                     const float* candidates /*[K][S]*/, const int* strides, const float* means /*[T][C]*/,
                     const float* controls, float* costs /*[K * num_samples]*/)
{
  const auto& dp = *(const typename DYN::P*)dpv;
  const auto& cp = *(const typename COST::P*)cpv;
  constexpr int S = DYN::S, C = DYN::C, O = DYN::O;
  for (int k = 0; k < num_candidates; k++)
    for (int j = 0; j < num_samples; j++)
    {
      const int global_idx = k * num_samples + j;
      float x[S], xn[S], xd[S], y[O], u[C];
      for (int i = 0; i < S; i++)
        x[i] = candidates[k * S + i];
      for (int i = 0; i < O; i++)
        y[i] = 0.0f;
      int crash = 0;
      float running = 0.0f;
      typename DYN::Carry carry;
      DYN::initCarry(aux, carry);
      for (int t = 0; t < T; t++)
      {
        const int ct = std::min(t + strides[k], T - 1);
        for (int i = 0; i < C; i++)
          u[i] = controls[((size_t)j * T + ct) * C + i];
        enforceConstraints<C>(dp.lim, u);
        dyn_step<DYN>(dp, aux, x, xn, xd, u, y, dt, &carry);
        running += COST::computeStateCost(cp, aux, y, t, &crash);
        // device call: computeLikelihoodRatioCost(u, theta_d, global_idx, t, 0, lambda, alpha) (rmppi_kernels.cu:331-333)
        running += likelihoodRatioCost(sp, &means[(size_t)t * C], u, C, 0, global_idx, N_sampler, lambda, alpha);
        for (int i = 0; i < S; i++)
          x[i] = xn[i];
      }
      running += COST::terminalCost(cp, aux, y);
      running /= T;
      costs[global_idx] = running;
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
// core/mppi_common.cu:968-985 (TsallisTransform), the weighting ColoredMPPIController uses when gamma and r are both
// non-zero (ColoredMPPI/colored_mppi_controller.cu:199-209)
static void tsallisTransform(float* c, int n, float gamma, float r, float baseline)
{
  for (int i = 0; i < n; i++)
  {
    float cost_dif = c[i] - baseline;
    if (cost_dif < gamma)
      c[i] = expf(logf(1.0 - cost_dif / gamma) / (r - 1));
    else
      c[i] = 0;
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
  typename DYN::Carry carry;
  DYN::initCarry(aux, carry);
  for (int t = 0; t < T - 1; ++t)
  {
    for (int i = 0; i < S; i++)
      state[i] = states[(size_t)t * S + i];
    for (int i = 0; i < C; i++)
      ui[i] = u[(size_t)t * C + i];
    enforceConstraints<C>(dp.lim, ui);
    dyn_step<DYN>(dp, aux, state, next_state, xdot, ui, output, dt, &carry);
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

// LSTM weights / architecture for MPPIB_DYN_RACER_LSTM: set once, used by every later call (kept by pointer).
static orc::Aux g_lstm;
void orc_set_lstm(const float* theta, int hidden_dim, int head_hidden)
{
  g_lstm.lstm_theta = theta;
  g_lstm.lstm_hidden = hidden_dim;
  g_lstm.lstm_head = head_hidden;
}
// elevation map of the RACER model (mppib_elevation_map_header + values, kept by pointer; nullptr = flat ground)
void orc_set_elevation_map(const void* blob)
{
  g_lstm.elev = static_cast<const mppib_elevation_map_header*>(blob);
}
float orc_elevation_at_world_pose(const void* blob, float x, float y, float z)
{
  return orc::elevationAtWorldPose(static_cast<const mppib_elevation_map_header*>(blob), x, y, z);
}
float orc_static_settling(const void* blob, float yaw, float x, float y, float* roll, float* pitch)
{
  return orc::staticSettling(static_cast<const mppib_elevation_map_header*>(blob), yaw, x, y, *roll, *pitch);
}
static void fill_lstm(orc::Aux& aux)
{
  aux.elev = g_lstm.elev;
  aux.lstm_theta = g_lstm.lstm_theta;
  aux.lstm_hidden = g_lstm.lstm_hidden;
  aux.lstm_head = g_lstm.lstm_head;
}

// LSTMHelper::forward(input, output) host path; h and c are updated in place. head_layers as in FNNHelper.
void orc_lstm_forward(const float* lstm_w, int input_dim, int hidden_dim, const float* head_theta,
                      const int* head_layers, int head_num_layers, const float* input, float* h, float* c,
                      float* output)
{
  orc::lstm_forward(lstm_w, input_dim, hidden_dim, head_theta, head_layers, head_num_layers, input, h, c, output);
}

// LSTMLSTMHelper::initializeLSTM (lstm_lstm_helper.cu:50-73): reset the init model to its initial hidden / cell state, run it
// over the last init_len columns of the buffer (forward without output, then forward with output on the last column), the
// output's head / tail are the prediction LSTM's hidden / cell state. buffer [cols][input_dim] (column t of the reference's
// matrix), out [2 * H_prediction].
void orc_lstm_initialize(const float* init_w, int input_dim, int hidden_dim, const float* head_theta, const int* head_layers,
                         int head_num_layers, int init_len, const float* buffer, int cols, float* out)
{
  const float* init = init_w + 4 * hidden_dim * hidden_dim + 4 * hidden_dim * input_dim + 4 * hidden_dim;
  float h[128], c[128], scratch[512];
  for (int i = 0; i < hidden_dim; i++)
  {
    h[i] = init[i];
    c[i] = init[hidden_dim + i];
  }
  int t = cols - init_len;
  for (; t < cols - 1; t++)
    orc::lstm_forward(init_w, input_dim, hidden_dim, head_theta, head_layers, head_num_layers,
                      buffer + (size_t)t * input_dim, h, c, scratch);
  orc::lstm_forward(init_w, input_dim, hidden_dim, head_theta, head_layers, head_num_layers, buffer + (size_t)t * input_dim, h,
                    c, out);
}

// One host step of RacerDubinsElevationLSTMSteering with explicit hidden/cell state (updated in place).
int orc_racer_step(const mppib_racer_lstm_dyn_params* p, const float* x, const float* u, float dt, float* h, float* c,
                   float* x_next, float* xdot, float* y)
{
  orc::Aux aux;
  fill_lstm(aux);
  if (!aux.lstm_theta)
    return -1;
  orc::RacerLSTM::Carry k;
  memcpy(k.h, h, sizeof(float) * aux.lstm_hidden);
  memcpy(k.c, c, sizeof(float) * aux.lstm_hidden);
  orc::dyn_step<orc::RacerLSTM>(*p, aux, x, x_next, xdot, u, y, dt, &k);
  memcpy(h, k.h, sizeof(float) * aux.lstm_hidden);
  memcpy(c, k.c, sizeof(float) * aux.lstm_hidden);
  return 0;
}

// ColoredNoiseDistribution noise block (before setGaussianControls) from the raw normals of one generateSamples call.
void orc_colored_noise(const float* normals, const mppib_gaussian_params* sp, int N, int C, int T, int offset_t,
                       float* eps, int nthreads)
{
  if (nthreads <= 1)
  {
    orc::coloredNoise(normals, *sp, N, C, T, offset_t, eps, 0, N);
    return;
  }
  std::vector<std::thread> th;
  for (int i = 0; i < nthreads; i++)
  {
    int b = (int)((long long)N * i / nthreads), e = (int)((long long)N * (i + 1) / nthreads);
    th.emplace_back([=]() { orc::coloredNoise(normals, *sp, N, C, T, offset_t, eps, b, e); });
  }
  for (auto& t : th)
    t.join();
}

void orc_colored_tables(const mppib_gaussian_params* sp, int C, int T, float* coeffs /*[C][T+1]*/, float* sigma)
{
  std::vector<float> c;
  orc::coloredNoiseTables(*sp, C, T, c, sigma);
  memcpy(coeffs, c.data(), c.size() * sizeof(float));
}

// RMPPI rollout (both systems) on already-sampled controls [2][N][T][C]; costs [2][N]. Returns 0 / -1 (unknown pair).
int orc_rmppi_rollout(int dyn_id, int cost_id, const void* dyn_params, const void* cost_params,
                      const mppib_gaussian_params* sp, const float* nn_theta, const float* costmap, int N, int T,
                      float dt, float lambda, float alpha, float value_func_threshold, int nominal_idx, const float* x0,
                      const float* means, const float* gains, float* samples, float* costs, int nthreads)
{
  orc::Aux aux;
  aux.nn_theta = nn_theta;
  aux.costmap = costmap;
  fill_lstm(aux);
  typedef void (*fn_t)(const void*, const void*, const mppib_gaussian_params&, const orc::Aux&, int, int, float, float,
                       float, float, int, const float*, const float*, const float*, float*, float*, int, int);
  fn_t f = nullptr;
  if (dyn_id == MPPIB_DYN_CARTPOLE && cost_id == MPPIB_COST_CARTPOLE_QUADRATIC)
    f = &orc::rmppiRollout<orc::Cartpole, orc::CartpoleQuadraticCost>;
  else if (dyn_id == MPPIB_DYN_DOUBLE_INTEGRATOR && cost_id == MPPIB_COST_DI_CIRCLE)
    f = &orc::rmppiRollout<orc::DoubleIntegrator, orc::DICircleCost>;
  else if (dyn_id == MPPIB_DYN_AUTORALLY_NN && cost_id == MPPIB_COST_AR_STANDARD)
    f = &orc::rmppiRollout<orc::AutorallyNN, orc::ARStandardCost>;
  else if (dyn_id == MPPIB_DYN_QUADROTOR && cost_id == MPPIB_COST_QUADROTOR_QUADRATIC)
    f = &orc::rmppiRollout<orc::Quadrotor, orc::QuadrotorQuadraticCost>;
  if (!f)
    return -1;
  if (nthreads <= 1)
  {
    f(dyn_params, cost_params, *sp, aux, N, T, dt, lambda, alpha, value_func_threshold, nominal_idx, x0, means, gains,
      samples, costs, 0, N);
    return 0;
  }
  std::vector<std::thread> th;
  for (int i = 0; i < nthreads; i++)
  {
    int b = (int)((long long)N * i / nthreads), e = (int)((long long)N * (i + 1) / nthreads);
    th.emplace_back([=, &aux]() {
      f(dyn_params, cost_params, *sp, aux, N, T, dt, lambda, alpha, value_func_threshold, nominal_idx, x0, means, gains,
        samples, costs, b, e);
    });
  }
  for (auto& t : th)
    t.join();
  return 0;
}

int orc_sampled_trajectory(int dyn_id, int cost_id, const void* dyn_params, const void* cost_params,
                           const mppib_gaussian_params* sp, const float* nn_theta, const float* costmap, int N, int T, int d,
                           int sample_index, int apply_constraints, float dt, float lambda, float alpha, const float* x0,
                           const float* means, const float* controls, float* outputs, float* costs, int* crash)
{
  orc::Aux aux;
  aux.nn_theta = nn_theta;
  aux.costmap = costmap;
  fill_lstm(aux);
  typedef void (*fn_t)(const void*, const void*, const mppib_gaussian_params&, const orc::Aux&, int, int, int, int, int, float,
                       float, float, const float*, const float*, const float*, float*, float*, int*);
  fn_t f = nullptr;
  if (dyn_id == MPPIB_DYN_CARTPOLE && cost_id == MPPIB_COST_CARTPOLE_QUADRATIC)
    f = &orc::sampledTrajectory<orc::Cartpole, orc::CartpoleQuadraticCost>;
  else if (dyn_id == MPPIB_DYN_DOUBLE_INTEGRATOR && cost_id == MPPIB_COST_DI_CIRCLE)
    f = &orc::sampledTrajectory<orc::DoubleIntegrator, orc::DICircleCost>;
  else if (dyn_id == MPPIB_DYN_AUTORALLY_NN && cost_id == MPPIB_COST_AR_STANDARD)
    f = &orc::sampledTrajectory<orc::AutorallyNN, orc::ARStandardCost>;
  else if (dyn_id == MPPIB_DYN_RACER_LSTM && cost_id == MPPIB_COST_RACER_QUADRATIC)
    f = &orc::sampledTrajectory<orc::RacerLSTM, orc::RacerQuadraticCost>;
  else if (dyn_id == MPPIB_DYN_QUADROTOR && cost_id == MPPIB_COST_QUADROTOR_QUADRATIC)
    f = &orc::sampledTrajectory<orc::Quadrotor, orc::QuadrotorQuadraticCost>;
  if (!f)
    return -1;
  f(dyn_params, cost_params, *sp, aux, N, T, d, sample_index, apply_constraints, dt, lambda, alpha, x0, means, controls,
    outputs, costs, crash);
  return 0;
}

int orc_init_eval(int dyn_id, int cost_id, const void* dyn_params, const void* cost_params,
                  const mppib_gaussian_params* sp, const float* nn_theta, const float* costmap, int N_sampler, int T,
                  float dt, float lambda, float alpha, int num_candidates, int num_samples, const float* candidates,
                  const int* strides, const float* means, const float* controls, float* costs)
{
  orc::Aux aux;
  aux.nn_theta = nn_theta;
  aux.costmap = costmap;
  fill_lstm(aux);
  if (dyn_id == MPPIB_DYN_CARTPOLE && cost_id == MPPIB_COST_CARTPOLE_QUADRATIC)
    orc::initEval<orc::Cartpole, orc::CartpoleQuadraticCost>(dyn_params, cost_params, *sp, aux, N_sampler, T, dt, lambda,
                                                             alpha, num_candidates, num_samples, candidates, strides,
                                                             means, controls, costs);
  else if (dyn_id == MPPIB_DYN_DOUBLE_INTEGRATOR && cost_id == MPPIB_COST_DI_CIRCLE)
    orc::initEval<orc::DoubleIntegrator, orc::DICircleCost>(dyn_params, cost_params, *sp, aux, N_sampler, T, dt, lambda,
                                                            alpha, num_candidates, num_samples, candidates, strides,
                                                            means, controls, costs);
  else if (dyn_id == MPPIB_DYN_AUTORALLY_NN && cost_id == MPPIB_COST_AR_STANDARD)
    orc::initEval<orc::AutorallyNN, orc::ARStandardCost>(dyn_params, cost_params, *sp, aux, N_sampler, T, dt, lambda, alpha,
                                                         num_candidates, num_samples, candidates, strides, means,
                                                         controls, costs);
  else if (dyn_id == MPPIB_DYN_QUADROTOR && cost_id == MPPIB_COST_QUADROTOR_QUADRATIC)
    orc::initEval<orc::Quadrotor, orc::QuadrotorQuadraticCost>(dyn_params, cost_params, *sp, aux, N_sampler, T, dt, lambda,
                                                               alpha, num_candidates, num_samples, candidates, strides,
                                                               means, controls, costs);
  else
    return -1;
  return 0;
}

// RobustMPPI host helpers: computeLineSearchWeights (robust_mppi_controller.cu:472-491, out [3][K] row-major),
// computeImportanceSamplerStride (:493-503), computeBestIndex (:519-537; returns -1 if no candidate passes, i.e. the
// reference leaves best_index_ unchanged)
void orc_rmppi_line_search_weights(int num_candidates, float* out)
{
  const int h = num_candidates / 2, K = num_candidates;
  for (int i = 0; i < 3 * K; i++)
    out[i] = 0.0f;
  for (int i = 0; i < h + 1; i++)
  {
    out[0 * K + i] = 1 - i / float(h);
    out[1 * K + i] = i / float(h);
    out[2 * K + i] = 0.0;
  }
  for (int i = 1; i < h + 1; i++)
  {
    out[0 * K + h + i] = 0.0;
    out[1 * K + h + i] = 1 - i / float(h);
    out[2 * K + h + i] = i / float(h);
  }
}
void orc_rmppi_strides(int num_candidates, int stride, int* out)
{
  std::vector<float> w(3 * num_candidates);
  orc_rmppi_line_search_weights(num_candidates, w.data());
  for (int i = 0; i < num_candidates; i++)
    out[i] = (int)roundf(0.0f * w[i] + stride * w[num_candidates + i] + stride * w[2 * num_candidates + i]);
}
int orc_rmppi_best_index(const float* costs, int num_candidates, int samples_per_candidate, float lambda,
                         float value_func_threshold, float* free_energy)
{
  const int n = num_candidates * samples_per_candidate;
  float baseline = costs[0];
  for (int i = 1; i < n; i++)
    if (costs[i] < baseline)
      baseline = costs[i];
  int best = -1;
  for (int i = 0; i < num_candidates; i++)
  {
    float fe = 0.0f;
    for (int j = 0; j < samples_per_candidate; j++)
      fe += expf(-1.0 / lambda * (costs[i * samples_per_candidate + j] - baseline));
    fe /= (1.0 * samples_per_candidate);
    fe = -lambda * logf(fe) + baseline;
    free_energy[i] = fe;
    if (fe < value_func_threshold)
      best = i;
  }
  return best;
}

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

// NLNDistribution::generateSamples' raw noise (sampling_distributions/nln/nln.cu:14-27,114-128) with the HOST XORWOW generator:
// `draws` whole generateSamples calls are made call after call from offset 0 like the reference (C log-normal planes of
// N*T with mean 0 / std dev sigma_c, then N*T*C normals); the last one's product normal[n][t][c] * log_normal[c][n][t] is
// returned in out [N][T][C].
int orc_nln_noise(unsigned long long seed, int draws, int N, int T, int C, const float* std_dev, float* out)
{
  curandGenerator_t g;
  if (curandCreateGeneratorHost(&g, CURAND_RNG_PSEUDO_DEFAULT))
    return -1;
  int rc = 0;
  if (curandSetPseudoRandomGeneratorSeed(g, seed))
    rc = -2;
  if (!rc && curandSetGeneratorOffset(g, 0ULL))
    rc = -3;
  const size_t plane = (size_t)N * T;
  std::vector<float> ln(plane * C);
  for (int k = 0; k < draws && !rc; k++)
  {
    for (int c = 0; c < C && !rc; c++)
      if (curandGenerateLogNormal(g, ln.data() + (size_t)c * plane, plane, 0.0f, std_dev[c]))
        rc = -4;
    if (!rc && curandGenerateNormal(g, out, plane * C, 0.0f, 1.0f))
      rc = -5;
  }
  curandDestroyGenerator(g);
  if (rc)
    return rc;
  for (size_t i = 0; i < plane; i++)
    for (int c = 0; c < C; c++)
      out[i * C + c] = out[i * C + c] * ln[(size_t)c * plane + i];
  return 0;
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
  fill_lstm(aux);
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
void orc_tsallis(float* costs, int n, float gamma, float r, float baseline)
{
  orc::tsallisTransform(costs, n, gamma, r, baseline);
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
    case MPPIB_DYN_QUADROTOR:
    {
      orc::NoCarry k;
      orc::dyn_step<orc::Quadrotor>(*(const mppib_quadrotor_dyn_params*)dyn_params, aux, x, x_next, xdot, u, y, dt, &k);
      return 0;
    }
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
    case MPPIB_COST_RACER_QUADRATIC:
      *cost_out = orc::RacerQuadraticCost::computeStateCost(*(const mppib_racer_quadratic_cost_params*)cost_params,
                                                            aux, y, t, crash);
      *terminal_out = 0;
      return 0;
    case MPPIB_COST_QUADROTOR_QUADRATIC:
      *cost_out = orc::QuadrotorQuadraticCost::computeStateCost(*(const mppib_quadrotor_cost_params*)cost_params, aux, y, t,
                                                                crash);
      *terminal_out = orc::QuadrotorQuadraticCost::terminalCost(*(const mppib_quadrotor_cost_params*)cost_params, aux, y);
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
    case MPPIB_DYN_RACER_LSTM:
      fill_lstm(aux);
      orc::outputTrajectory<orc::RacerLSTM>(dyn_params, aux, x0, u, T, dt, states, outputs);
      return 0;
    case MPPIB_DYN_QUADROTOR:
      orc::outputTrajectory<orc::Quadrotor>(dyn_params, aux, x0, u, T, dt, states, outputs);
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
