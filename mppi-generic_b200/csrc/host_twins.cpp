/*
 * host_twins.cpp — host-side (CPU) twins of the plugins and the controller tail, exported from libmppi_b200.so.
 *
 * In the reference every Dynamics / Cost class carries Eigen host methods next to its device methods
 * (include/mppi/dynamics/dynamics.cuh:250-300, include/mppi/cost_functions/cost.cuh:136-219) and the controller
 * finishes computeControl on the host: Savitzky-Golay smoothing, nominal state roll-forward and control clamping
 * (include/mppi/controllers/controller.cuh:557-663, controllers/MPPI/mppi_controller.cu:225-231). Those stay on the
 * host here as well (north_star: "host side stays header-only C++/Eigen"); they are compiled once into the library so
 * that the header-only C++ layer (include/mppi_b200/) and the ctypes mirror (mppi-generic_b200/host.py) share one
 * implementation. These functions are host conveniences of the plugin surface, NOT a fallback for the rollout path:
 * mppib_solve has no CPU route.
 */
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/mppi_b200.h"
#include "../../include/mppi_b200/host_twins.h"

namespace
{
inline float sign_ref(float v)  // utils/math_utils.h:744-747
{
  return v >= 0 ? 1.0f : -1.0f;
}

void enforce(const mppib_control_limits& lim, float* u, int C)  // dynamics.cuh:250-264
{
  for (int i = 0; i < C; i++)
  {
    if (fabsf(u[i]) < lim.deadband[i])
      u[i] = lim.zero_control[i];
    else
      u[i] += lim.deadband[i] * -sign_ref(u[i]);
    u[i] = fminf(fmaxf(lim.rng_lo[i], u[i]), lim.rng_hi[i]);
  }
}

const mppib_control_limits* limits_of(int dyn_id, const void* p)
{
  switch (dyn_id)
  {
    case MPPIB_DYN_CARTPOLE:
      return &((const mppib_cartpole_dyn_params*)p)->lim;
    case MPPIB_DYN_DOUBLE_INTEGRATOR:
      return &((const mppib_di_dyn_params*)p)->lim;
    case MPPIB_DYN_AUTORALLY_NN:
      return &((const mppib_ar_nn_dyn_params*)p)->lim;
    case MPPIB_DYN_RACER_LSTM:
      return &((const mppib_racer_lstm_dyn_params*)p)->lim;
    case MPPIB_DYN_QUADROTOR:
      return &((const mppib_quadrotor_dyn_params*)p)->lim;
  }
  return nullptr;
}

// FNNHelper::forward host twin (utils/nn_helpers/fnn_helper.cu:354-382) for the fixed 6-32-32-4 net.
// The controller calls it T times per computeControl for the nominal trajectory (controller.cuh:643-663), which next to a
// GPU solve of a few hundred microseconds is a visible share of the call (112 us of ~480 us at T = 100 with the plain
// scalar loops + tanhf). So: weights transposed once per trajectory ([in][out], the inner loop runs over the OUT
// neurons and vectorises without reassociating any neuron's k-ascending sum), a branch-free tanh (below), and AVX2+FMA /
// baseline clones selected at load time.
struct FnnT
{
  float WT1[6 * 32], b1[32], WT2[32 * 32], b2[32], W3[4 * 32], b3[4];  // W3 keeps the reference's [out][in] order
};
void fnn_transpose(const float* theta, FnnT& t)
{
  for (int j = 0; j < 32; j++)
    for (int k = 0; k < 6; k++)
      t.WT1[k * 32 + j] = theta[j * 6 + k];
  memcpy(t.b1, theta + 192, sizeof(t.b1));
  for (int j = 0; j < 32; j++)
    for (int k = 0; k < 32; k++)
      t.WT2[k * 32 + j] = theta[224 + j * 32 + k];
  memcpy(t.b2, theta + 1248, sizeof(t.b2));
  memcpy(t.W3, theta + 1280, sizeof(t.W3));
  memcpy(t.b3, theta + 1408, sizeof(t.b3));
}

// tanh(x) = 1 - 2 / (exp(2x) + 1), exp by Cody-Waite reduction + degree-7 polynomial: |abs error| < 2e-7 over the whole
// range (the same bound as the device kernels' tanh_fast; the reference calls tanhf, activation_functions.cuh:15-26).
// Straight-line code so the loops over neurons vectorise.
static inline float exp_host(float z)
{
  z = z > 30.0f ? 30.0f : (z < -30.0f ? -30.0f : z);
  const float nf = (z * 1.44269504088896341f + 12582912.0f) - 12582912.0f;  // round to nearest integer
  float r = z - nf * 0.693145751953125f;                                    // ln2 high part
  r = r - nf * 1.42860682030941723212e-6f;                                  // ln2 low part
  float p = 1.0f / 5040.0f;
  p = p * r + 1.0f / 720.0f;
  p = p * r + 1.0f / 120.0f;
  p = p * r + 1.0f / 24.0f;
  p = p * r + 1.0f / 6.0f;
  p = p * r + 0.5f;
  p = p * r + 1.0f;
  p = p * r + 1.0f;
  int bits = ((int)nf + 127) << 23;
  float scale;
  memcpy(&scale, &bits, sizeof(scale));
  return p * scale;
}
static inline float tanh_host(float x)
{
  return 1.0f - 2.0f / (exp_host(2.0f * x) + 1.0f);
}
static inline float sigmoid_host(float x)  // activation_functions.cuh:49-59 (host branch: 1 / (1 + expf(-x)))
{
  return 1.0f / (1.0f + exp_host(-x));
}

// GCC / clang generic vectors: the same source lowers to 8-lane AVX2 + FMA in the x86-64-v3 clone and to SSE2 pairs in the
// baseline clone. Written out by hand because the auto-vectoriser keeps tanh scalar (clamp branches, float -> int -> float
// round trip): 64 scalar tanh with a vdivss each were 2/3 of the forward pass.
typedef float v8f __attribute__((vector_size(32), aligned(4)));
typedef int v8i __attribute__((vector_size(32), aligned(4)));
static inline v8f splat8(float v)
{
  return v8f{ v, v, v, v, v, v, v, v };
}
static inline v8f load8(const float* p)
{
  v8f v;
  memcpy(&v, p, sizeof(v));
  return v;
}
static inline void store8(float* p, v8f v)
{
  memcpy(p, &v, sizeof(v));
}
// exp_host / tanh_host above, eight lanes at a time (same constants, same operation order)
static inline v8f exp8(v8f z)
{
  const v8f hi = splat8(30.0f), lo = splat8(-30.0f);
  z = z > hi ? hi : (z < lo ? lo : z);
  const v8f magic = splat8(12582912.0f);
  const v8f nf = (z * splat8(1.44269504088896341f) + magic) - magic;
  v8f r = z - nf * splat8(0.693145751953125f);
  r = r - nf * splat8(1.42860682030941723212e-6f);
  v8f p = splat8(1.0f / 5040.0f);
  p = p * r + splat8(1.0f / 720.0f);
  p = p * r + splat8(1.0f / 120.0f);
  p = p * r + splat8(1.0f / 24.0f);
  p = p * r + splat8(1.0f / 6.0f);
  p = p * r + splat8(0.5f);
  p = p * r + splat8(1.0f);
  p = p * r + splat8(1.0f);
  const v8i bits = (__builtin_convertvector(nf, v8i) + 127) << 23;
  v8f scale;
  memcpy(&scale, &bits, sizeof(scale));
  return p * scale;
}
static inline v8f tanh8(v8f x)
{
  return splat8(1.0f) - splat8(2.0f) / (exp8(splat8(2.0f) * x) + splat8(1.0f));
}

__attribute__((target_clones("arch=x86-64-v3", "default"))) void fnn_6_32_32_4(const FnnT& t, const float* in6,
                                                                              float* out4)
{
  // layers 1 and 2: the 32 outputs are four 8-lane vectors, k ascending per neuron (fnn_helper.cu:354-382), bias last
  v8f acc[4], a1[4], a2[4];
  for (int v = 0; v < 4; v++)
    acc[v] = splat8(0.0f);
  for (int k = 0; k < 6; k++)
  {
    const v8f xk = splat8(in6[k]);
    for (int v = 0; v < 4; v++)
      acc[v] += load8(t.WT1 + k * 32 + 8 * v) * xk;
  }
  for (int v = 0; v < 4; v++)
    a1[v] = tanh8(acc[v] + load8(t.b1 + 8 * v));
  float a1s[32];
  for (int v = 0; v < 4; v++)
    store8(a1s + 8 * v, a1[v]);
  // layer 2: even and odd k accumulate separately (two FMA chains of 16 instead of one of 32 on the critical path) and
  // are added at the end; like layer 3 below a reassociation of a few ulp
  v8f acc_odd[4];
  for (int v = 0; v < 4; v++)
    acc[v] = acc_odd[v] = splat8(0.0f);
  for (int k = 0; k < 32; k += 2)
  {
    const v8f xa = splat8(a1s[k]), xb = splat8(a1s[k + 1]);
    for (int v = 0; v < 4; v++)
    {
      acc[v] += load8(t.WT2 + k * 32 + 8 * v) * xa;
      acc_odd[v] += load8(t.WT2 + (k + 1) * 32 + 8 * v) * xb;
    }
  }
  for (int v = 0; v < 4; v++)
    a2[v] = tanh8((acc[v] + acc_odd[v]) + load8(t.b2 + 8 * v));
  // layer 3: four 32-term dot products, eight interleaved partial sums each (one 8-lane FMA chain of length 4) — a
  // reassociation of the kind Eigen's packet products in the reference's host code make as well
  for (int j = 0; j < 4; j++)
  {
    v8f part = load8(t.W3 + j * 32) * a2[0];
    for (int v = 1; v < 4; v++)
      part += load8(t.W3 + j * 32 + 8 * v) * a2[v];
    out4[j] = (((part[0] + part[4]) + (part[1] + part[5])) + ((part[2] + part[6]) + (part[3] + part[7]))) + t.b3[j];
  }
}

static inline int state_deriv(int dyn_id, const void* p, const FnnT* nn, const float* x, const float* u, float* xdot)
{
  switch (dyn_id)
  {
    case MPPIB_DYN_CARTPOLE:
    {  // dynamics/cartpole/cartpole_dynamics.cu:48-69
      const auto& q = *(const mppib_cartpole_dyn_params*)p;
      const float st = sinf(x[2]), ct = cosf(x[2]);
      const float m_c = q.cart_mass, m_p = q.pole_mass, l_p = q.pole_length, g = q.gravity;
      xdot[0] = x[1];
      xdot[1] = 1.0f / (m_c + m_p * st * st) * (u[0] + m_p * st * (l_p * x[3] * x[3] + g * ct));
      xdot[2] = x[3];
      xdot[3] = 1.0f / (l_p * (m_c + m_p * st * st)) *
                (-u[0] * ct - m_p * l_p * x[3] * x[3] * ct * st - (m_c + m_p) * g * st);
      return 0;
    }
    case MPPIB_DYN_DOUBLE_INTEGRATOR:  // dynamics/double_integrator/di_dynamics.cu:14-22
      xdot[0] = x[2];
      xdot[1] = x[3];
      xdot[2] = u[0];
      xdot[3] = u[1];
      return 0;
    case MPPIB_DYN_QUADROTOR:
    {  // dynamics/quadrotor/quadrotor_dynamics.cu:70-112; DCM column 2 as Eigen's toRotationMatrix evaluates it
      const auto& q = *(const mppib_quadrotor_dyn_params*)p;
      const float qw = x[6], qx = x[7], qy = x[8], qz = x[9];
      const float tx = 2.0f * qx, ty = 2.0f * qy, tz = 2.0f * qz;
      const float col2[3] = { tz * qx + ty * qw, tz * qy - tx * qw, 1.0f - (tx * qx + ty * qy) };
      const float tau_inv[3] = { 1 / q.tau_roll, 1 / q.tau_pitch, 1 / q.tau_yaw };
      for (int i = 0; i < 3; i++)
      {
        xdot[i] = x[3 + i];
        xdot[3 + i] = (u[3] / q.mass) * col2[i];
        xdot[10 + i] = tau_inv[i] * (u[i] - x[10 + i]);
      }
      xdot[5] -= MPPIB_GRAVITY;
      const float pp = x[10], qq = x[11], rr = x[12];
      xdot[6] = 0.5f * (-pp * qx - qq * qy - rr * qz);
      xdot[7] = 0.5f * (pp * qw - qq * qz + rr * qy);
      xdot[8] = 0.5f * (pp * qz + qq * qw - rr * qx);
      xdot[9] = 0.5f * (-pp * qy + qq * qx + rr * qw);
      return 0;
    }
    case MPPIB_DYN_AUTORALLY_NN:
    {  // dynamics/autorally/ar_nn_model.cu:90-119
      if (!nn)
        return MPPIB_ERR_INVALID_ARG;
      const float cs = cosf(x[2]), sn = sinf(x[2]);
      xdot[0] = cs * x[4] - sn * x[5];
      xdot[1] = sn * x[4] + cs * x[5];
      xdot[2] = -x[6];
      const float in6[6] = { x[3], x[4], x[5], x[6], u[0], u[1] };
      fnn_6_32_32_4(*nn, in6, xdot + 3);
      return 0;
    }
  }
  return MPPIB_ERR_UNSUPPORTED;
}

// ---- RacerDubinsElevationLSTMSteering host twin ------------------------------------------------------------------
// State / output indices: racer_dubins_elevation.cuh:18-39, racer_dubins.cuh:35-76
enum
{
  R_VEL_X = 0, R_YAW, R_POS_X, R_POS_Y, R_STEER_ANGLE, R_BRAKE_STATE, R_ROLL, R_PITCH, R_STEER_ANGLE_RATE, R_UNC0
};
inline float normalize_angle(float a)  // utils/angle_utils.cuh
{
  const float two_pi = 6.283185307179586f, pi = 3.14159265358979f;
  float r = fmodf(a + pi, two_pi);
  return r <= 0.0f ? r + pi : r - pi;
}

// LSTMHelper::forward (host), utils/nn_helpers/lstm_helper.cu:267-339: gates from W_*m h + W_*i x + b, FNN head on [h; x]
// In-place activations of n values, eight lanes at a time through exp8 (the scalar exp_host per value — ~40 of them per
// step at H = 4, L1 = 20 — was most of the roll-forward): kind 0 = sigmoid 1 / (1 + exp(-x)), kind 1 = tanh.
static inline void activate_n(float* v, int n, int kind)
{
  for (int i = 0; i < n; i += 8)
  {
    float buf[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    const int m = n - i < 8 ? n - i : 8;
    memcpy(buf, v + i, sizeof(float) * m);
    const v8f x = load8(buf);
    const v8f r = kind ? tanh8(x) : splat8(1.0f) / (splat8(1.0f) + exp8(-x));
    store8(buf, r);
    memcpy(v + i, buf, sizeof(float) * m);
  }
}

__attribute__((target_clones("arch=x86-64-v3", "default"))) float lstm_head_forward(const mppib_host_lstm* net,
                                                                                   const float* in4)
{
  const int H = net->hidden_dim, I = MPPIB_RACER_LSTM_INPUT_DIM, L1 = net->head_hidden;
  const int HH = H * H, IH = H * I;
  const float* w = net->theta;
  const float* bias = w + 4 * HH + 4 * IH;
  float g[4][64];  // gate pre-activations [input, forget, output, cell-update][H], H <= 64 (engine limit)
  for (int k = 0; k < 4; k++)
    for (int i = 0; i < H; i++)
    {
      const float* Wm = w + k * HH + i * H;
      const float* Wi = w + 4 * HH + k * IH + i * I;
      float hm = 0.0f, im = 0.0f;
      for (int j = 0; j < H; j++)
        hm += Wm[j] * net->hidden[j];
      for (int j = 0; j < I; j++)
        im += Wi[j] * in4[j];
      g[k][i] = (hm + im) + bias[k * H + i];
    }
  // exp-based activations without libm calls (exp8 above; |abs error| < 2e-7 against expf / tanhf)
  activate_n(g[0], H, 0);
  activate_n(g[1], H, 0);
  activate_n(g[2], H, 0);
  activate_n(g[3], H, 1);
  float hn[64], cn[64];
  for (int i = 0; i < H; i++)
    cn[i] = hn[i] = g[0][i] * g[3][i] + g[1][i] * net->cell[i];
  activate_n(hn, H, 1);
  for (int i = 0; i < H; i++)
    hn[i] = g[2][i] * hn[i];
  memcpy(net->hidden, hn, sizeof(float) * H);
  memcpy(net->cell, cn, sizeof(float) * H);
  const float* hd = w + 4 * HH + 4 * IH + 6 * H;  // head {H+I, L1, 1}: W1 | b1 | W2 | b2 (fnn_helper.cu:176-183)
  const int IN = H + I;
  float a1[64];
  for (int k = 0; k < L1; k++)
  {
    float a = 0.0f;
    for (int j = 0; j < H; j++)
      a += hd[k * IN + j] * hn[j];
    for (int j = 0; j < I; j++)
      a += hd[k * IN + H + j] * in4[j];
    a1[k] = a + hd[L1 * IN + k];
  }
  activate_n(a1, L1, 1);
  float out = 0.0f;
  for (int k = 0; k < L1; k++)
    out += hd[L1 * IN + L1 + k] * a1[k];
  return out + hd[L1 * IN + 2 * L1];
}

// TextureHelper::worldPoseToTexCoord (texture_helper.cu:94-134) + TwoDTextureHelper::queryTextureCPU
// (two_d_texture_helper.cu:151-243) for the TextureParams defaults: clamp addressing, bilinear filter
float elevation_at_world_pose(const mppib_elevation_map_header* h, float wx, float wy, float wz)
{
  const float* data = reinterpret_cast<const float*>(h + 1);
  const float dx = wx - h->origin[0], dy = wy - h->origin[1], dz = wz - h->origin[2];
  const float mx = h->rotations[0] * dx + h->rotations[1] * dy + h->rotations[2] * dz;
  const float my = h->rotations[3] * dx + h->rotations[4] * dy + h->rotations[5] * dz;
  float qx = ((mx / h->resolution[0]) / (float)h->width) * (float)h->width - 0.5f;
  float qy = ((my / h->resolution[1]) / (float)h->height) * (float)h->height - 0.5f;
  if (qx > (float)(h->width - 1))
    qx = (float)(h->width - 1);
  else if (qx <= 0.0f)
    qx = 0.0f;
  if (qy > (float)(h->height - 1))
    qy = (float)(h->height - 1);
  else if (qy <= 0.0f)
    qy = 0.0f;
  if (std::isnan(qx) || std::isnan(qy))
    return NAN;
  const int x0 = std::min((int)std::floor(qx), h->width - 2), y0 = std::min((int)std::floor(qy), h->height - 2);
  const int w = h->width;
  const float q11 = data[(size_t)y0 * w + x0], q12 = data[(size_t)y0 * w + x0 + 1];
  const float q21 = data[(size_t)(y0 + 1) * w + x0], q22 = data[(size_t)(y0 + 1) * w + x0 + 1];
  const float lo = q11 * ((float)(x0 + 1) - qx) + q12 * (qx - (float)x0);
  const float hi = q21 * ((float)(x0 + 1) - qx) + q22 * (qx - (float)x0);
  return lo * ((float)(y0 + 1) - qy) + hi * (qy - (float)y0);
}
// RACER::computeStaticSettling, racer_dubins.cu:359-434 (host branch of math::Euler2DCM_NWU: sincosf without normalisation)
float static_settling(const mppib_elevation_map_header* map, float yaw, float x, float y, float& roll, float& pitch)
{
  if (!map || !map->use)
  {
    roll = 0.0f;
    pitch = 0.0f;
    return 0.0f;
  }
  float sr, cr, sp, cp, sy, cy;
  sincosf(roll, &sr, &cr);
  sincosf(pitch, &sp, &cp);
  sincosf(yaw, &sy, &cy);
  const float M00 = cp * cy, M01 = sr * sp * cy - cr * sy, M10 = cp * sy, M11 = sr * sp * sy + cr * cy, M20 = -sp,
              M21 = sr * cp;
  float hgt[4];
  for (int k = 0; k < 4; k++)
  {  // front left, front right, rear left, rear right (:364-367)
    const float ox = (k < 2) ? 2.981f : 0.0f, oy = (k & 1) ? -0.737f : 0.737f;
    hgt[k] = elevation_at_world_pose(map, M00 * ox + M01 * oy + x, M10 * ox + M11 * oy + y, M20 * ox + M21 * oy + 0.0f);
  }
  const float fl = hgt[0], fr = hgt[1], rl = hgt[2], rr = hgt[3];
  const float front_diff = fmaxf(fminf(fl - fr, 0.736f * 2.0f), -0.736f * 2.0f);
  const float rear_diff = fmaxf(fminf(rl - rr, 0.736f * 2.0f), -0.736f * 2.0f);
  roll = (asinf(front_diff / (0.737f * 2.0f)) + asinf(rear_diff / (0.737f * 2.0f))) / 2.0f;
  const float left_diff = fmaxf(fminf(rl - fl, 2.98f), -2.98f);
  const float right_diff = fmaxf(fminf(rr - fr, 2.98f), -2.98f);
  pitch = (asinf(left_diff / 2.981f) + asinf(right_diff / 2.981f)) / 2.0f;
  float height = (rl + rr) / 2.0f;
  if (!std::isfinite(roll) || fabsf(roll) > (float)M_PI)
    roll = 2.0f * (float)M_PI;
  if (!std::isfinite(pitch) || fabsf(pitch) > (float)M_PI)
    pitch = 2.0f * (float)M_PI;
  if (!std::isfinite(height))
    height = 0.0f;
  return height;
}

// racer_dubins_elevation_lstm_steering.cu:90-118 (host step) and the host methods it calls
void racer_step(const mppib_racer_lstm_dyn_params& p, const mppib_host_lstm* net, const float* x, const float* u,
                float dt, float* xn, float* xd, float* y)
{
  const float vx = x[R_VEL_X];
  const int index = (fabsf(vx) > 0.2f && fabsf(vx) <= 3.0f) + (fabsf(vx) > 3.0f) * 2;
  const bool enable_brake = u[0] < 0.0f;
  const float brake_error = (enable_brake * -u[0] - x[R_BRAKE_STATE]);  // racer_dubins.cu:306-319
  xd[R_BRAKE_STATE] = fminf(fmaxf((brake_error > 0) * brake_error * p.brake_delay_constant +
                                      (brake_error < 0) * brake_error * p.brake_delay_constant_neg,
                                  -p.max_brake_rate_neg),
                            p.max_brake_rate_pos);
  const float brake_state = fminf(fmaxf(x[R_BRAKE_STATE], 0.0f), 0.25f);  // racer_dubins_elevation.cu:32-67
  float throttle = p.c_t[index] * u[0];
  float brake = p.c_b[index] * brake_state * (vx >= 0.0f ? -1.0f : 1.0f);
  if (fabsf(vx) <= 0.2f)
  {
    throttle = p.c_t[index] * fmaxf(u[0] - p.low_min_throttle, 0.0f);
    brake = p.c_b[index] * brake_state * -vx;
  }
  xd[R_VEL_X] = (!enable_brake) * throttle * p.gear_sign + brake - p.c_v[index] * vx + p.c_0;
  xd[R_VEL_X] = fminf(fmaxf(xd[R_VEL_X], -p.clamp_ax), p.clamp_ax);
  if (fabsf(x[R_PITCH]) < 1.57079632679489661923f)
    xd[R_VEL_X] -= p.gravity * sinf(x[R_PITCH]);
  const float delta = x[R_STEER_ANGLE] / p.steer_angle_scale;
  const float tan_delta = tanf(delta);
  xd[R_YAW] = (vx / p.wheel_base) * tan_delta;
  const float sy = sinf(x[R_YAW]), cy = cosf(x[R_YAW]);
  xd[R_POS_X] = vx * cy;
  xd[R_POS_Y] = vx * sy;
  {  // computeLSTMSteering, lstm_steering.cu:66-88
    const float parametric_accel = (u[1] * p.steer_command_angle_scale - x[R_STEER_ANGLE]) * p.steering_constant;
    xd[R_STEER_ANGLE_RATE] = fmaxf(fminf((parametric_accel - x[R_STEER_ANGLE_RATE]) * p.steer_accel_constant -
                                             x[R_STEER_ANGLE_RATE] * p.steer_accel_drag_constant,
                                         p.max_steer_rate),
                                   -p.max_steer_rate);
    const float in4[4] = { x[R_STEER_ANGLE] * 0.2f, x[R_STEER_ANGLE_RATE] * 0.2f, u[1], xd[R_STEER_ANGLE_RATE] * 0.2f };
    xd[R_STEER_ANGLE_RATE] += lstm_head_forward(net, in4) * 5.0f;
    xd[R_STEER_ANGLE] = x[R_STEER_ANGLE_RATE];
  }
  for (int i = 0; i < 6; i++)  // updateState, lstm_steering.cu:267-285
    xn[i] = x[i] + xd[i] * dt;
  xn[R_YAW] = normalize_angle(xn[R_YAW]);
  xn[R_STEER_ANGLE] = fmaxf(fminf(xn[R_STEER_ANGLE], p.max_steer_angle), -p.max_steer_angle);
  xn[R_STEER_ANGLE_RATE] = x[R_STEER_ANGLE_RATE] + xd[R_STEER_ANGLE_RATE] * dt;
  xn[R_BRAKE_STATE] = fminf(fmaxf(xn[R_BRAKE_STATE], 0.0f), -p.lim.rng_lo[0]);
  {  // computeUncertaintyPropagation, racer_dubins_elevation.cu:662-741; matrices column-major 4x4 over
     // (VEL_X, YAW, POS_X, POS_Y); state order of the 10 covariance entries: racer_dubins_elevation.cuh:29-38
    float A[4][4] = {}, Sg[4][4], Tm[4][4], Q[4][4] = {};  // [row][col]
    const float c2 = cosf(delta) * cosf(delta);
    A[0][0] = -p.c_v[index] - p.K_vel_x - (index == 0 ? 1.0f : 0.0f) * p.c_b[0] * brake_state;
    A[0][2] = -p.K_x * cy;
    A[0][3] = -p.K_x * sy;
    A[1][0] = tan_delta / p.wheel_base;
    A[1][1] = -fabsf(vx) * p.K_yaw / (p.wheel_base * c2);
    A[1][2] = vx * p.K_y * sy / (p.wheel_base * c2);
    A[1][3] = -vx * p.K_y * cy / (p.wheel_base * c2);
    A[2][0] = cy;
    A[2][1] = -sy * vx;
    A[3][0] = sy;
    A[3][1] = cy * vx;
    const float* s = x + R_UNC0;  // POS_X, POS_Y, YAW, VEL_X, POS_X_Y, POS_X_YAW, POS_X_VEL_X, POS_Y_YAW, POS_Y_VEL_X, YAW_VEL_X
    Sg[0][0] = s[3], Sg[1][1] = s[2], Sg[2][2] = s[0], Sg[3][3] = s[1];
    Sg[1][0] = Sg[0][1] = s[9];
    Sg[2][0] = Sg[0][2] = s[6];
    Sg[3][0] = Sg[0][3] = s[8];
    Sg[2][1] = Sg[1][2] = s[5];
    Sg[3][1] = Sg[1][3] = s[7];
    Sg[3][2] = Sg[2][3] = s[4];
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++)
        A[r][c] = (r == c) + A[r][c] * dt;
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++)
      {
        float acc = 0.0f;
        for (int k = 0; k < 4; k++)
          acc += A[r][k] * Sg[k][c];
        Tm[r][c] = acc;
      }
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++)
      {
        float acc = 0.0f;
        for (int k = 0; k < 4; k++)
          acc += Tm[r][k] * A[c][k];
        Sg[r][c] = acc;
      }
    const float abs_vx = fabsf(vx);
    const float side_force = abs_vx * abs_vx * tan_delta / p.wheel_base + p.gravity * sinf(x[R_ROLL]);
    const float Q_11 = fabsf(p.Q_y_f * fabsf(side_force) * fmaxf(abs_vx - 2, 0.0f));
    Q[0][0] = p.Q_x_acc * fabsf(xd[R_VEL_X]) + p.Q_x_v[index] * abs_vx;
    Q[1][1] = abs_vx * (p.Q_omega_steering * fabsf(delta) + p.Q_omega_v);
    Q[2][2] = Q_11 * sy * sy;
    Q[2][3] = Q[3][2] = -Q_11 * sy * cy;
    Q[3][3] = Q_11 * cy * cy;
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++)
        Sg[r][c] += Q[r][c] * dt;
    float* o = xn + R_UNC0;
    o[0] = Sg[2][2], o[1] = Sg[3][3], o[2] = Sg[1][1], o[3] = Sg[0][0], o[4] = Sg[3][2], o[5] = Sg[2][1], o[6] = Sg[2][0],
    o[7] = Sg[3][1], o[8] = Sg[3][0], o[9] = Sg[1][0];
  }
  // computeStaticSettling (lstm_steering.cu:105-112): current roll / pitch, next yaw and position; flat without a map
  float roll = x[R_ROLL], pitch = x[R_PITCH];
  const float height = static_settling(net->map, xn[R_YAW], xn[R_POS_X], xn[R_POS_Y], roll, pitch);
  xn[R_ROLL] = roll;
  xn[R_PITCH] = pitch;
  // setOutputs, racer_dubins_elevation.cu:69-227 (output order racer_dubins.cuh:35-76)
  y[0] = xn[R_VEL_X], y[1] = 0.0f, y[2] = xn[R_POS_X], y[3] = xn[R_POS_Y], y[4] = height, y[5] = xn[R_YAW];
  y[6] = xn[R_ROLL], y[7] = xn[R_PITCH], y[8] = xn[R_STEER_ANGLE], y[9] = xn[R_STEER_ANGLE_RATE];
  y[10] = y[11] = y[12] = NAN;
  y[13] = xd[R_VEL_X], y[14] = 0.0f, y[15] = xd[R_YAW], y[16] = fabsf(xn[R_VEL_X]);
  for (int i = 0; i < 10; i++)
    y[17 + i] = xn[R_UNC0 + i];
  y[27] = 0.0f;
}
// controller.cuh:643-663 (computeOutputTrajectoryHelper) with the model's dimensions and its derivative known at compile
// time: the T-step loop is part of every computeControl, so the per-step switch / allocation overhead of the generic
// entry points is kept out of it
// Dynamics::updateState (dynamics.cuh:277-281) and the one override in tree: QuadrotorDynamics renormalises the
// quaternion after the Euler step (quadrotor_dynamics.cu:114-122)
static inline void update_state(int dyn_id, int S, const float* x, const float* xd, float dt, float* xn)
{
  for (int i = 0; i < S; i++)
    xn[i] = x[i] + xd[i] * dt;
  if (dyn_id == MPPIB_DYN_QUADROTOR)
  {
    float* q = xn + 6;
    const float norm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const float div = (float)((double)norm * copysign(1.0, (double)q[0]));
    for (int i = 0; i < 4; i++)
      q[i] /= div;
  }
}

template <int DYN_ID, int S, int C, int O>
static int output_trajectory_impl(const void* dyn_params, const FnnT* nn, const mppib_control_limits& lim, const float* x0,
                                  const float* u, int T, float dt, float* states, float* outputs)
{
  float xn[S], xd[S], y[O], ui[C];
  memcpy(states, x0, sizeof(float) * S);
  for (int i = 0; i < O; i++)
    y[i] = (i < S) ? x0[i] : 0.0f;  // initializeDynamics (dynamics.cuh:416-423)
  memcpy(outputs, y, sizeof(float) * O);
  for (int t = 0; t < T - 1; t++)
  {
    const float* x = states + (size_t)t * S;
    for (int i = 0; i < C; i++)
      ui[i] = u[(size_t)t * C + i];
    enforce(lim, ui, C);
    for (int i = 0; i < S; i++)
      xd[i] = 0.0f;
    const int rc = state_deriv(DYN_ID, dyn_params, nn, x, ui, xd);  // constant id: the switch folds away
    if (rc)
      return rc;
    update_state(DYN_ID, S, x, xd, dt, xn);
    for (int i = 0; i < O && i < S; i++)
      y[i] = xn[i];  // dynamics.cuh:292-300
    memcpy(states + (size_t)(t + 1) * S, xn, sizeof(float) * S);
    memcpy(outputs + (size_t)(t + 1) * O, y, sizeof(float) * O);
  }
  return MPPIB_OK;
}

}  // namespace

extern "C" {

int mppib_host_dims(int dyn_id, int* S, int* C, int* O)
{
  int s = 0, c = 0, o = 0;
  switch (dyn_id)
  {
    case MPPIB_DYN_CARTPOLE:
      s = 4, c = 1, o = 4;
      break;
    case MPPIB_DYN_DOUBLE_INTEGRATOR:
      s = 4, c = 2, o = 4;
      break;
    case MPPIB_DYN_AUTORALLY_NN:
      s = 7, c = 2, o = 8;
      break;
    case MPPIB_DYN_RACER_LSTM:
      s = 19, c = 2, o = 28;
      break;
    case MPPIB_DYN_QUADROTOR:
      s = 13, c = 4, o = 13;
      break;
    default:
      return MPPIB_ERR_UNSUPPORTED;
  }
  if (S)
    *S = s;
  if (C)
    *C = c;
  if (O)
    *O = o;
  return MPPIB_OK;
}

int mppib_host_enforce_constraints(int dyn_id, const void* dyn_params, float* u)
{
  int S, C, O;
  if (mppib_host_dims(dyn_id, &S, &C, &O) || !dyn_params || !u)
    return MPPIB_ERR_INVALID_ARG;
  enforce(*limits_of(dyn_id, dyn_params), u, C);
  return MPPIB_OK;
}

static int host_step_impl(int dyn_id, const void* dyn_params, const FnnT* nn, const float* x, const float* u, float dt,
                          float* x_next, float* xdot, float* y)
{
  int S, C, O;
  if (mppib_host_dims(dyn_id, &S, &C, &O) || !dyn_params || !x || !u || !x_next || !xdot || !y)
    return MPPIB_ERR_INVALID_ARG;
  for (int i = 0; i < S; i++)
    xdot[i] = 0.0f;
  int rc = state_deriv(dyn_id, dyn_params, nn, x, u, xdot);
  if (rc)
    return rc;
  update_state(dyn_id, S, x, xdot, dt, x_next);
  for (int i = 0; i < O && i < S; i++)
    y[i] = x_next[i];  // dynamics.cuh:292-300
  return MPPIB_OK;
}

int mppib_host_step(int dyn_id, const void* dyn_params, const float* nn_theta, const float* x, const float* u,
                    float dt, float* x_next, float* xdot, float* y)
{
  FnnT nn;
  if (dyn_id == MPPIB_DYN_AUTORALLY_NN && nn_theta)
    fnn_transpose(nn_theta, nn);
  return host_step_impl(dyn_id, dyn_params, (dyn_id == MPPIB_DYN_AUTORALLY_NN && nn_theta) ? &nn : nullptr, x, u, dt,
                        x_next, xdot, y);
}

void mppib_host_smooth_controls(float* u, const float* history, int T, int C)
{
  // controller.cuh:557-586: coefficients (-3 12 17 12 -3)/35 over [history(2) | u(T) | u_last u_last]
  const float coef[5] = { -3.0f / 35.0f, 12.0f / 35.0f, 17.0f / 35.0f, 12.0f / 35.0f, -3.0f / 35.0f };
  std::vector<float> buf((size_t)(T + 4) * C);
  for (int c = 0; c < C; c++)
  {
    buf[c] = history[c];
    buf[C + c] = history[C + c];
    for (int t = 0; t < T; t++)
      buf[(size_t)(t + 2) * C + c] = u[(size_t)t * C + c];
    buf[(size_t)(T + 2) * C + c] = u[(size_t)(T - 1) * C + c];
    buf[(size_t)(T + 3) * C + c] = u[(size_t)(T - 1) * C + c];
  }
  for (int t = 0; t < T; t++)
    for (int c = 0; c < C; c++)
    {
      float acc = 0.0f;
      for (int k = 0; k < 5; k++)
        acc += coef[k] * buf[(size_t)(t + k) * C + c];
      u[(size_t)t * C + c] = acc;
    }
}

void mppib_host_slide_controls(float* u, int steps, int T, int C, const float* zero_control, const float* scale)
{
  // controller.cuh:588-600
  for (int i = 0; i < T; ++i)
  {
    const int ind = std::min(i + steps, T - 1);
    for (int c = 0; c < C; c++)
    {
      float v = u[(size_t)ind * C + c];
      if (i + steps > T - 1)
        v = (v - zero_control[c]) * scale[c] + zero_control[c];
      u[(size_t)i * C + c] = v;
    }
  }
}

int mppib_host_output_trajectory(int dyn_id, const void* dyn_params, const float* nn_theta, const float* x0,
                                 const float* u, int T, float dt, float* states, float* outputs)
{
  if (!dyn_params || !x0 || !u || !states || !outputs || T <= 0)
    return MPPIB_ERR_INVALID_ARG;
  switch (dyn_id)
  {
    case MPPIB_DYN_CARTPOLE:
      return output_trajectory_impl<MPPIB_DYN_CARTPOLE, 4, 1, 4>(dyn_params, nullptr, *limits_of(dyn_id, dyn_params), x0, u,
                                                                T, dt, states, outputs);
    case MPPIB_DYN_DOUBLE_INTEGRATOR:
      return output_trajectory_impl<MPPIB_DYN_DOUBLE_INTEGRATOR, 4, 2, 4>(dyn_params, nullptr,
                                                                         *limits_of(dyn_id, dyn_params), x0, u, T, dt,
                                                                         states, outputs);
    case MPPIB_DYN_AUTORALLY_NN:
    {
      if (!nn_theta)
        return MPPIB_ERR_INVALID_ARG;
      FnnT nn;
      fnn_transpose(nn_theta, nn);
      return output_trajectory_impl<MPPIB_DYN_AUTORALLY_NN, 7, 2, 8>(dyn_params, &nn, *limits_of(dyn_id, dyn_params), x0,
                                                                    u, T, dt, states, outputs);
    }
    case MPPIB_DYN_QUADROTOR:
      return output_trajectory_impl<MPPIB_DYN_QUADROTOR, 13, 4, 13>(dyn_params, nullptr, *limits_of(dyn_id, dyn_params), x0,
                                                                   u, T, dt, states, outputs);
  }
  return MPPIB_ERR_UNSUPPORTED;
}

// LSTMLSTMHelper::initializeLSTM, lstm_lstm_helper.cu:50-73, with LSTMHelper::forward (host, lstm_helper.cu:267-339) and
// FNNHelper::forward (host: tanh between layers, linear output) for arbitrary dimensions. Runs once per re-initialisation
// (a few hundred microseconds of scalar code), not per solve.
int mppib_host_lstm_initialize(const mppib_host_init_lstm* net, const float* buffer, int cols, float* out)
{
  if (!net || !net->lstm_theta || !net->head_theta || !net->head_layers || !buffer || !out)
    return MPPIB_ERR_INVALID_ARG;
  const int I = net->input_dim, H = net->hidden_dim, L = net->head_num_layers;
  if (I <= 0 || H <= 0 || L < 2 || net->init_len <= 0 || cols < net->init_len || net->head_layers[0] != H + I)
    return MPPIB_ERR_INVALID_ARG;
  const int HH = H * H, IH = H * I;
  const float* w = net->lstm_theta;
  const float* bias = w + 4 * HH + 4 * IH;
  std::vector<float> h(bias + 4 * H, bias + 5 * H), c(bias + 5 * H, bias + 6 * H);  // resetHiddenCellCPU
  std::vector<float> hn(H), cn(H);
  for (int t = cols - net->init_len; t < cols; t++)
  {
    const float* x = buffer + (size_t)t * I;
    for (int i = 0; i < H; i++)
    {
      float g[4];
      for (int k = 0; k < 4; k++)
      {
        const float* Wm = w + k * HH + i * H;
        const float* Wi = w + 4 * HH + k * IH + i * I;
        float hm = 0.0f, im = 0.0f;
        for (int j = 0; j < H; j++)
          hm += Wm[j] * h[j];
        for (int j = 0; j < I; j++)
          im += Wi[j] * x[j];
        g[k] = (hm + im) + bias[k * H + i];
      }
      const float gi = 1.0f / (1.0f + expf(-g[0])), gf = 1.0f / (1.0f + expf(-g[1])), go = 1.0f / (1.0f + expf(-g[2]));
      cn[i] = gi * tanhf(g[3]) + gf * c[i];
      hn[i] = go * tanhf(cn[i]);
    }
    h = hn;
    c = cn;
  }
  // head on [h; x_last]
  int widest = 0;
  for (int l = 0; l < L; l++)
  {
    if (net->head_layers[l] <= 0)
      return MPPIB_ERR_INVALID_ARG;
    widest = std::max(widest, net->head_layers[l]);
  }
  std::vector<float> a(widest), b(widest);
  for (int i = 0; i < H; i++)
    a[i] = h[i];
  for (int i = 0; i < I; i++)
    a[H + i] = buffer[(size_t)(cols - 1) * I + i];
  const float* th = net->head_theta;
  for (int l = 0; l + 1 < L; l++)
  {
    const int in = net->head_layers[l], on = net->head_layers[l + 1];
    const float* W = th;
    const float* bb = th + (size_t)in * on;
    for (int o = 0; o < on; o++)
    {
      float acc = 0.0f;
      for (int j = 0; j < in; j++)
        acc += W[(size_t)o * in + j] * a[j];
      acc += bb[o];
      b[o] = (l + 2 < L) ? tanhf(acc) : acc;
    }
    std::swap(a, b);
    th += (size_t)in * on + on;
  }
  memcpy(out, a.data(), sizeof(float) * net->head_layers[L - 1]);
  return MPPIB_OK;
}

float mppib_host_elevation_at_world_pose(const mppib_elevation_map_header* map, float x, float y, float z)
{
  return map ? elevation_at_world_pose(map, x, y, z) : 0.0f;
}
float mppib_host_static_settling(const mppib_elevation_map_header* map, float yaw, float x, float y, float* roll, float* pitch)
{
  float r = roll ? *roll : 0.0f, p = pitch ? *pitch : 0.0f;
  const float h = static_settling(map, yaw, x, y, r, p);
  if (roll)
    *roll = r;
  if (pitch)
    *pitch = p;
  return h;
}

int mppib_host_step_lstm(const void* dyn_params, const mppib_host_lstm* net, const float* x, const float* u, float dt,
                         float* x_next, float* xdot, float* y)
{
  if (!dyn_params || !net || !net->theta || !net->hidden || !net->cell || !x || !u || !x_next || !xdot || !y)
    return MPPIB_ERR_INVALID_ARG;
  for (int i = 0; i < 19; i++)
    xdot[i] = 0.0f;
  racer_step(*(const mppib_racer_lstm_dyn_params*)dyn_params, net, x, u, dt, x_next, xdot, y);
  return MPPIB_OK;
}

int mppib_host_output_trajectory_lstm(const void* dyn_params, const mppib_host_lstm* net, const float* x0,
                                      const float* u, int T, float dt, float* states, float* outputs)
{
  // controller.cuh:643-663; initializeDynamics resets the LSTM to its initial hidden / cell state
  // (lstm_steering.cu:230-237), which the weight blob carries after the biases (lstm_helper.cu:86-87)
  if (!dyn_params || !net || !net->theta || !x0 || !u || !states || !outputs || T <= 0)
    return MPPIB_ERR_INVALID_ARG;
  const int S = 19, C = 2, O = 28, H = net->hidden_dim;
  std::vector<float> h(H), c(H), xn(S), xd(S), y(O, 0.0f);
  const float* init = net->theta + 4 * H * H + 4 * H * MPPIB_RACER_LSTM_INPUT_DIM + 4 * H;
  memcpy(h.data(), init, sizeof(float) * H);
  memcpy(c.data(), init + H, sizeof(float) * H);
  mppib_host_lstm local = *net;
  local.hidden = h.data();
  local.cell = c.data();
  memcpy(states, x0, sizeof(float) * S);
  for (int i = 0; i < O && i < S; i++)
    y[i] = x0[i];
  memcpy(outputs, y.data(), sizeof(float) * O);
  const auto& p = *(const mppib_racer_lstm_dyn_params*)dyn_params;
  for (int t = 0; t < T - 1; t++)
  {
    float ui[2] = { u[(size_t)t * C], u[(size_t)t * C + 1] };
    enforce(p.lim, ui, C);
    for (int i = 0; i < S; i++)
      xd[i] = 0.0f;
    racer_step(p, &local, states + (size_t)t * S, ui, dt, xn.data(), xd.data(), y.data());
    memcpy(states + (size_t)(t + 1) * S, xn.data(), sizeof(float) * S);
    memcpy(outputs + (size_t)(t + 1) * O, y.data(), sizeof(float) * O);
  }
  return MPPIB_OK;
}

// ---- RobustMPPI host logic (controllers/R-MPPI/robust_mppi_controller.cu) --------------------------------------------
void mppib_host_rmppi_line_search_weights(int num_candidates, float* out /*[3][K]*/)
{  // computeLineSearchWeights, :472-491
  const int K = num_candidates, h = K / 2;
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

void mppib_host_rmppi_candidates(int num_candidates, int S, const float* nominal_x_k, const float* nominal_x_kp1,
                                 const float* real_x_kp1, int stride, float* candidates /*[K][S]*/, int* strides /*[K]*/)
{  // getInitNominalStateCandidates :351-362 (points * line_search_weights) and computeImportanceSamplerStride :493-503
  std::vector<float> w(3 * (size_t)num_candidates);
  mppib_host_rmppi_line_search_weights(num_candidates, w.data());
  const int K = num_candidates;
  for (int k = 0; k < K; k++)
  {
    for (int i = 0; i < S; i++)
      candidates[(size_t)k * S + i] = nominal_x_k[i] * w[k] + nominal_x_kp1[i] * w[K + k] + real_x_kp1[i] * w[2 * K + k];
    strides[k] = (int)roundf(0.0f * w[k] + (float)stride * w[K + k] + (float)stride * w[2 * K + k]);
  }
}

int mppib_host_rmppi_best_index(const float* costs, int num_candidates, int samples_per_candidate, float lambda,
                                float value_func_threshold, int previous_best, float* free_energy /*[K] or NULL*/)
{  // computeCandidateBaseline + computeBestIndex, :505-537 (the LAST candidate under the threshold wins; if none does
   // best_index_ keeps its previous value)
  const int n = num_candidates * samples_per_candidate;
  float baseline = costs[0];
  for (int i = 1; i < n; i++)
    if (costs[i] < baseline)
      baseline = costs[i];
  int best = previous_best;
  for (int i = 0; i < num_candidates; i++)
  {
    float fe = 0.0f;
    for (int j = 0; j < samples_per_candidate; j++)
      fe += expf(-1.0 / lambda * (costs[i * samples_per_candidate + j] - baseline));
    fe /= (1.0 * samples_per_candidate);
    fe = -lambda * logf(fe) + baseline;
    if (free_energy)
      free_energy[i] = fe;
    if (fe < value_func_threshold)
      best = i;
  }
  return best;
}

void mppib_host_free_energy(const mppib_solve_stats* st, int num_rollouts, float lambda, float* out3)
{
  // core/mppi_common.cu:1065-1081 from (eta, sum w^2): norm = eta/N, var = sum w^2
  const float norm = st->normalizer / num_rollouts;
  out3[0] = -lambda * logf(norm) + st->baseline;
  out3[1] = lambda * (st->sum_w2 / num_rollouts - norm * norm);
  const float weird_term = out3[1] / (norm * sqrtf(1.0f * num_rollouts));
  out3[2] = lambda * (weird_term + 0.5f * weird_term * weird_term);
}


int mppib_host_merge_records(const float* records, int nrec, int D, int TC, int pstride, float lambda, int normalize,
                             float* out)
{
  // CPU twin of combine_kernel (csrc/combine_kernel.cuh): merges per-block or per-rank records
  // [beta_b, eta_b, sum w^2_b, -, V_b[TC]] with s_b = expf(-(beta_b - beta)/lambda). Used by launchers and tests that
  // reason about rollout sharding without a GPU; the engine itself always merges on the device.
  if (!records || !out || nrec <= 0 || D <= 0 || TC <= 0 || pstride < 4 + TC || !(lambda > 0.0f))
    return MPPIB_ERR_INVALID_ARG;
  const float lambda_inv = (float)(1.0 / lambda);
  for (int d = 0; d < D; d++)
  {
    const float* rec = records + (size_t)d * pstride;
    const size_t rstride = (size_t)D * pstride;
    float beta = rec[0];
    for (int b = 1; b < nrec; b++)
      beta = fminf(beta, rec[b * rstride]);
    double eta = 0.0, w2 = 0.0;
    std::vector<float> acc(TC, 0.0f);
    for (int b = 0; b < nrec; b++)
    {
      const float* r = rec + b * rstride;
      const float s = expf(-lambda_inv * (r[0] - beta));
      eta += (double)s * (double)r[1];
      w2 += (double)s * (double)s * (double)r[2];
      for (int c = 0; c < TC; c++)
        acc[c] = fmaf(s, r[4 + c], acc[c]);
    }
    float* o = out + (size_t)d * pstride;
    o[0] = beta;
    o[1] = (float)eta;
    o[2] = (float)w2;
    o[3] = 0.0f;
    for (int c = 0; c < TC; c++)
      o[4 + c] = normalize ? acc[c] / (float)eta : acc[c];
  }
  return MPPIB_OK;
}

}  // extern "C"
