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
  }
  return nullptr;
}

// FNNHelper::forward host twin (utils/nn_helpers/fnn_helper.cu:354-382) for the fixed 6-32-32-4 net
void fnn_6_32_32_4(const float* theta, const float* in6, float* out4)
{
  float a1[32], a2[32];
  const float* W1 = theta;
  const float* b1 = theta + 192;
  const float* W2 = theta + 224;
  const float* b2 = theta + 1248;
  const float* W3 = theta + 1280;
  const float* b3 = theta + 1408;
  for (int j = 0; j < 32; j++)
  {
    float s = 0.0f;
    for (int k = 0; k < 6; k++)
      s += W1[j * 6 + k] * in6[k];
    a1[j] = tanhf(s + b1[j]);
  }
  for (int j = 0; j < 32; j++)
  {
    float s = 0.0f;
    for (int k = 0; k < 32; k++)
      s += W2[j * 32 + k] * a1[k];
    a2[j] = tanhf(s + b2[j]);
  }
  for (int j = 0; j < 4; j++)
  {
    float s = 0.0f;
    for (int k = 0; k < 32; k++)
      s += W3[j * 32 + k] * a2[k];
    out4[j] = s + b3[j];
  }
}

int state_deriv(int dyn_id, const void* p, const float* nn_theta, const float* x, const float* u, float* xdot)
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
    case MPPIB_DYN_AUTORALLY_NN:
    {  // dynamics/autorally/ar_nn_model.cu:90-119
      if (!nn_theta)
        return MPPIB_ERR_INVALID_ARG;
      xdot[0] = cosf(x[2]) * x[4] - sinf(x[2]) * x[5];
      xdot[1] = sinf(x[2]) * x[4] + cosf(x[2]) * x[5];
      xdot[2] = -x[6];
      const float in6[6] = { x[3], x[4], x[5], x[6], u[0], u[1] };
      fnn_6_32_32_4(nn_theta, in6, xdot + 3);
      return 0;
    }
  }
  return MPPIB_ERR_UNSUPPORTED;
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

int mppib_host_step(int dyn_id, const void* dyn_params, const float* nn_theta, const float* x, const float* u,
                    float dt, float* x_next, float* xdot, float* y)
{
  int S, C, O;
  if (mppib_host_dims(dyn_id, &S, &C, &O) || !dyn_params || !x || !u || !x_next || !xdot || !y)
    return MPPIB_ERR_INVALID_ARG;
  for (int i = 0; i < S; i++)
    xdot[i] = 0.0f;
  int rc = state_deriv(dyn_id, dyn_params, nn_theta, x, u, xdot);
  if (rc)
    return rc;
  for (int i = 0; i < S; i++)
    x_next[i] = x[i] + xdot[i] * dt;  // dynamics.cuh:277-281
  for (int i = 0; i < O && i < S; i++)
    y[i] = x_next[i];  // dynamics.cuh:292-300
  return MPPIB_OK;
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
  // controller.cuh:643-663 (computeOutputTrajectoryHelper)
  int S, C, O;
  if (mppib_host_dims(dyn_id, &S, &C, &O) || !dyn_params || !x0 || !u || !states || !outputs || T <= 0)
    return MPPIB_ERR_INVALID_ARG;
  std::vector<float> xn(S), xd(S), y(O, 0.0f), ui(C);
  memcpy(states, x0, sizeof(float) * S);
  for (int i = 0; i < O && i < S; i++)
    y[i] = x0[i];  // initializeDynamics (dynamics.cuh:416-423)
  memcpy(outputs, y.data(), sizeof(float) * O);
  for (int t = 0; t < T - 1; t++)
  {
    memcpy(ui.data(), u + (size_t)t * C, sizeof(float) * C);
    enforce(*limits_of(dyn_id, dyn_params), ui.data(), C);
    int rc = mppib_host_step(dyn_id, dyn_params, nn_theta, states + (size_t)t * S, ui.data(), dt, xn.data(), xd.data(),
                             y.data());
    if (rc)
      return rc;
    memcpy(states + (size_t)(t + 1) * S, xn.data(), sizeof(float) * S);
    memcpy(outputs + (size_t)(t + 1) * O, y.data(), sizeof(float) * O);
  }
  return MPPIB_OK;
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
