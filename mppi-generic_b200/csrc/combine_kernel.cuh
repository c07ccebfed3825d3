/*
 * combine_kernel.cuh — K2, the baseline / normaliser / weighted-average merge, for sm_100a.
 *
 * Completes what the reference does with three host round trips (controllers/MPPI/mppi_controller.cu:187-218):
 *   computeBaselineCost  (host min over N,        core/mppi_common.cu:858-900)
 *   normExpKernel        (w = expf(-(c-beta)/lambda), :686-701,958-966)
 *   computeNormalizer    (host double sum,         :1055-1063)   and the sums computeFreeEnergy needs (:1065-1081)
 *   weightedReductionKernel (U_t = sum_n (w_n/eta) u_n,t   :710-737,1115-1160)
 * K1 already produced, per block b, (beta_b, eta_b, sum w^2_b, V_b[t][c] = sum_n w_n u_n) against its LOCAL baseline.
 * With beta = min_b beta_b and s_b = expf(-(beta_b - beta)/lambda):
 *     eta = sum_b s_b eta_b,   sum w^2 = sum_b s_b^2 w2_b,   U = (sum_b s_b V_b) / eta
 * which is algebraically the reference's two-pass formula (w_n = expf(-(c_n-beta_b)/lambda) * s_b). The same kernel
 * merges the per-GPU records after the NCCL all-gather (records = ranks, normalize = 1).
 *
 * Launch: grid (ceil(TC/64), D), block 256 = 4 record-groups x 64 columns; block-wide warp-shuffle reductions.
 * Output record layout == input record layout: [beta, eta, sum_w2, pad, V or U (TC floats)].
 */
#pragma once
#include "device_utils.cuh"

namespace mppib
{
constexpr int kCombineCols = 64;
constexpr int kCombineGroups = 4;

__global__ void __launch_bounds__(kCombineCols* kCombineGroups)
    combine_kernel(const float* __restrict__ records,  // [nrec][D][pstride]
                   int nrec, int D, int TC, int pstride, float lambda_inv, int normalize,
                   float* __restrict__ out,    // [D][pstride] (device or mapped host)
                   float* __restrict__ out2)   // optional second copy (mapped host result), may be nullptr
{
  __shared__ float red_f[8];
  __shared__ float beta_sh;
  __shared__ double eta_sh[kCombineGroups], w2_sh[kCombineGroups];
  __shared__ float acc_sh[kCombineGroups][kCombineCols];

  const int d = blockIdx.y;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int group = tid / kCombineCols, cl = tid % kCombineCols;
  const int col = blockIdx.x * kCombineCols + cl;
  const float* rec = records + (size_t)d * pstride;
  const size_t rstride = (size_t)D * pstride;

  // global baseline: first-minimum value == plain min (mppi_common.cu:858-900)
  float m = INFINITY;
  for (int b = tid; b < nrec; b += blockDim.x)
    m = fminf(m, rec[b * rstride + 0]);
  m = warp_min(m);
  if (lane == 0)
    red_f[warp] = m;
  __syncthreads();
  if (tid == 0)
  {
    float v = red_f[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); i++)
      v = fminf(v, red_f[i]);
    beta_sh = v;
  }
  __syncthreads();
  const float beta = beta_sh;

  float acc = 0.0f;
  double eta = 0.0, w2 = 0.0;
  for (int b = group; b < nrec; b += kCombineGroups)
  {
    const float* r = rec + b * rstride;
    const float s = expf(-lambda_inv * (r[0] - beta));
    eta += (double)s * (double)r[1];
    w2 += (double)s * (double)s * (double)r[2];
    if (col < TC)
      acc = fmaf(s, r[kPartialHeader + col], acc);
  }
  acc_sh[group][cl] = acc;
  if (cl == 0)
  {
    eta_sh[group] = eta;
    w2_sh[group] = w2;
  }
  __syncthreads();
  if (group == 0)
  {
    double e = 0.0, q = 0.0;
    float a = 0.0f;
#pragma unroll
    for (int gq = 0; gq < kCombineGroups; gq++)
    {
      e += eta_sh[gq];
      q += w2_sh[gq];
      a += acc_sh[gq][cl];
    }
    const float eta_f = (float)e;  // mppi_common.cu:1055-1063: double accumulate, narrowed to float
    float* o = out + (size_t)d * pstride;
    float* o2 = out2 ? out2 + (size_t)d * pstride : nullptr;
    if (col < TC)
    {
      const float v = normalize ? a / eta_f : a;
      o[kPartialHeader + col] = v;
      if (o2)
        o2[kPartialHeader + col] = v;
    }
    if (blockIdx.x == 0 && cl == 0)
    {
      o[0] = beta;
      o[1] = eta_f;
      o[2] = (float)q;
      o[3] = 0.0f;
      if (o2)
      {
        o2[0] = beta;
        o2[1] = eta_f;
        o2[2] = (float)q;
        o2[3] = 0.0f;
      }
    }
  }
}

// w_n = expf(-(c_n - beta)/lambda) for read-back (trajectory_costs_d_ after launchNormExpKernel); not on the hot path.
__global__ void weights_kernel(const float* __restrict__ costs, const float* __restrict__ final_rec, int n, int pstride,
                               float lambda_inv, float* __restrict__ w)
{
  const int d = blockIdx.y;
  const float beta = final_rec[(size_t)d * pstride];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    w[(size_t)d * n + i] = expf(-lambda_inv * (costs[(size_t)d * n + i] - beta));
}

}  // namespace mppib
