/*
 * noise_colored.cuh — K0c, the ColoredNoise sampler's extra passes around the draw (sm_100a).
 *
 * Reference: ColoredNoiseDistributionImpl::generateSamples, sampling_distributions/colored_noise/colored_noise.cu:286-372
 *   curandGenerateNormal(spectrum, 2 * batch * (T+1))           :343
 *   configureFrequencyNoise  (scale by f^(-beta_c/2), zero the imaginary part of DC / Nyquist)   :12-37
 *   cufftExecC2R, plan (2T, C2R, batch = N*C)                   :358, plan :280-282
 *   rearrangeNoise  ([n][c][2T] -> [n][t][c], keep t < T, subtract decay^t * value at t = optimization_stride,
 *                    divide by sigma_c * 2T)                    :39-56
 * followed by the same setGaussianControls as the Gaussian sampler, which K1 applies on the fly.
 *
 * The frequency table and sigma are computed on the host exactly like the reference does every call (:294-338) — here
 * once per parameter change (engine.cu) — and kept on the device. The inverse transform is the same cuFFT plan, so the
 * time series are the reference's bit for bit given the same normals; the two elementwise passes are below. The draw
 * + transform runs one solve ahead on the side stream like the Gaussian draw (the block depends only on the RNG
 * position and on optimization_stride, which is re-applied from the retained time-domain buffer if a solve asks for a
 * different stride than the prefetch assumed).
 */
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace mppib
{
// configureFrequencyNoise over the flat complex array: element p is (row = p / F, f = p % F), c = row % C.
__global__ void colored_scale_kernel(float2* __restrict__ spec, const float* __restrict__ coeffs /*[C][F]*/,
                                     size_t ncomplex, int C, int F)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const bool zero_last = (F % 2) == 1;  // colored_noise.cu:28
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < ncomplex; p += stride)
  {
    const size_t row = p / (size_t)F;
    const int f = (int)(p - row * (size_t)F);
    const int c = (int)(row % (size_t)C);
    const float v = coeffs[c * F + f];
    float2 z = spec[p];
    z.x *= v;
    if (f == 0 || (zero_last && f == F - 1))
      z.y = 0.0f;
    else
      z.y *= v;
    spec[p] = z;
  }
}

// powf(decay_rate, t) for t < T, evaluated once per parameter change by the same device powf the reference calls per
// element (colored_noise.cu:45)
__global__ void colored_decay_table_kernel(float* __restrict__ table, int T, float decay_rate)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T)
    table[t] = decay_rate == 0 ? 0 : powf(decay_rate, t);
}

// rearrangeNoise: one thread per (n, t) pair of the flattened [n][t] index, all C components: reads C rows coalesced
// along t, writes C contiguous floats.
template <int C>
__global__ void colored_rearrange_kernel(const float* __restrict__ time /*[n][c][2T]*/, float* __restrict__ eps /*[n][t][c]*/,
                                         const float* __restrict__ sigma /*[C]*/, const float* __restrict__ decay_pow /*[T]*/,
                                         int n_local, int T, int offset_t)
{
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_local * T)
    return;
  const int n = (int)(idx / (size_t)T);
  const int t = (int)(idx - (size_t)n * T);
  const float decayed_offset = decay_pow[t];
  float out[C];
#pragma unroll
  for (int c = 0; c < C; c++)
  {
    const float* row = time + ((size_t)n * C + c) * 2 * T;
    out[c] = (row[t] - row[offset_t] * decayed_offset) / (sigma[c] * 2 * T);
  }
  float* dst = eps + idx * C;
#pragma unroll
  for (int c = 0; c < C; c++)
    dst[c] = out[c];
}

// ---- NLN sampler (sampling_distributions/nln/nln.cu:14-27, createNLNNoise) --------------------------------------------------
// normal[n][t][c] *= log_normal[c][n][t]; one thread per (n, t), the C factors of a step read from the C planes.
template <int C>
__global__ void nln_combine_kernel(float* __restrict__ normal, const float* __restrict__ log_normal, int n_rollouts, int T)
{
  const size_t plane = (size_t)n_rollouts * T;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x)
  {
#pragma unroll
    for (int c = 0; c < C; c++)
      normal[i * C + c] *= log_normal[c * plane + i];
  }
}

}  // namespace mppib
