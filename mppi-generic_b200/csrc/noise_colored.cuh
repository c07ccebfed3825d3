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

// rearrangeNoise: one thread per (n, t), all C components (reads C rows coalesced along t, writes C contiguous floats).
template <int C>
__global__ void colored_rearrange_kernel(const float* __restrict__ time /*[n][c][2T]*/, float* __restrict__ eps /*[n][t][c]*/,
                                         const float* __restrict__ sigma /*[C]*/, int n_local, int T, int offset_t,
                                         float decay_rate)
{
  const int t = blockIdx.y * blockDim.x + threadIdx.x;
  const int n = blockIdx.x;  // grid.x carries the rollouts (up to 2^31-1), grid.y the time chunks
  if (t >= T || n >= n_local)
    return;
  const float decayed_offset = decay_rate == 0 ? 0 : powf(decay_rate, t);  // colored_noise.cu:45
  float out[C];
#pragma unroll
  for (int c = 0; c < C; c++)
  {
    const float* row = time + ((size_t)n * C + c) * 2 * T;
    out[c] = (row[t] - row[offset_t] * decayed_offset) / (sigma[c] * 2 * T);
  }
  float* dst = eps + ((size_t)n * T + t) * C;
#pragma unroll
  for (int c = 0; c < C; c++)
    dst[c] = out[c];
}

}  // namespace mppib
