// mma.sync (FP16 hi/lo split, 3 products) evaluation of the Autorally 6-32-32-4 network, one warp = 32 samples (GPU box):
//   1. accuracy of nn_mma::forward against an FP64 host evaluation, next to the shipped FFMA + tanh_fast arithmetic
//   2. cycles per network evaluation with 4 / 7 / 8 warps per SM running a dependent recurrence (output feeds input)
// Decides whether K1's network should move from shared-memory-fed FFMA2 to register-level MMAs
// (profiles/r01_autorally_k1_notes.md).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/mma_probe.cu
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mppi-generic_b200/csrc/plugins/nn_mma.cuh"

using namespace mppib;

__device__ __forceinline__ float tanh_fast(float x)
{
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}

// scalar FP32 evaluation in the reference's order (k ascending, bias last) with tanh_fast: the shipped arithmetic
__device__ void forward_ffma(const float* __restrict__ g, const float (&in)[6], float (&out)[4])
{
  float a1[32], a2[32];
  for (int j = 0; j < 32; j++)
  {
    float s = 0.0f;
    for (int k = 0; k < 6; k++)
      s = fmaf(g[j * 6 + k], in[k], s);
    a1[j] = tanh_fast(s + g[192 + j]);
  }
  for (int j = 0; j < 32; j++)
  {
    float s = 0.0f;
    for (int k = 0; k < 32; k++)
      s = fmaf(g[224 + j * 32 + k], a1[k], s);
    a2[j] = tanh_fast(s + g[1248 + j]);
  }
  for (int j = 0; j < 4; j++)
  {
    float s = 0.0f;
    for (int k = 0; k < 32; k++)
      s = fmaf(g[1280 + j * 32 + k], a2[k], s);
    out[j] = s + g[1408 + j];
  }
}

template <int MMA>
__global__ void __launch_bounds__(256) probe(const float* __restrict__ gw, const float* __restrict__ inputs, int steps,
                                            float* outputs, long long* cycles)
{
  extern __shared__ __align__(16) float theta_s[];
  if (MMA)
    nn_mma::load_weights(gw, theta_s);
  __syncthreads();
  float* scratch = theta_s + nn_mma::kFixedFloats + (threadIdx.x >> 5) * nn_mma::kScratchPerWarp;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  float in[6], out[4];
  for (int k = 0; k < 6; k++)
    in[k] = inputs[gid * 6 + k];
  const long long t0 = clock64();
  for (int s = 0; s < steps; s++)
  {
    if (MMA == 1)
      nn_mma::forward<0>(theta_s, scratch, in, out);
    else if (MMA == 2)
      nn_mma::forward<1>(theta_s, scratch, in, out);
    else if (MMA == 3)
      nn_mma::forward<2>(theta_s, scratch, in, out);
    else
      forward_ffma(gw, in, out);
    if (s + 1 < steps)
    {  // recurrence: the next input depends on this output (like the state update of the rollout)
      for (int k = 0; k < 4; k++)
        in[k] = 0.9f * in[k] + 0.05f * out[k];
      in[4] = 0.9f * in[4] + 0.02f * out[0];
      in[5] = 0.9f * in[5] - 0.02f * out[3];
    }
  }
  const long long t1 = clock64();
  for (int k = 0; k < 4; k++)
    outputs[gid * 4 + k] = out[k];
  if (threadIdx.x == 0 && blockIdx.x == 0)
    *cycles = t1 - t0;
}

static void host_forward(const std::vector<float>& g, const double* in, double* out)
{
  double a1[32], a2[32];
  for (int j = 0; j < 32; j++)
  {
    double s = 0;
    for (int k = 0; k < 6; k++)
      s += (double)g[j * 6 + k] * in[k];
    a1[j] = tanh(s + g[192 + j]);
  }
  for (int j = 0; j < 32; j++)
  {
    double s = 0;
    for (int k = 0; k < 32; k++)
      s += (double)g[224 + j * 32 + k] * a1[k];
    a2[j] = tanh(s + g[1248 + j]);
  }
  for (int j = 0; j < 4; j++)
  {
    double s = 0;
    for (int k = 0; k < 32; k++)
      s += (double)g[1280 + j * 32 + k] * a2[k];
    out[j] = s + g[1408 + j];
  }
}

int main()
{
  std::vector<float> w(1412);
  srand(1);
  auto uni = []() { return (rand() / (float)RAND_MAX) * 2.0f - 1.0f; };
  for (int i = 0; i < 192; i++)
    w[i] = uni() / sqrtf(6.0f);
  for (int i = 192; i < 224; i++)
    w[i] = 0.1f * uni();
  for (int i = 224; i < 1248; i++)
    w[i] = uni() / sqrtf(32.0f) * 2.0f;
  for (int i = 1248; i < 1280; i++)
    w[i] = 0.1f * uni();
  for (int i = 1280; i < 1408; i++)
    w[i] = uni() / sqrtf(32.0f) * 2.0f;
  for (int i = 1408; i < 1412; i++)
    w[i] = 0.1f * uni();
  const int max_threads = 148 * 256;
  std::vector<float> in(max_threads * 6);
  for (auto& v : in)
    v = 2.0f * uni();
  float *gw, *gin, *gout;
  long long* cyc_d;
  cudaMalloc(&gw, w.size() * 4);
  cudaMalloc(&gin, in.size() * 4);
  cudaMalloc(&gout, max_threads * 4 * 4);
  cudaMalloc(&cyc_d, 8);
  cudaMemcpy(gw, w.data(), w.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(gin, in.data(), in.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = nn_mma::sharedFloats(256) * sizeof(float);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);

  // ---- accuracy: one evaluation
  std::vector<float> o_mma(max_threads * 4), o_ffma(max_threads * 4);
  std::vector<float> o_mma2(max_threads * 4), o_mma3(max_threads * 4);
  probe<2><<<148, 224, smem>>>(gw, gin, 1, gout, cyc_d);
  cudaDeviceSynchronize();
  cudaMemcpy(o_mma2.data(), gout, 148 * 224 * 16, cudaMemcpyDeviceToHost);
  probe<3><<<148, 224, smem>>>(gw, gin, 1, gout, cyc_d);
  cudaDeviceSynchronize();
  cudaMemcpy(o_mma3.data(), gout, 148 * 224 * 16, cudaMemcpyDeviceToHost);
  probe<1><<<148, 224, smem>>>(gw, gin, 1, gout, cyc_d);
  cudaError_t e1 = cudaDeviceSynchronize();
  cudaMemcpy(o_mma.data(), gout, 148 * 224 * 16, cudaMemcpyDeviceToHost);
  probe<0><<<148, 224, smem>>>(gw, gin, 1, gout, cyc_d);
  cudaError_t e2 = cudaDeviceSynchronize();
  cudaMemcpy(o_ffma.data(), gout, 148 * 224 * 16, cudaMemcpyDeviceToHost);
  double err_mma = 0, err_mma2 = 0, err_mma3 = 0, err_ffma = 0, mag = 0;
  for (int s = 0; s < 148 * 224; s++)
  {
    double di[6], dout[4];
    for (int k = 0; k < 6; k++)
      di[k] = in[s * 6 + k];
    host_forward(w, di, dout);
    for (int k = 0; k < 4; k++)
    {
      err_mma = fmax(err_mma, fabs(o_mma[s * 4 + k] - dout[k]));
      err_mma2 = fmax(err_mma2, fabs(o_mma2[s * 4 + k] - dout[k]));
      err_mma3 = fmax(err_mma3, fabs(o_mma3[s * 4 + k] - dout[k]));
      err_ffma = fmax(err_ffma, fabs(o_ffma[s * 4 + k] - dout[k]));
      mag = fmax(mag, fabs(dout[k]));
    }
  }
  printf("accuracy vs FP64 over %d samples (|out| up to %.3f): mma f16 x3 max abs err %.3e (Newton rcp: %.3e, pair rcp: %.3e), FFMA + tanh_fast %.3e  (%s / %s)\n",
         148 * 224, mag, err_mma, err_mma2, err_mma3, err_ffma, cudaGetErrorString(e1), cudaGetErrorString(e2));

  // ---- timing: dependent recurrence of 100 evaluations
  const int steps = 100;
  for (int threads : { 128, 224, 256 })
  {
    for (int mode = 0; mode < 4; mode++)
    {
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      for (int rep = 0; rep < 2; rep++)
      {
        cudaEventRecord(a);
        if (mode == 0)
          probe<1><<<148, threads, smem>>>(gw, gin, steps, gout, cyc_d);
        else if (mode == 1)
          probe<2><<<148, threads, smem>>>(gw, gin, steps, gout, cyc_d);
        else if (mode == 3)
          probe<3><<<148, threads, smem>>>(gw, gin, steps, gout, cyc_d);
        else if (threads == 128)
          probe<0><<<148, threads, smem>>>(gw, gin, steps, gout, cyc_d);
        cudaEventRecord(b);
        cudaDeviceSynchronize();
      }
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      long long cyc = 0;
      cudaMemcpy(&cyc, cyc_d, 8, cudaMemcpyDeviceToHost);
      printf("%-22s %3d threads/SM (%d warps): %8.1f cycles per evaluation (block 0), %7.1f us per %d steps\n",
             mode == 0 ? "mma f16 x3, MUFU rcp" : (mode == 1 ? "mma f16 x3, Newton rcp" : (mode == 3 ? "mma f16 x3, pair rcp" : "scalar FFMA (global w)")), threads, threads / 32, (double)cyc / steps,
             ms * 1000.0, steps);
    }
  }
  return 0;
}
