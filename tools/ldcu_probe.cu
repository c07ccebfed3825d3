// Constant-bank working-set probe (GPU box): cycles per 32x32 parameter-bank layer when a step cycles through NL
// different layers (NL * 4.2 KB of kernel parameters). Finds the size at which LDCU starts missing its cache.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
template <int NL>
struct Weights
{
  float w[NL][32 * 32 + 32];
};
template <int NL>
__global__ void __launch_bounds__(256) probe(const __grid_constant__ Weights<NL> pw, int reps, float* out, long long* cycles)
{
  float a[32];
#pragma unroll
  for (int i = 0; i < 32; i++)
    a[i] = 0.01f * (float)((threadIdx.x + i) & 7);
  const long long t0 = clock64();
  for (int r = 0; r < reps; r++)
  {
#pragma unroll
    for (int l = 0; l < NL; l++)
    {
      float acc[32];
#pragma unroll
      for (int j = 0; j < 32; j++)
        acc[j] = 0.0f;
#pragma unroll
      for (int k = 0; k < 32; k++)
#pragma unroll
        for (int j = 0; j < 32; j++)
          acc[j] = fmaf(pw.w[l][k * 32 + j], a[k], acc[j]);
#pragma unroll
      for (int j = 0; j < 32; j++)
        a[j] = (acc[j] + pw.w[l][1024 + j]) * 0.03125f;
    }
  }
  const long long t1 = clock64();
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < 32; j++)
    s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0)
    *cycles = t1 - t0;
}
template <int NL>
static void run(float* out, long long* cyc_d)
{
  static Weights<NL> hw;
  for (int l = 0; l < NL; l++)
    for (float& v : hw.w[l])
      v = ((rand() / (float)RAND_MAX) * 2 - 1) / 5.65f;
  const int reps = 100;
  for (int wps : { 1, 2 })
  {
    probe<NL><<<148, 128 * wps>>>(hw, reps, out, cyc_d);
    cudaDeviceSynchronize();
    probe<NL><<<148, 128 * wps>>>(hw, reps, out, cyc_d);
    cudaError_t err = cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, cyc_d, sizeof(cyc), cudaMemcpyDeviceToHost);
    printf("working set %5.1f KB, warps/SMSP %d: %8.1f cycles per layer per warp (%s)\n", NL * 4.125, wps,
           (double)cyc / reps / NL, cudaGetErrorString(err));
  }
}
int main()
{
  float* out;
  long long* cyc_d;
  cudaMalloc(&out, 148 * 256 * sizeof(float));
  cudaMalloc(&cyc_d, sizeof(long long));
  run<1>(out, cyc_d);
  run<2>(out, cyc_d);
  run<3>(out, cyc_d);
  run<4>(out, cyc_d);
  run<6>(out, cyc_d);
  return 0;
}
