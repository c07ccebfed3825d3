// FFMA operand-form probe (GPU box): how fast does one SM sub-partition issue the 32x32 dense layer of the Autorally
// network when the weights come from (a) the kernel-parameter constant bank as FFMA constant operands, (b) a __constant__
// array, (c) shared memory as broadcast LDS.128 feeding FFMA2 (the shipped K1 path)? Prints cycles per layer evaluation
// per warp for 1, 2 and 4 warps per scheduler. Decides whether K1's weights should move from shared memory to the
// constant bank (profiles/r01_autorally_k1_notes.md).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct Weights
{
  float w[32 * 32 + 32];  // WT[k][j] then bias
};
__constant__ Weights c_weights;

__device__ __forceinline__ float tanh_fast(float x)
{
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}

template <int MODE, bool TANH>
__global__ void __launch_bounds__(512) probe(const __grid_constant__ Weights pw, const float* __restrict__ gw, int reps,
                                            float* out, long long* cycles)
{
  __shared__ __align__(16) float sw[32 * 32 + 32];
  for (int i = threadIdx.x; i < 32 * 32 + 32; i += blockDim.x)
    sw[i] = gw[i];
  __syncthreads();
  float a[32];
#pragma unroll
  for (int i = 0; i < 32; i++)
    a[i] = 0.01f * (float)((threadIdx.x + i) & 7);
  const long long t0 = clock64();
  for (int r = 0; r < reps; r++)
  {
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; j++)
      acc[j] = 0.0f;
    if (MODE == 0)
    {  // kernel-parameter constant bank
#pragma unroll
      for (int k = 0; k < 32; k++)
#pragma unroll
        for (int j = 0; j < 32; j++)
          acc[j] = fmaf(pw.w[k * 32 + j], a[k], acc[j]);
#pragma unroll
      for (int j = 0; j < 32; j++)
        acc[j] += pw.w[1024 + j];
    }
    else if (MODE == 1)
    {  // __constant__ array
#pragma unroll
      for (int k = 0; k < 32; k++)
#pragma unroll
        for (int j = 0; j < 32; j++)
          acc[j] = fmaf(c_weights.w[k * 32 + j], a[k], acc[j]);
#pragma unroll
      for (int j = 0; j < 32; j++)
        acc[j] += c_weights.w[1024 + j];
    }
    else
    {  // shared memory, broadcast LDS.128 + FFMA2 (two adjacent outputs per instruction)
      float2 acc2[16];
#pragma unroll
      for (int j = 0; j < 16; j++)
        acc2[j] = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int k = 0; k < 32; k++)
      {
        const float2 xk = make_float2(a[k], a[k]);
#pragma unroll
        for (int j4 = 0; j4 < 8; j4++)
        {
          const float4 w = *reinterpret_cast<const float4*>(&sw[k * 32 + 4 * j4]);
          acc2[2 * j4] = __ffma2_rn(make_float2(w.x, w.y), xk, acc2[2 * j4]);
          acc2[2 * j4 + 1] = __ffma2_rn(make_float2(w.z, w.w), xk, acc2[2 * j4 + 1]);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; j++)
      {
        acc[2 * j] = acc2[j].x + sw[1024 + 2 * j];
        acc[2 * j + 1] = acc2[j].y + sw[1024 + 2 * j + 1];
      }
    }
#pragma unroll
    for (int j = 0; j < 32; j++)
      a[j] = TANH ? tanh_fast(acc[j]) : acc[j] * 0.03125f;
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

template <int MODE, bool TANH>
static void run(const char* name, const Weights& hw, const float* gw, float* out, long long* cyc_d)
{
  const int reps = 200;
  for (int warps_per_smsp : { 1, 2, 4 })
  {
    const int threads = 32 * 4 * warps_per_smsp;
    probe<MODE, TANH><<<148, threads>>>(hw, gw, reps, out, cyc_d);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE, TANH><<<148, threads>>>(hw, gw, reps, out, cyc_d);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long cyc = 0;
    cudaMemcpy(&cyc, cyc_d, sizeof(cyc), cudaMemcpyDeviceToHost);
    printf("%-34s warps/SMSP %d: %8.1f cycles per layer per warp, %7.1f cycles per layer per SMSP, %.3f ms (%s)\n", name,
           warps_per_smsp, (double)cyc / reps, (double)cyc / reps / warps_per_smsp, ms, cudaGetErrorString(err));
  }
}

int main()
{
  Weights hw;
  srand(1);
  for (float& v : hw.w)
    v = ((rand() / (float)RAND_MAX) * 2 - 1) / 5.65f;
  float *gw, *out;
  long long* cyc_d;
  cudaMalloc(&gw, sizeof(hw));
  cudaMalloc(&out, 148 * 512 * sizeof(float));
  cudaMalloc(&cyc_d, sizeof(long long));
  cudaMemcpy(gw, hw.w, sizeof(hw), cudaMemcpyHostToDevice);
  cudaMemcpyToSymbol(c_weights, &hw, sizeof(hw));
  run<0, false>("param-bank FFMA", hw, gw, out, cyc_d);
  run<1, false>("__constant__ FFMA", hw, gw, out, cyc_d);
  run<2, false>("shared LDS.128 + FFMA2", hw, gw, out, cyc_d);
  run<0, true>("param-bank FFMA + tanh", hw, gw, out, cyc_d);
  run<2, true>("shared LDS.128 + FFMA2 + tanh", hw, gw, out, cyc_d);
  return 0;
}
