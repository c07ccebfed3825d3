// tcgen05 probe (GPU box): D[128 x N] = A[128 x K] * W[N x K]^T with kind::tf32 UMMA, A and W staged in shared memory in
// the K-major SWIZZLE_NONE canonical layout ([k/4][row] float4 slabs), accumulator in TMEM, read back with tcgen05.ld.
// Checks (a) single-pass TF32 and (b) the 3xTF32 split (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo) against an FP64 reference.
// This is the building block of the tensor-core NN forward pass; validated here before integration.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x)
{
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// K-major, no swizzle: element (row, k) at base + (k/4)*LBO + (row/8)*SBO + (row%8)*16 + (k%4)*4 bytes
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  // base_offset = 0, lbo_mode = 0, layout_type = 0 (SWIZZLE_NONE)
  return d;
}

__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N)
{
  uint32_t d = 0;
  d |= 1u << 4;                    // c_format = F32
  d |= 2u << 7;                    // a_format = TF32
  d |= 2u << 10;                   // b_format = TF32
  d |= 0u << 15;                   // a_major = K
  d |= 0u << 16;                   // b_major = K
  d |= (uint32_t)(N >> 3) << 17;   // n_dim
  d |= (uint32_t)(M >> 4) << 24;   // m_dim
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int N, int K, bool SPLIT3>
__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ D)
{
  constexpr int KC = K / 4;
  __shared__ __align__(128) float4 a_hi[KC][128];
  __shared__ __align__(128) float4 a_lo[KC][128];
  __shared__ __align__(128) float4 w_hi[KC][N];
  __shared__ __align__(128) float4 w_lo[KC][N];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // stage operands
  for (int kc = 0; kc < KC; kc++)
  {
    float v[4], h[4], l[4];
    for (int i = 0; i < 4; i++)
    {
      v[i] = A[tid * K + kc * 4 + i];
      h[i] = to_tf32(v[i]);
      l[i] = to_tf32(v[i] - h[i]);
    }
    a_hi[kc][tid] = make_float4(h[0], h[1], h[2], h[3]);
    a_lo[kc][tid] = make_float4(l[0], l[1], l[2], l[3]);
  }
  for (int i = tid; i < KC * N; i += 128)
  {
    const int kc = i / N, n = i % N;
    float h[4], l[4];
    for (int j = 0; j < 4; j++)
    {
      const float v = W[n * K + kc * 4 + j];
      h[j] = to_tf32(v);
      l[j] = to_tf32(v - h[j]);
    }
    w_hi[kc][n] = make_float4(h[0], h[1], h[2], h[3]);
    w_lo[kc][n] = make_float4(l[0], l[1], l[2], l[3]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0)
  {
    const uint32_t idesc = make_idesc_tf32(128, N);
    uint32_t acc = 0;
    for (int kb = 0; kb < K / 8; kb++)
    {
      // A: slabs of 128 rows x 16 B => LBO (next k-group) = 2048 B, SBO (next 8 rows) = 128 B
      const uint64_t dah = make_desc(smem_u32(&a_hi[kb * 2][0]), 2048, 128);
      const uint64_t dal = make_desc(smem_u32(&a_lo[kb * 2][0]), 2048, 128);
      // W: slabs of N rows x 16 B => LBO = N*16 B, SBO = 128 B
      const uint64_t dwh = make_desc(smem_u32(&w_hi[kb * 2][0]), N * 16, 128);
      const uint64_t dwl = make_desc(smem_u32(&w_lo[kb * 2][0]), N * 16, 128);
      mma_tf32(tmem_base, dah, dwh, idesc, acc);
      acc = 1;
      if (SPLIT3)
      {
        mma_tf32(tmem_base, dal, dwh, idesc, 1);
        mma_tf32(tmem_base, dah, dwl, idesc, 1);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // wait for the MMAs
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(smem_u32(&mbar)), "r"(0)
          : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < N; j++)
    D[tid * N + j] = __uint_as_float(r[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32));
}


// Round-trip latency of one layer's MMAs: thread 0 issues them + commit, every thread then waits on the mbarrier
// (WAIT = 0: mbarrier.try_wait loop, 1: mbarrier.test_wait spin) and reads 32 accumulator columns back from TMEM.
template <int N, int K, int WAIT>
__global__ void __launch_bounds__(128) latency(float* out, long long* cycles, int reps)
{
  constexpr int KC = K / 4;
  __shared__ __align__(128) float4 a_hi[KC][128];
  __shared__ __align__(128) float4 a_lo[KC][128];
  __shared__ __align__(128) float4 w_hi[KC][N];
  __shared__ __align__(128) float4 w_lo[KC][N];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int kc = 0; kc < KC; kc++)
  {
    a_hi[kc][tid] = make_float4(0.5f, 0.25f, 0.125f, 1.0f);
    a_lo[kc][tid] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = tid; i < KC * N; i += 128)
  {
    w_hi[i / N][i % N] = make_float4(0.01f, 0.02f, 0.03f, 0.04f);
    w_lo[i / N][i % N] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = make_idesc_tf32(128, N);
  uint32_t phase = 0;
  float sum = 0.0f;
  const long long t0 = clock64();
  for (int r = 0; r < reps; r++)
  {
    if (tid == 0)
    {
      uint32_t acc = 0;
      for (int kb = 0; kb < K / 8; kb++)
      {
        const uint64_t dah = make_desc(smem_u32(&a_hi[kb * 2][0]), 2048, 128);
        const uint64_t dal = make_desc(smem_u32(&a_lo[kb * 2][0]), 2048, 128);
        const uint64_t dwh = make_desc(smem_u32(&w_hi[kb * 2][0]), N * 16, 128);
        const uint64_t dwl = make_desc(smem_u32(&w_lo[kb * 2][0]), N * 16, 128);
        mma_tf32(tmem_base, dah, dwh, idesc, acc);
        mma_tf32(tmem_base, dal, dwh, idesc, 1);
        mma_tf32(tmem_base, dah, dwl, idesc, 1);
        acc = 1;
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok)
    {
      if (WAIT == 0)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
    }
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[4];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    sum += __uint_as_float(v[0]) + __uint_as_float(v[3]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // the kernel's per-layer barrier: operands of the next layer are written by all threads
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  const long long t1 = clock64();
  out[blockIdx.x * 128 + tid] = sum;
  if (tid == 0 && blockIdx.x == 0)
    *cycles = t1 - t0;
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32));
}

template <int N, int K, int WAIT>
static void run_latency(const char* name, int ctas)
{
  float* out;
  long long* cyc;
  cudaMalloc(&out, ctas * 128 * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  const int reps = 300;
  latency<N, K, WAIT><<<ctas, 128>>>(out, cyc, reps);
  cudaDeviceSynchronize();
  latency<N, K, WAIT><<<ctas, 128>>>(out, cyc, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  printf("%s N=%d K=%d (%d MMAs) wait=%s ctas=%d: %.0f cycles per issue->commit->wait->tmem-ld->barrier round trip (%s)\n", name, N,
         K, 3 * K / 8, WAIT ? "test_wait" : "try_wait", ctas, (double)c / reps, cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(cyc);
}

template <int N, int K, bool SPLIT3>
static void run(const char* name)
{
  std::vector<float> A(128 * K), W(N * K), D(128 * N);
  srand(1);
  for (auto& v : A) v = (rand() / (float)RAND_MAX) * 2 - 1;
  for (auto& v : W) v = ((rand() / (float)RAND_MAX) * 2 - 1) / sqrtf((float)K);
  float *dA, *dW, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  probe<N, K, SPLIT3><<<1, 128>>>(dA, dW, dD);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; m++)
    for (int n = 0; n < N; n++)
    {
      double ref = 0;
      for (int k = 0; k < K; k++) ref += (double)A[m * K + k] * (double)W[n * K + k];
      maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("%s N=%d K=%d split3=%d: cuda=%s max|err|=%.3e (max|ref|=%.3f)  D[0][0..3]= %g %g %g %g\n", name, N, K, (int)SPLIT3,
         cudaGetErrorString(e), maxerr, maxref, D[0], D[1], D[2], D[3]);
  cudaFree(dA); cudaFree(dW); cudaFree(dD);
}

int main()
{
  run<32, 32, false>("L2 tf32   ");
  run<32, 32, true>("L2 3xtf32 ");
  run<32, 8, true>("L1 3xtf32 ");
  run<8, 32, true>("L3 3xtf32 ");
  run_latency<32, 32, 0>("L2", 148);
  run_latency<32, 32, 1>("L2", 148);
  run_latency<32, 8, 0>("L1", 148);
  run_latency<32, 8, 1>("L1", 148);
  run_latency<32, 32, 0>("L2", 296);
  run_latency<32, 32, 1>("L2", 296);
  return 0;
}
