/*
 * rollout_kernel_nn_tc.cuh — K1 for the Autorally pair (NeuralNetModel<7,2,3> + ARStandardCost) with the 6-32-32-4
 * forward pass on the 5th-generation tensor cores (tcgen05 + TMEM). Same contract and outputs as the generic
 * rollout_kernel (rollout_kernel.cuh); only the place where the 1344 multiply-adds per step happen changes.
 *
 * Mapping. One CTA = 128 threads = 128 samples = one UMMA tile (M = 128): thread i owns sample i AND accumulator row i
 * (TMEM lane i), so the per-sample code (control sampling, constraints, kinematics, Euler update, map cost) stays
 * thread-private exactly as in the SIMT kernel. Per layer the threads write their activations to shared memory in the
 * UMMA K-major / no-swizzle canonical layout ([k/4][row] slabs of 16 B), one elected thread issues the tcgen05.mma
 * sequence, completion arrives on an mbarrier via tcgen05.commit, and every thread reads its own 32 pre-activations back
 * with tcgen05.ld (32x32b).
 *
 * Precision. kind::tf32 alone (10-bit mantissa) is ~4e-4 off per layer and fails the FP32 parity bar of a 100-step
 * recurrence, so every product is the 3xTF32 split  a_hi*w_hi + a_lo*w_hi + a_hi*w_lo  (a_hi = rna_tf32(a), a_lo = a - a_hi);
 * tools/tc_probe.cu measures 5.5e-7 max error on a 128x32x32 tile against FP64 (plain TF32: 3.8e-4). Biases ride an
 * extra K block against a constant-one activation column; weights and biases of the two tanh layers are pre-scaled by
 * 2*log2(e) so tanh(z) = 1 - 2/(exp2(z') + 1) needs no multiply.
 *
 * Shared memory per CTA ~82 KB (2-slab noise ring 32 KB, activations hi/lo 36 KB, weights hi/lo 14 KB, small tables),
 * so two CTAs are resident per SM and one covers the other's MMA round trips; the noise rows stream through a
 * two-slab TMA ring (the whole-horizon tile of the SIMT kernel would not leave room), and the epilogue re-streams them
 * (L2-resident) to form the block's exp-weighted control sum.
 */
#pragma once
#include "rollout_kernel.cuh"
#include "plugins/costs.cuh"
#include "plugins/dynamics.cuh"

namespace mppib
{
// MMA completion wait. Measured on B200: parking the thread between polls (mbarrier.try_wait with a suspend-time hint),
// or letting only warp 0 poll while the others sit at the CTA barrier, changes nothing (379-383 us); neither does starting
// the two CTAs of an SM half a step apart. One CTA per SM alone takes 320 us: the per-step chain of a tile (stage -> sync ->
// MMA round trip -> tcgen05.ld -> 32 tanh at 2 MUFU each on one warp per scheduler -> split -> store -> fence, three times)
// is ~6300 cycles, longer than the FFMA2 kernel's ~4500 for a lone warp (profiles/r01_autorally_k1_notes.md).
#define MMA_WAIT(bar, phase) mbar_wait(bar, phase)

namespace nn_tc
{
constexpr int kRows = 128;          // samples per CTA == UMMA M
constexpr int kSlabBytes = kRows * kChunkBytes;
constexpr int kAChunks = 10;        // 8 activation chunks (K = 32) + 2 chunks holding the constant-one column
constexpr float kTanhScale = 2.8853900817779268f;  // 2 * log2(e)

struct Smem
{
  uint32_t slabs, a_hi, a_lo, w1_hi, w1_lo, w2_hi, w2_lo, w3_hi, w3_lo, b3, means, theta_c, weights, red, bars, tmem, total;
};
__host__ __device__ inline Smem layout(int TC, int T)
{
  Smem s;
  uint32_t off = 0;
  s.slabs = off;
  off += 2 * kSlabBytes;
  s.a_hi = off;
  off += kAChunks * kRows * 16;
  s.a_lo = off;
  off += 8 * kRows * 16;
  s.w1_hi = off;
  off += 2 * 32 * 16;
  s.w1_lo = off;
  off += 2 * 32 * 16;
  s.w2_hi = off;
  off += kAChunks * 32 * 16;
  s.w2_lo = off;
  off += kAChunks * 32 * 16;
  s.w3_hi = off;
  off += kAChunks * 8 * 16;
  s.w3_lo = off;
  off += kAChunks * 8 * 16;
  s.b3 = off;
  off += 16;
  s.means = off;
  off += ((uint32_t)(TC + 3) / 4) * 16;
  s.theta_c = off;
  off += ((uint32_t)(T + 3) / 4) * 16;
  s.weights = off;
  off += kRows * 4;
  s.red = off;
  off += 4 * 32 * 4 + 3 * 32 * 4;
  s.bars = off;
  off += 4 * 8;
  s.tmem = off;
  off += 16;
  s.total = off + 1024;
  return s;
}

__device__ __forceinline__ float rna_tf32(float x)
{
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout): element (row, k) lives at
// base + (k/4)*LBO + (row/8)*SBO + (row%8)*16 + (k%4)*4
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M x N
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N)
{
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
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
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before()
{
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after()
{
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; i++)
    v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4])
{
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; i++)
    v[i] = __uint_as_float(r[i]);
}
// tanh of a pre-activation that was already scaled by 2*log2(e)
__device__ __forceinline__ float tanh_prescaled(float zs)
{
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(zs));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}
// 32 activations -> hi/lo TF32 parts -> the thread's row of chunks 0..7 of the A operand
__device__ __forceinline__ void store_activations(float4* a_hi, float4* a_lo, int row, const float (&a)[32])
{
#pragma unroll
  for (int kc = 0; kc < 8; kc++)
  {
    float h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
      h[i] = rna_tf32(a[kc * 4 + i]);
      l[i] = a[kc * 4 + i] - h[i];  // exact in FP32; the tensor core keeps its leading 11 bits
    }
    a_hi[kc * kRows + row] = make_float4(h[0], h[1], h[2], h[3]);
    a_lo[kc * kRows + row] = make_float4(l[0], l[1], l[2], l[3]);
  }
}
}  // namespace nn_tc

template <bool WRITEBACK>
__global__ void __launch_bounds__(nn_tc::kRows, 2)
    rollout_kernel_ar_tc(const __grid_constant__ RolloutArgs<plugins::AutorallyNNDynamics, plugins::ARStandardCost> args,
                         const __grid_constant__ CUtensorMap tmap)
{
  using namespace nn_tc;
  using DYN = plugins::AutorallyNNDynamics;
  using COST = plugins::ARStandardCost;
  constexpr int S = 7, C = 2, O = 8;

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = args.T, TC = T * C, nchunks = args.nchunks;
  const Smem L = layout(TC, T);
  unsigned char* slabs = smem + L.slabs;
  float4* a_hi = reinterpret_cast<float4*>(smem + L.a_hi);
  float4* a_lo = reinterpret_cast<float4*>(smem + L.a_lo);
  float4* w1_hi = reinterpret_cast<float4*>(smem + L.w1_hi);
  float4* w1_lo = reinterpret_cast<float4*>(smem + L.w1_lo);
  float4* w2_hi = reinterpret_cast<float4*>(smem + L.w2_hi);
  float4* w2_lo = reinterpret_cast<float4*>(smem + L.w2_lo);
  float4* w3_hi = reinterpret_cast<float4*>(smem + L.w3_hi);
  float4* w3_lo = reinterpret_cast<float4*>(smem + L.w3_lo);
  float* means_s = reinterpret_cast<float*>(smem + L.means);
  float* theta_c = reinterpret_cast<float*>(smem + L.theta_c);
  float* w_s = reinterpret_cast<float*>(smem + L.weights);
  float* red_s = reinterpret_cast<float*>(smem + L.red);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);  // [0],[1]: slab full; [2]: MMA done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem);

  pdl_launch_dependents();
  const int row0 = blockIdx.x * kRows;
  const int n_loc = row0 + tid;
  const bool valid = n_loc < args.n_local;
  const int n_glob = args.n_offset + n_loc;

  // ---- one-time setup: TMEM, barriers, first two noise slabs, weights ---------------------------------------------
  if (warp == 0)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0)
  {
    tma_prefetch_desc(&tmap);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0)
  {
    for (int k = 0; k < 2 && k < nchunks; k++)
    {
      mbar_arrive_expect_tx(&bars[k], kSlabBytes);
      tma_load_2d(slabs + k * kSlabBytes, &tmap, k * kChunkFloats, row0, &bars[k]);
    }
  }
  {
    // weights: reference packed layout (fnn_helper.cu:176-183) -> K-major [k/4][n] float4 slabs, hi/lo TF32 parts;
    // k == fan_in is the bias column; the tanh layers are pre-scaled by 2*log2(e)
    const float* g = args.dyn_aux.theta_d;
    auto put = [](float4* hi, float4* lo, int idx, int comp, float v) {
      const float h = rna_tf32(v);
      reinterpret_cast<float*>(&hi[idx])[comp] = h;
      reinterpret_cast<float*>(&lo[idx])[comp] = v - h;
    };
    for (int i = tid; i < 2 * 32 * 4; i += kRows)
    {  // layer 1: K = 8 (6 inputs, bias at k = 6, zero at k = 7), N = 32
      const int kc = i / (32 * 4), n = (i / 4) % 32, c = i % 4, k = kc * 4 + c;
      const float v = (k < 6) ? g[n * 6 + k] : (k == 6 ? g[192 + n] : 0.0f);
      put(w1_hi, w1_lo, kc * 32 + n, c, v * kTanhScale);
    }
    for (int i = tid; i < kAChunks * 32 * 4; i += kRows)
    {  // layer 2: K = 32 (+ bias at k = 32), N = 32
      const int kc = i / (32 * 4), n = (i / 4) % 32, c = i % 4, k = kc * 4 + c;
      const float v = (k < 32) ? g[224 + n * 32 + k] : (k == 32 ? g[1248 + n] : 0.0f);
      put(w2_hi, w2_lo, kc * 32 + n, c, v * kTanhScale);
    }
    for (int i = tid; i < kAChunks * 8 * 4; i += kRows)
    {  // layer 3: K = 32 (+ bias), N = 8 (4 outputs, rows 4..7 zero), linear
      const int kc = i / (8 * 4), n = (i / 4) % 8, c = i % 4, k = kc * 4 + c;
      float v = 0.0f;
      if (n < 4)
        v = (k < 32) ? g[1280 + n * 32 + k] : (k == 32 ? g[1408 + n] : 0.0f);
      put(w3_hi, w3_lo, kc * 8 + n, c, v);
    }
    // constant-one activation column (chunk 8 = (1,0,0,0), chunk 9 = 0) — written once
    a_hi[8 * kRows + tid] = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
    a_hi[9 * kRows + tid] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  }
  for (int i = tid; i < TC; i += kRows)
    means_s[i] = args.means[i];
  COST::initializeCosts(args.cost, args.cost_aux, theta_c, T);

  float x[S], y[O];
#pragma unroll
  for (int i = 0; i < S; i++)
    x[i] = args.x0[i];
#pragma unroll
  for (int i = 0; i < O; i++)
    y[i] = 0.0f;
#pragma unroll
  for (int i = 0; i < 7; i++)
    y[i] = x[i];  // initializeDynamics (dynamics.cuh:429-435)
  float running_cost = 0.0f;
  int crash_status = 0;

  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc32 = umma_idesc_tf32(kRows, 32), idesc8 = umma_idesc_tf32(kRows, 8);
  const uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo);
  uint32_t mma_phase = 0;

  const bool pure_noise = (float)n_glob >= args.samp.pure_noise_threshold;
  const bool zero_noise_sample = (n_glob == 0);
  float lr_scale[C];
  bool lr_on = false;
#pragma unroll
  for (int c = 0; c < C; c++)
  {
    lr_scale[c] = args.samp.control_cost_coeff[c] / (args.samp.std_dev[0][c] * args.samp.std_dev[0][c]);
    lr_on = lr_on || (args.samp.control_cost_coeff[c] != 0.0f);
  }
  const float half_lambda_1ma = 0.5f * args.lambda * (1.0f - args.alpha);

  // issues the MMAs of one layer: K = 4*nkc real columns (+ the bias block when `bias`), W slabs of N rows
  auto issue_layer = [&](uint32_t w_hi_s, uint32_t w_lo_s, int N, int nkb, bool bias_block, uint32_t idesc) {
    uint32_t acc = 0;
    for (int kb = 0; kb < nkb; kb++)
    {
      const uint64_t dah = umma_desc(a_hi_s + kb * 2 * kRows * 16, kRows * 16, 128);
      const uint64_t dal = umma_desc(a_lo_s + kb * 2 * kRows * 16, kRows * 16, 128);
      const uint64_t dwh = umma_desc(w_hi_s + kb * 2 * N * 16, N * 16, 128);
      const uint64_t dwl = umma_desc(w_lo_s + kb * 2 * N * 16, N * 16, 128);
      umma_tf32(tmem_base, dah, dwh, idesc, acc);
      umma_tf32(tmem_base, dal, dwh, idesc, 1);
      umma_tf32(tmem_base, dah, dwl, idesc, 1);
      acc = 1;
    }
    if (bias_block)
    {  // constant-one column (chunks 8,9 of A) x bias row (chunks 8,9 of W): hi and lo part of the bias
      const uint64_t dah = umma_desc(a_hi_s + 8 * kRows * 16, kRows * 16, 128);
      const uint64_t dwh = umma_desc(w_hi_s + 8 * N * 16, N * 16, 128);
      const uint64_t dwl = umma_desc(w_lo_s + 8 * N * 16, N * 16, 128);
      umma_tf32(tmem_base, dah, dwh, idesc, 1);
      umma_tf32(tmem_base, dah, dwl, idesc, 1);
    }
    umma_commit(&bars[2]);
  };
  const uint32_t w1h = smem_u32(w1_hi), w1l = smem_u32(w1_lo), w2h = smem_u32(w2_hi), w2l = smem_u32(w2_lo),
                 w3h = smem_u32(w3_hi), w3l = smem_u32(w3_lo);

  // ---- the horizon -------------------------------------------------------------------------------------------------
  uint32_t slab_use[2] = { 0, 0 };
  for (int k = 0; k < nchunks; k++)
  {
    const int buf = k & 1;
    mbar_wait(&bars[buf], slab_use[buf] & 1);
    slab_use[buf]++;
    const unsigned char* slab = slabs + buf * kSlabBytes;
#pragma unroll 1
    for (int g = 0; g < 8; g++)
    {
      const int col0 = k * kChunkFloats + g * 4;
      if (col0 >= TC)
        break;
      const float4 e4 = *reinterpret_cast<const float4*>(slab + tid * kChunkBytes + ((g ^ (tid & 7)) << 4));
#pragma unroll 1
      for (int s = 0; s < 2; s++)
      {
        const int t = col0 / C + s;
        if (t >= T)
          break;
        const bool use_mean = zero_noise_sample || (t < args.opt_stride);
        const float* mean_t = means_s + t * C;
        float u[C];
        u[0] = sample_control(mean_t[0], args.samp.std_dev_decayed[0][0], s == 0 ? e4.x : e4.z, use_mean, pure_noise);
        u[1] = sample_control(mean_t[1], args.samp.std_dev_decayed[0][1], s == 0 ? e4.y : e4.w, use_mean, pure_noise);
        DYN::enforceConstraints(args.dyn, x, u);
        if (WRITEBACK)
        {
          if (valid)
          {
            float* dst = args.controls_out + ((size_t)n_loc * T + t) * C;
            dst[0] = u[0];
            dst[1] = u[1];
          }
        }
        float xdot[S];
        DYN::computeKinematics(args.dyn, x, xdot);  // ar_nn_model.cu:123-128

        // ---- layer 1: inputs (roll, vx, vy, yaw rate, steering, throttle, 1, 0) ------------------------------------
        {
          const float in[8] = { x[3], x[4], x[5], x[6], u[0], u[1], 1.0f, 0.0f };
          float h[8], l[8];
#pragma unroll
          for (int i = 0; i < 8; i++)
          {
            h[i] = rna_tf32(in[i]);
            l[i] = in[i] - h[i];
          }
          a_hi[tid] = make_float4(h[0], h[1], h[2], h[3]);
          a_hi[kRows + tid] = make_float4(h[4], h[5], h[6], h[7]);
          a_lo[tid] = make_float4(l[0], l[1], l[2], l[3]);
          a_lo[kRows + tid] = make_float4(l[4], l[5], l[6], l[7]);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0)
        {
          tc_fence_after();
          issue_layer(w1h, w1l, 32, 1, false, idesc32);
          // every thread has consumed this slab's last group once it passed the barrier above: refill the buffer
          if (g == 7 && s == 0 && k + 2 < nchunks)
          {
            mbar_arrive_expect_tx(&bars[buf], kSlabBytes);
            tma_load_2d(slabs + buf * kSlabBytes, &tmap, (k + 2) * kChunkFloats, row0, &bars[buf]);
          }
        }
        MMA_WAIT(&bars[2], mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        float act[32];
        tmem_ld32(tmem_row, act);
#pragma unroll
        for (int i = 0; i < 32; i++)
          act[i] = tanh_prescaled(act[i]);
        store_activations(a_hi, a_lo, tid, act);

        // ---- layer 2 ------------------------------------------------------------------------------------------------
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0)
        {
          tc_fence_after();
          issue_layer(w2h, w2l, 32, 4, true, idesc32);
        }
        MMA_WAIT(&bars[2], mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        tmem_ld32(tmem_row, act);
#pragma unroll
        for (int i = 0; i < 32; i++)
          act[i] = tanh_prescaled(act[i]);
        store_activations(a_hi, a_lo, tid, act);

        // ---- layer 3 (linear, 4 outputs) ----------------------------------------------------------------------------
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0)
        {
          tc_fence_after();
          issue_layer(w3h, w3l, 8, 4, true, idesc8);
        }
        MMA_WAIT(&bars[2], mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        float out4[4];
        tmem_ld4(tmem_row, out4);
        xdot[3] = out4[0];
        xdot[4] = out4[1];
        xdot[5] = out4[2];
        xdot[6] = out4[3];

        // ---- Euler update, output, costs (dynamics.cu:118-155, mppi_common.cu:120-128) --------------------------------
#pragma unroll
        for (int i = 0; i < S; i++)
          x[i] = x[i] + xdot[i] * args.dt;
#pragma unroll
        for (int i = 0; i < S; i++)
          y[i] = x[i];
        float step_cost = COST::computeRunningCost(args.cost, args.cost_aux, theta_c, y, u, t, &crash_status);
        if (lr_on)
          step_cost += likelihood_ratio_cost<C>(lr_scale, mean_t, u, pure_noise, half_lambda_1ma);
        running_cost += step_cost;
      }
    }
  }

  // ---- per-sample cost and block partial (same as rollout_kernel) ------------------------------------------------------
  const float cost = running_cost / (float)T + COST::terminalCost(args.cost, args.cost_aux, y) / (float)T;
  if (valid)
    args.costs[n_loc] = cost;
  float* scratch = red_s + 4 * 32;
  {
    float m = warp_min(valid ? cost : INFINITY);
    if (lane == 0)
      scratch[warp] = m;
    __syncthreads();
    float beta_b = scratch[0];
    for (int i = 1; i < 4; i++)
      beta_b = fminf(beta_b, scratch[i]);
    const float w = valid ? expf(-args.lambda_inv * (cost - beta_b)) : 0.0f;
    w_s[tid] = w;
    const float sw = warp_sum(w), sw2 = warp_sum(w * w);
    if (lane == 0)
    {
      scratch[32 + warp] = sw;
      scratch[64 + warp] = sw2;
    }
    __syncthreads();
    if (tid == 0)
    {
      float eta_b = 0.0f, w2_b = 0.0f;
      for (int i = 0; i < 4; i++)
      {
        eta_b += scratch[32 + i];
        w2_b += scratch[64 + i];
      }
      args.headers[blockIdx.x] = make_float4(beta_b, eta_b, w2_b, 0.0f);
    }
  }

  // exp-weighted sum of the constrained controls: re-stream the noise slabs (L2-resident) through the same ring;
  // thread (q, j) = (tid / 32, tid % 32) sums column j over rows [32q, 32q + 32), then the 4 quarters are added.
  __syncthreads();
  const int rows_here = min(kRows, args.n_local - row0);
  if (tid == 0)
  {
    for (int k = 0; k < 2 && k < nchunks; k++)
    {
      mbar_arrive_expect_tx(&bars[k], kSlabBytes);
      tma_load_2d(slabs + k * kSlabBytes, &tmap, k * kChunkFloats, row0, &bars[k]);
    }
  }
  float* out = args.partials + (size_t)blockIdx.x * args.pstride + kPartialHeader;
  for (int k = 0; k < nchunks; k++)
  {
    const int buf = k & 1;
    mbar_wait(&bars[buf], slab_use[buf] & 1);
    slab_use[buf]++;
    const unsigned char* slab = slabs + buf * kSlabBytes;
    const int col = k * kChunkFloats + lane;
    float acc = 0.0f;
    if (col < TC)
    {
      const int t = col >> 1, c = col & 1;
      const float* mean_t = means_s + t * C;
      const bool t_uses_mean = t < args.opt_stride;
      const int pair = (lane >> 1) << 1;  // first column of this time step inside the slab
      const int r_begin = warp * 32, r_end = min(r_begin + 32, rows_here);
      for (int r = r_begin; r < r_end; r++)
      {
        const float2 e2 = *reinterpret_cast<const float2*>(slab + r * kChunkBytes + (((pair >> 2) ^ (r & 7)) << 4) +
                                                           ((pair & 3) << 2));
        const int ng = args.n_offset + row0 + r;
        const bool pn = (float)ng >= args.samp.pure_noise_threshold;
        const bool um = t_uses_mean || (ng == 0);
        float u[C];
        u[0] = sample_control(mean_t[0], args.samp.std_dev_decayed[0][0], e2.x, um, pn);
        u[1] = sample_control(mean_t[1], args.samp.std_dev_decayed[0][1], e2.y, um, pn);
        DYN::enforceConstraints(args.dyn, nullptr, u);
        acc = fmaf(w_s[r], c == 0 ? u[0] : u[1], acc);
      }
    }
    red_s[warp * 32 + lane] = acc;
    __syncthreads();
    if (tid < 32 && col < TC)
      out[col] = (red_s[lane] + red_s[32 + lane]) + (red_s[64 + lane] + red_s[96 + lane]);
    if (tid == 0 && k + 2 < nchunks)
    {  // every thread passed the barrier above, i.e. is done with this buffer
      mbar_arrive_expect_tx(&bars[buf], kSlabBytes);
      tma_load_2d(slabs + buf * kSlabBytes, &tmap, (k + 2) * kChunkFloats, row0, &bars[buf]);
    }
    __syncthreads();  // red_s reuse
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32));
}

}  // namespace mppib
