/*
 * plugins/nn_mma.cuh — the Autorally 6-32-32-4 network evaluated by a WARP for its 32 samples with register-level tensor
 * core MMAs (mma.sync, FP16 inputs, FP32 accumulate); the reference is FNNHelper::forward
 * (include/mppi/utils/nn_helpers/fnn_helper.cu:419-484).
 *
 * Why: the FFMA2 kernel (plugins/dynamics.cuh) is bound by shared-memory wavefronts — every 4 FMAs of a thread need one
 * broadcast LDS.128 of weights, 672 wavefronts per warp-step, 72 % of the data pipe (profiles/r01_autorally_k1_notes.md).
 * Here the weights are B fragments (one LDS.128 per lane per tile carrying the hi and lo parts: 14 loads, ~50 wavefronts per warp-step) and the
 * activations never leave registers between layers: with 16-bit inputs the C fragment of n-tiles 2j, 2j+1 of layer l,
 * packed to half2, IS the A fragment of k-tile j of layer l+1.
 *
 * Precision: a single 11-bit-significand product (FP16 or TF32) is ~4e-4 off per layer and fails the FP32 parity bar of
 * a 100-step recurrence (rollout_kernel_nn_tc.cuh), so every operand is split v = hi + lo / 2048 with hi = half(v),
 * lo = half((v - hi) * 2048) (the scale keeps the residual out of FP16's subnormals) and every product is
 * hi*hi + (hi*lo + lo*hi) / 2048, the two cross terms in their own accumulator. The legacy tensor path issues one
 * mma.sync per ~14 cycles per SM sub-partition whatever the operand type (tools/mma_probe.cu: 3xTF32 m16n8k8, 144 per
 * evaluation, 1900-2100 cycles per warp-evaluation with two warps per scheduler), so FP16 m16n8k16 — twice the k per
 * instruction, 84 per evaluation — is the cheaper encoding of the same three-product scheme.
 *
 * Fragment layouts (PTX ISA; g = lane >> 2, t = lane & 3; every register is a half2 of adjacent columns / k):
 *   m16n8k16 A row:  a0 (g, 2t..)  a1 (g+8, 2t..)  a2 (g, 2t+8..)  a3 (g+8, 2t+8..)
 *   m16n8k16 B col:  b0 (k = 2t.., n = g)  b1 (k = 2t+8.., n = g)
 *   m16n8k8  A / B:  a0 a1 / b0 as above
 *   C 16x8 (f32):    c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1)
 * A warp's 32 samples are two m-tiles (rows 0-15, 16-31); sample s of the warp is lane s's.
 */
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mppib
{
namespace nn_mma
{
// shared-memory layout in floats (block-wide part, then 256 floats of scratch per warp)
constexpr int kW1F = 0;               // layer 1: 4 n-tiles x 32 lanes x (b0_hi, b0_lo)               uint2
constexpr int kW2F = kW1F + 4 * 64;   // layer 2: 8 tiles (n-tile i major, k-tile j minor) x 32 x (b0_hi, b1_hi, b0_lo, b1_lo)
constexpr int kW3F = kW2F + 8 * 128;  // layer 3: 2 k-tiles x 32 x (b0_hi, b1_hi, b0_lo, b1_lo)
constexpr int kB1 = kW3F + 2 * 128;   // 32 (pre-scaled like the weights)
constexpr int kB2 = kB1 + 32;         // 32
constexpr int kB3 = kB2 + 32;         // 8 (4 real)
constexpr int kFixedFloats = kB3 + 8;  // 1608
constexpr int kScratchPerWarp = 256;   // [2][32][4] input halves, reused as [32][4] outputs
__host__ __device__ constexpr int sharedFloats(int block_threads)
{
  return kFixedFloats + (block_threads / 32) * kScratchPerWarp;
}
// tanh(x) = 1 - 2 / (exp2(2 log2(e) x) + 1): the factor is folded into the weights and biases feeding a tanh
constexpr float kTanhScale = 2.8853900817779268f;
constexpr float kLoScale = 2048.0f, kLoInv = 1.0f / 2048.0f;

__device__ __forceinline__ uint32_t h2_bits(__half2 h)
{
  return *reinterpret_cast<uint32_t*>(&h);
}
// (v0, v1) -> half2 hi, half2 lo with v = hi + lo / 2048
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo)
{
  const __half2 h = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(h);
  hi = h2_bits(h);
#ifdef MPPIB_EXP_PACKED  // experimental (round 2): the residual with packed FP32x2 instructions (same IEEE results per lane)
  const float2 r = __fmul2_rn(__fadd2_rn(make_float2(v0, v1), make_float2(-hf.x, -hf.y)), make_float2(kLoScale, kLoScale));
  lo = h2_bits(__floats2half2_rn(r.x, r.y));
#else
  lo = h2_bits(__floats2half2_rn((v0 - hf.x) * kLoScale, (v1 - hf.y) * kLoScale));
#endif
}
__device__ __forceinline__ void mma16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma8(float (&c)[4], const uint32_t (&a)[2], uint32_t b0)
{
  asm("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(b0));
}
// tanh of two pre-scaled arguments z = 2 log2(e) x:  1 - 2 / (exp2(z) + 1), ex2 + rcp on the MUFU unit
__device__ __forceinline__ float2 tanh2_prescaled(float z0, float z1)
{
  float e0, e1, r0, r1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(z0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(z1));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(e0 + 1.0f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(e1 + 1.0f));
  return make_float2(fmaf(-2.0f, r0, 1.0f), fmaf(-2.0f, r1, 1.0f));
}
// Same, with the reciprocal moved to the FP32 pipe: integer-subtract seed (5 % off) + three Newton steps as packed FFMA2s,
// relative error < 1.5e-7. z is clamped at 64 so that exp2(z) + 1 stays finite for the seed. Timed by tools/mma_probe.cu.
__device__ __forceinline__ float2 tanh2_prescaled_newton(float z0, float z1)
{
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(z0, 64.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(z1, 64.0f)));
  const float2 d = __fadd2_rn(make_float2(e0, e1), make_float2(1.0f, 1.0f));
  float2 r = make_float2(__uint_as_float(0x7EF311C7u - __float_as_uint(d.x)),
                         __uint_as_float(0x7EF311C7u - __float_as_uint(d.y)));
  const float2 nd = make_float2(-d.x, -d.y), one = make_float2(1.0f, 1.0f);
#pragma unroll
  for (int it = 0; it < 3; it++)
  {
    const float2 err = __ffma2_rn(nd, r, one);  // 1 - d r
    r = __ffma2_rn(r, err, r);                  // r + r (1 - d r)
  }
  return __ffma2_rn(make_float2(-2.0f, -2.0f), r, one);
}
// One reciprocal for two values (experimental, tools/mma_probe.cu): 1/d0 = d1 / (d0 d1), 1/d1 = d0 / (d0 d1) — three MUFU per
// pair instead of four on the busiest pipe (XU 45 % in K1), two more FMULs on a short chain. z is clamped at 60 so that the
// product (<= 2^120) stays finite; tanh is 1 to the last bit long before that.
__device__ __forceinline__ float2 tanh2_prescaled_pair(float z0, float z1)
{
  float e0, e1, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(z0, 60.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(z1, 60.0f)));
  const float d0 = e0 + 1.0f, d1 = e1 + 1.0f;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d0 * d1));
  return make_float2(fmaf(-2.0f, r * d1, 1.0f), fmaf(-2.0f, r * d0, 1.0f));
}
// Finished tile (hi*hi sum in c, cross terms in x) -> tanh -> the two half2 registers (rows g and g+8) of the next layer's
// A fragment, hi and lo parts
// TANH: 0 = ex2 + rcp on MUFU (shipped), 1 = Newton reciprocal on the FP32 pipe (measured slower), 2 = one rcp per pair
template <int TANH>
__device__ __forceinline__ void activate(const float (&c)[4], const float (&x)[4], uint32_t& top_hi, uint32_t& bot_hi,
                                         uint32_t& top_lo, uint32_t& bot_lo)
{
#ifdef MPPIB_EXP_PACKED
  const float2 inv = make_float2(kLoInv, kLoInv);
  const float2 za = __ffma2_rn(make_float2(x[0], x[1]), inv, make_float2(c[0], c[1]));
  const float2 zb = __ffma2_rn(make_float2(x[2], x[3]), inv, make_float2(c[2], c[3]));
  const float z0 = za.x, z1 = za.y, z2 = zb.x, z3 = zb.y;
#else
  const float z0 = fmaf(x[0], kLoInv, c[0]), z1 = fmaf(x[1], kLoInv, c[1]), z2 = fmaf(x[2], kLoInv, c[2]),
              z3 = fmaf(x[3], kLoInv, c[3]);
#endif
  const float2 u = TANH == 1 ? tanh2_prescaled_newton(z0, z1) : (TANH == 2 ? tanh2_prescaled_pair(z0, z1) : tanh2_prescaled(z0, z1));
  const float2 v = TANH == 1 ? tanh2_prescaled_newton(z2, z3) : (TANH == 2 ? tanh2_prescaled_pair(z2, z3) : tanh2_prescaled(z2, z3));
  split2(u.x, u.y, top_hi, top_lo);
  split2(v.x, v.y, bot_hi, bot_lo);
}

// Block-cooperative: reference packed weights (per layer W row-major out x in, then b; fnn_helper.cu:176-183) ->
// fragment-ordered hi / lo half2 parts. Layers 1 and 2 feed a tanh, so their weights and biases carry kTanhScale.
__device__ __forceinline__ void load_weights(const float* __restrict__ g, float* theta_s)
{
  uint32_t* w1f = reinterpret_cast<uint32_t*>(theta_s + kW1F);
  uint32_t* w2f = reinterpret_cast<uint32_t*>(theta_s + kW2F);
  uint32_t* w3f = reinterpret_cast<uint32_t*>(theta_s + kW3F);
  for (int idx = threadIdx.x; idx < 14 * 32; idx += blockDim.x)
  {
    const int tile = idx >> 5, lane = idx & 31, gg = lane >> 2, t = lane & 3;
    if (tile < 4)
    {  // layer 1, n-tile i = tile, one k8 tile: k = 2t, 2t+1 (inputs 6, 7 are padding)
      const int n = 8 * tile + gg;
      const float v0 = (2 * t < 6) ? g[n * 6 + 2 * t] * kTanhScale : 0.0f;
      const float v1 = (2 * t + 1 < 6) ? g[n * 6 + 2 * t + 1] * kTanhScale : 0.0f;
      uint32_t hi, lo;
      split2(v0, v1, hi, lo);
      w1f[idx * 2 + 0] = hi;
      w1f[idx * 2 + 1] = lo;
    }
    else
    {
      const bool l2 = tile < 12;
      const int q = l2 ? tile - 4 : tile - 12;
      const int i = l2 ? (q >> 1) : 0, j = l2 ? (q & 1) : q;  // n-tile, k16-tile
      const int n = 8 * i + gg;
      const float* W = l2 ? g + 224 : g + 1280;
      const bool real = l2 || gg < 4;  // layer 3 has 4 outputs
      const float sc = l2 ? kTanhScale : 1.0f;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; e++)
      {
        const int k = 16 * j + 2 * t + (e & 1) + ((e >> 1) ? 8 : 0);
        v[e] = real ? W[n * 32 + k] * sc : 0.0f;
      }
      uint32_t h0, l0, h1, l1;
      split2(v[0], v[1], h0, l0);
      split2(v[2], v[3], h1, l1);
      uint32_t* dst = (l2 ? w2f : w3f) + (q * 32 + lane) * 4;
      dst[0] = h0, dst[1] = h1, dst[2] = l0, dst[3] = l1;
    }
  }
  for (int i = threadIdx.x; i < 72; i += blockDim.x)
  {
    float v;
    if (i < 32)
      v = g[192 + i] * kTanhScale;
    else if (i < 64)
      v = g[1248 + (i - 32)] * kTanhScale;
    else
      v = (i - 64 < 4) ? g[1408 + (i - 64)] : 0.0f;
    theta_s[kB1 + i] = v;
  }
}

// Forward pass for the calling warp's 32 samples (all 32 lanes must call it): in[6] / out[4] are the lane's own sample.
template <int TANH = 0>
__device__ __forceinline__ void forward(const float* theta_s, float* scratch, const float (&in)[6], float (&out)[4])
{
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // inputs to fragment layout through the warp's scratch: [half][sample][4] (columns 6, 7 are zero padding)
  float4* s4 = reinterpret_cast<float4*>(scratch);
  s4[lane] = make_float4(in[0], in[1], in[2], in[3]);
  s4[32 + lane] = make_float4(in[4], in[5], 0.0f, 0.0f);
  __syncwarp();
  uint32_t a_hi[2][2], a_lo[2][2];  // layer 1 A fragments (m16n8k8), [m-tile][row half]
  {
    const int off = (t >> 1) * 128 + (t & 1) * 2;  // columns 2t, 2t+1
#pragma unroll
    for (int m = 0; m < 2; m++)
    {
      const float2 top = *reinterpret_cast<const float2*>(scratch + off + (16 * m + g) * 4);
      const float2 bot = *reinterpret_cast<const float2*>(scratch + off + (16 * m + g + 8) * 4);
      split2(top.x, top.y, a_hi[m][0], a_lo[m][0]);
      split2(bot.x, bot.y, a_hi[m][1], a_lo[m][1]);
    }
  }
  __syncwarp();
  const uint2* w1f = reinterpret_cast<const uint2*>(theta_s + kW1F);
  const uint4* w2f = reinterpret_cast<const uint4*>(theta_s + kW2F);
  const uint4* w3f = reinterpret_cast<const uint4*>(theta_s + kW3F);
  // ---- layer 1: 8 (6) -> 32. h = layer 2's A fragments [m-tile][k16-tile][frag]: n-tile i lands in k-tile i / 2, register
  // pair (i & 1) * 2 + {0: rows g, 1: rows g+8}
  uint32_t h_hi[2][2][4], h_lo[2][2][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
  {
    const float2 b = *reinterpret_cast<const float2*>(theta_s + kB1 + 8 * i + 2 * t);
    const uint2 wf = w1f[i * 32 + lane];  // (b0_hi, b0_lo)
#pragma unroll
    for (int m = 0; m < 2; m++)
    {
      float c[4] = { b.x, b.y, b.x, b.y }, x[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
      mma8(x, a_lo[m], wf.x);
      mma8(x, a_hi[m], wf.y);
      mma8(c, a_hi[m], wf.x);
      activate<TANH>(c, x, h_hi[m][i >> 1][(i & 1) * 2], h_hi[m][i >> 1][(i & 1) * 2 + 1], h_lo[m][i >> 1][(i & 1) * 2],
                           h_lo[m][i >> 1][(i & 1) * 2 + 1]);
    }
  }
  // ---- layer 2: 32 -> 32
  uint32_t q_hi[2][2][4], q_lo[2][2][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
  {
    const float2 b = *reinterpret_cast<const float2*>(theta_s + kB2 + 8 * i + 2 * t);
    float c[2][4], x[2][4];
#pragma unroll
    for (int m = 0; m < 2; m++)
    {
      c[m][0] = b.x, c[m][1] = b.y, c[m][2] = b.x, c[m][3] = b.y;
      x[m][0] = x[m][1] = x[m][2] = x[m][3] = 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 2; j++)
    {
      const uint4 wf = w2f[(i * 2 + j) * 32 + lane];  // (b0_hi, b1_hi, b0_lo, b1_lo)
#pragma unroll
      for (int m = 0; m < 2; m++)
      {
        mma16(x[m], h_lo[m][j], wf.x, wf.y);
        mma16(x[m], h_hi[m][j], wf.z, wf.w);
        mma16(c[m], h_hi[m][j], wf.x, wf.y);
      }
    }
#pragma unroll
    for (int m = 0; m < 2; m++)
      activate<TANH>(c[m], x[m], q_hi[m][i >> 1][(i & 1) * 2], q_hi[m][i >> 1][(i & 1) * 2 + 1],
                           q_lo[m][i >> 1][(i & 1) * 2], q_lo[m][i >> 1][(i & 1) * 2 + 1]);
  }
  // ---- layer 3: 32 -> 8 (4)
  float o[2][4];
  {
    const float2 b = *reinterpret_cast<const float2*>(theta_s + kB3 + 2 * t);
    float x[2][4];
#pragma unroll
    for (int m = 0; m < 2; m++)
    {
      o[m][0] = b.x, o[m][1] = b.y, o[m][2] = b.x, o[m][3] = b.y;
      x[m][0] = x[m][1] = x[m][2] = x[m][3] = 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 2; j++)
    {
      const uint4 wf = w3f[j * 32 + lane];
#pragma unroll
      for (int m = 0; m < 2; m++)
      {
        mma16(x[m], q_lo[m][j], wf.x, wf.y);
        mma16(x[m], q_hi[m][j], wf.z, wf.w);
        mma16(o[m], q_hi[m][j], wf.x, wf.y);
      }
    }
#pragma unroll
    for (int m = 0; m < 2; m++)
#pragma unroll
      for (int e = 0; e < 4; e++)
        o[m][e] = fmaf(x[m][e], kLoInv, o[m][e]);
  }
  // outputs back to one sample per lane through the scratch: [sample][4]; columns 4..7 of the tile are padding
  if (t < 2)
  {
#pragma unroll
    for (int m = 0; m < 2; m++)
    {
      *reinterpret_cast<float2*>(scratch + (16 * m + g) * 4 + 2 * t) = make_float2(o[m][0], o[m][1]);
      *reinterpret_cast<float2*>(scratch + (16 * m + g + 8) * 4 + 2 * t) = make_float2(o[m][2], o[m][3]);
    }
  }
  __syncwarp();
  const float4 r = s4[lane];
  out[0] = r.x, out[1] = r.y, out[2] = r.z, out[3] = r.w;
  __syncwarp();
}

}  // namespace nn_mma
}  // namespace mppib
