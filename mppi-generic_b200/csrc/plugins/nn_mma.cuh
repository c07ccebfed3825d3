/*
 * plugins/nn_mma.cuh — the Autorally 6-32-32-4 network evaluated by a WARP for its 32 / 16 / 8 samples with register-level tensor
 * core MMAs (mma.sync, FP16 inputs, FP32 accumulate); the reference is FNNHelper::forward
 * (include/mppi/utils/nn_helpers/fnn_helper.cu:419-484).
 *
 * Why: the FFMA2 kernel (plugins/dynamics.cuh) is bound by shared-memory wavefronts — every 4 FMAs of a thread need one
 * broadcast LDS.128 of weights, 672 wavefronts per warp-step, 72 % of the data pipe (profiles/r01_autorally_k1_notes.md).
 * Here the weights are B fragments (one LDS.128 per lane per tile carrying the hi and lo parts: 14 loads, ~50 wavefronts per warp-step; the biases as ready-made C fragments, 9 more) and the
 * activations never leave registers between layers: with 16-bit inputs the C fragment of n-tiles 2j, 2j+1 of layer l,
 * packed to half2, IS the A fragment of k-tile j of layer l+1.
 *
 * Precision: a single 11-bit-significand product (FP16 or TF32) is ~4e-4 off per layer and fails the FP32 parity bar of
 * a 100-step recurrence (rollout_kernel_nn_tc.cuh), so every operand is split v = hi + lo with hi = half(v),
 * lo = half(v - hi) and every product is hi*hi + hi*lo + lo*hi, all three in ONE FP32 accumulator that starts at the bias.
 * The residual is not rescaled: for |v| <= 1 it is at worst an FP16 subnormal (spacing 2^-24, i.e. half an FP32 ulp of a
 * value in [0.5, 1)), which the tensor cores take at full rate — this removes the x2048 / (1/2048) multiplies and the second
 * accumulator of the round-1 kernel (8 of 36 instructions per tile). Measured against FP64 (tests/test_nn_mma_scheme.py, CPU
 * emulation; tools/mma_probe.cu on the device): 4e-7 max, plain FP32 FMA chains reach 1.8e-7 on the same inputs.
 * The legacy tensor path issues one mma.sync per ~14 cycles per SM sub-partition whatever the operand type
 * (tools/mma_probe.cu), so FP16 m16n8k16 — twice the k of TF32 m16n8k8 per instruction — is the cheaper encoding: 38 MMAs
 * per 16-row tile and evaluation (layer 1: 8, layer 2: 24, layer 3: 6).
 *
 * Fragment layouts (PTX ISA; g = lane >> 2, t = lane & 3; every register is a half2 of adjacent columns / k):
 *   m16n8k16 A row:  a0 (g, 2t..)  a1 (g+8, 2t..)  a2 (g, 2t+8..)  a3 (g+8, 2t+8..)
 *   m16n8k16 B col:  b0 (k = 2t.., n = g)  b1 (k = 2t+8.., n = g)
 *   m16n8k8  A / B:  a0 a1 / b0 as above
 *   C 16x8 (f32):    c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1)
 * SPW = 32: a warp's samples are two m-tiles (rows 0-15, 16-31), sample s is lane s's; SPW = 16 / 8: one m-tile, see forward().
 */
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mppib
{
namespace nn_mma
{
// shared-memory layout in floats (block-wide part, then scratch per warp)
constexpr int kW1F = 0;               // layer 1: 4 n-tiles x 32 lanes x (b_hi, b_hi, b_lo, 0): the k16 cross-term MMA takes the
                                      // hi part in BOTH of its B registers, stored as a pair so that no MOV builds it  uint4
constexpr int kW2F = kW1F + 4 * 128;  // layer 2: 8 tiles (n-tile i major, k-tile j minor) x 32 x (b0_hi, b1_hi, b0_lo, b1_lo)
constexpr int kW3F = kW2F + 8 * 128;  // layer 3: 2 k-tiles x 32 x (b0_hi, b1_hi, b0_lo, b1_lo)
// biases as ready-made C fragments: [n-tile][t] x (b[2t], b[2t+1], b[2t], b[2t+1]) — one LDS.128 is the accumulator's start
// (rows g and g + 8 share the bias; as a float2 the quad cost two MOVs per tile, 26 of the 314 instructions of a step)
constexpr int kB1 = kW3F + 2 * 128;   // 4 x 4 x 4 (pre-scaled like the weights)
constexpr int kB2 = kB1 + 64;         // 4 x 4 x 4
constexpr int kB3 = kB2 + 64;         // 4 x 4 (columns 4..7 are padding)
constexpr int kFixedFloats = kB3 + 16;  // 1936
// per warp: [2][SPW][4] input halves, then [SPW][4] outputs (separate regions: two __syncwarp per evaluation, not four)
__host__ __device__ constexpr int scratchPerWarp(int spw)
{
  return 12 * spw;
}
__host__ __device__ constexpr int sharedFloats(int block_threads, int spw)
{
  return kFixedFloats + (block_threads / 32) * scratchPerWarp(spw);
}
// tanh(x) = 1 - 2 r with r = 1 / (exp2(2 log2(e) x) + 1). The factor 2 log2(e) is folded into the weights and biases
// FEEDING a tanh; the affine map 1 - 2 r is folded into the weights and biases CONSUMING it (W' = -2 W, b' = b + sum_k W_k),
// so the activation a layer hands on is r itself: one ex2, one add, one rcp per value.
constexpr float kTanhScale = 2.8853900817779268f;

__device__ __forceinline__ uint32_t h2_bits(__half2 h)
{
  return *reinterpret_cast<uint32_t*>(&h);
}
// (v0, v1) -> half2 hi, half2 lo with v = hi + lo to ~2^-24 absolute for |v| <= 1 (the residual of a value below 1/8 is an
// FP16 subnormal: 2^-24 spacing, which the tensor cores take at full rate)
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo)
{
  const __half2 h = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(h);
  hi = h2_bits(h);
  const float2 r = __fadd2_rn(make_float2(v0, v1), make_float2(-hf.x, -hf.y));  // exact (Sterbenz)
  lo = h2_bits(__floats2half2_rn(r.x, r.y));
}
__device__ __forceinline__ void mma16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                      uint32_t b1)
{
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0)
{
  asm("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
// r = 1 / (exp2(z) + 1) for two pre-scaled arguments; ex2 + rcp on the MUFU unit. exp2 overflowing to +inf gives r = 0.
__device__ __forceinline__ float2 sigmoid2_prescaled(float z0, float z1)
{
  float e0, e1, r0, r1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(z0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(z1));
  const float2 d = __fadd2_rn(make_float2(e0, e1), make_float2(1.0f, 1.0f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d.y));
  return make_float2(r0, r1);
}
// Finished tile (bias + three products, one accumulator) -> r -> the half2 registers (rows g and, when BOT, g+8) of the next
// layer's A fragment, hi and lo parts
template <bool BOT>
__device__ __forceinline__ void activate(const float (&c)[4], uint32_t& top_hi, uint32_t& bot_hi, uint32_t& top_lo,
                                         uint32_t& bot_lo)
{
  const float2 u = sigmoid2_prescaled(c[0], c[1]);
  split2(u.x, u.y, top_hi, top_lo);
  if (BOT)
  {
    const float2 v = sigmoid2_prescaled(c[2], c[3]);
    split2(v.x, v.y, bot_hi, bot_lo);
  }
  else
  {
    bot_hi = 0u;
    bot_lo = 0u;
  }
}

// Block-cooperative: reference packed weights (per layer W row-major out x in, then b; fnn_helper.cu:176-183) ->
// fragment-ordered hi / lo half2 parts. Layers 1 and 2 feed a tanh (kTanhScale); layers 2 and 3 consume r = (1 - tanh) / 2
// (W' = -2 W, b' = b + row sum of W, the sum taken in double).
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
      w1f[idx * 4 + 0] = hi;
      w1f[idx * 4 + 1] = hi;
      w1f[idx * 4 + 2] = lo;
      w1f[idx * 4 + 3] = 0u;
    }
    else
    {
      const bool l2 = tile < 12;
      const int q = l2 ? tile - 4 : tile - 12;
      const int i = l2 ? (q >> 1) : 0, j = l2 ? (q & 1) : q;  // n-tile, k16-tile
      const int n = 8 * i + gg;
      const float* W = l2 ? g + 224 : g + 1280;
      const bool real = l2 || gg < 4;  // layer 3 has 4 outputs
      const float sc = l2 ? -2.0f * kTanhScale : -2.0f;
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
    {
      double s = (double)g[1248 + (i - 32)];
      for (int k = 0; k < 32; k++)
        s += (double)g[224 + (i - 32) * 32 + k];
      v = (float)s * kTanhScale;
    }
    else if (i - 64 < 4)
    {
      double s = (double)g[1408 + (i - 64)];
      for (int k = 0; k < 32; k++)
        s += (double)g[1280 + (i - 64) * 32 + k];
      v = (float)s;
    }
    else
      v = 0.0f;
    // bias i of its layer (column 8 n + 2 t + e of n-tile n) -> elements e and e + 2 of quad (n, t)
    const int layer = i < 32 ? 0 : (i < 64 ? 1 : 2), col = i - (layer == 0 ? 0 : (layer == 1 ? 32 : 64));
    float* q = theta_s + (layer == 0 ? kB1 : (layer == 1 ? kB2 : kB3)) + ((col >> 3) * 4 + ((col & 7) >> 1)) * 4 + (col & 1);
    q[0] = v;
    q[2] = v;
  }
}

// The three layers on operands that are already in fragment layout. a_hi / a_lo: layer 1's A fragments per m-tile, [0] =
// rows g, [1] = rows g + 8, columns 2t, 2t+1 of the 8 (6 + 2 padding) inputs. o: layer 3's C fragment per m-tile, (rows g:
// o[0], o[1]; rows g + 8: o[2], o[3]) = output columns 2t, 2t+1 — the 4 real outputs sit on lanes t < 2, lanes t >= 2 hold
// exact zeros (padding rows of W3 and b3). Note the symmetry the warp-specialised rollout uses (rollout_kernel_ar_ws.cuh):
// a lane's slice of the OUTPUT (d/dt of state 3 + 2t, 4 + 2t) is the derivative of its slice of the INPUT.
template <int MT, bool BOT>
__device__ __forceinline__ void forward_frag(const float* theta_s, const uint32_t (&a_hi)[MT][2], const uint32_t (&a_lo)[MT][2],
                                             float (&o)[MT][4])
{
  const int lane = threadIdx.x & 31, t = lane & 3;
  const uint4* w1f = reinterpret_cast<const uint4*>(theta_s + kW1F);
  const uint4* w2f = reinterpret_cast<const uint4*>(theta_s + kW2F);
  const uint4* w3f = reinterpret_cast<const uint4*>(theta_s + kW3F);
  // ---- layer 1: 8 (6) -> 32. The k = 8 operands leave half of a k16 MMA free, so [a_hi | a_lo] x [w_hi ; w_hi] brings
  // a_hi w_hi + a_lo w_hi in one instruction and a k8 MMA adds a_hi w_lo: 2 MMAs per tile.
  // h = layer 2's A fragments [m-tile][k16-tile][frag]: n-tile i lands in k-tile i / 2, registers (i & 1) * 2 + {0: rows g, 1: g+8}
  uint32_t h_hi[MT][2][4], h_lo[MT][2][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
  {
    const float4 b = *reinterpret_cast<const float4*>(theta_s + kB1 + (4 * i + t) * 4);
    const uint4 wf = w1f[i * 32 + lane];  // (b_hi, b_hi, b_lo, -)
#pragma unroll
    for (int m = 0; m < MT; m++)
    {
      float c[4] = { b.x, b.y, b.z, b.w };
      mma16(c, a_hi[m][0], a_hi[m][1], a_lo[m][0], a_lo[m][1], wf.x, wf.y);
      mma8(c, a_hi[m][0], a_hi[m][1], wf.z);
      activate<BOT>(c, h_hi[m][i >> 1][(i & 1) * 2], h_hi[m][i >> 1][(i & 1) * 2 + 1], h_lo[m][i >> 1][(i & 1) * 2],
                    h_lo[m][i >> 1][(i & 1) * 2 + 1]);
    }
  }
  // ---- layer 2: 32 -> 32
  uint32_t q_hi[MT][2][4], q_lo[MT][2][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
  {
    const float4 b = *reinterpret_cast<const float4*>(theta_s + kB2 + (4 * i + t) * 4);
    float c[MT][4];
#pragma unroll
    for (int m = 0; m < MT; m++)
      c[m][0] = b.x, c[m][1] = b.y, c[m][2] = b.z, c[m][3] = b.w;
#pragma unroll
    for (int j = 0; j < 2; j++)
    {
      const uint4 wf = w2f[(i * 2 + j) * 32 + lane];  // (b0_hi, b1_hi, b0_lo, b1_lo)
#pragma unroll
      for (int m = 0; m < MT; m++)
      {
        mma16(c[m], h_hi[m][j][0], h_hi[m][j][1], h_hi[m][j][2], h_hi[m][j][3], wf.x, wf.y);
        mma16(c[m], h_lo[m][j][0], h_lo[m][j][1], h_lo[m][j][2], h_lo[m][j][3], wf.x, wf.y);
        mma16(c[m], h_hi[m][j][0], h_hi[m][j][1], h_hi[m][j][2], h_hi[m][j][3], wf.z, wf.w);
      }
    }
#pragma unroll
    for (int m = 0; m < MT; m++)
      activate<BOT>(c[m], q_hi[m][i >> 1][(i & 1) * 2], q_hi[m][i >> 1][(i & 1) * 2 + 1], q_lo[m][i >> 1][(i & 1) * 2],
                    q_lo[m][i >> 1][(i & 1) * 2 + 1]);
  }
  // ---- layer 3: 32 -> 8 (4); one accumulator per k-tile so that the two chains of three MMAs run side by side
  {
    const float4 b = *reinterpret_cast<const float4*>(theta_s + kB3 + t * 4);
    float o2[MT][4];
#pragma unroll
    for (int m = 0; m < MT; m++)
    {
      o[m][0] = b.x, o[m][1] = b.y, o[m][2] = b.z, o[m][3] = b.w;
      o2[m][0] = o2[m][1] = o2[m][2] = o2[m][3] = 0.0f;
    }
    const uint4 wf0 = w3f[lane], wf1 = w3f[32 + lane];
#pragma unroll
    for (int m = 0; m < MT; m++)
    {
      mma16(o[m], q_hi[m][0][0], q_hi[m][0][1], q_hi[m][0][2], q_hi[m][0][3], wf0.x, wf0.y);
      mma16(o2[m], q_hi[m][1][0], q_hi[m][1][1], q_hi[m][1][2], q_hi[m][1][3], wf1.x, wf1.y);
      mma16(o[m], q_lo[m][0][0], q_lo[m][0][1], q_lo[m][0][2], q_lo[m][0][3], wf0.x, wf0.y);
      mma16(o2[m], q_lo[m][1][0], q_lo[m][1][1], q_lo[m][1][2], q_lo[m][1][3], wf1.x, wf1.y);
      mma16(o[m], q_hi[m][0][0], q_hi[m][0][1], q_hi[m][0][2], q_hi[m][0][3], wf0.z, wf0.w);
      mma16(o2[m], q_hi[m][1][0], q_hi[m][1][1], q_hi[m][1][2], q_hi[m][1][3], wf1.z, wf1.w);
    }
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
      for (int e = 0; e < 4; e++)
        o[m][e] += o2[m][e];
  }
}

// Forward pass for the calling warp's SPW samples (all 32 lanes must call it). SPW = 32: two m16 tiles, lane l owns sample l.
// SPW = 16: one m16 tile, lanes l and l + 16 both carry sample l & 15. SPW = 8: rows 8..15 of the tile are padding (their
// activations are skipped), lanes l, l + 8, l + 16, l + 24 carry sample l & 7.
// in[6] / out[4] are the lane's own sample; `scratch` is the warp's scratchPerWarp(SPW) floats.
template <int SPW>
__device__ __forceinline__ void forward(const float* theta_s, float* scratch, const float (&in)[6], float (&out)[4])
{
  static_assert(SPW == 32 || SPW == 16 || SPW == 8, "samples per warp");
  constexpr int MT = SPW == 32 ? 2 : 1;
  constexpr bool BOT = SPW >= 16;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int srow = lane & (SPW - 1);
  // inputs to fragment layout through the warp's scratch: [half][sample][4] (columns 6, 7 are zero padding); lanes that
  // share a sample store identical values
  float4* s4 = reinterpret_cast<float4*>(scratch);
  s4[srow] = make_float4(in[0], in[1], in[2], in[3]);
  s4[SPW + srow] = make_float4(in[4], in[5], 0.0f, 0.0f);
  __syncwarp();
  uint32_t a_hi[MT][2], a_lo[MT][2];  // layer 1 A fragments, [m-tile][rows g | g+8], columns 2t, 2t+1
  {
    const int off = (t >> 1) * (SPW * 4) + (t & 1) * 2;
#pragma unroll
    for (int m = 0; m < MT; m++)
    {
      const float2 top = *reinterpret_cast<const float2*>(scratch + off + (16 * m + g) * 4);
      split2(top.x, top.y, a_hi[m][0], a_lo[m][0]);
      if (BOT)
      {
        const float2 bot = *reinterpret_cast<const float2*>(scratch + off + (16 * m + g + 8) * 4);
        split2(bot.x, bot.y, a_hi[m][1], a_lo[m][1]);
      }
      else
        a_hi[m][1] = a_lo[m][1] = 0u;
    }
  }
  float o[MT][4];
  forward_frag<MT, BOT>(theta_s, a_hi, a_lo, o);
  // outputs back to one sample per lane through the scratch's output region: [sample][4]; columns 4..7 of the tile are padding
  float* so = scratch + 8 * SPW;
  if (t < 2)
  {
#pragma unroll
    for (int m = 0; m < MT; m++)
    {
      *reinterpret_cast<float2*>(so + (16 * m + g) * 4 + 2 * t) = make_float2(o[m][0], o[m][1]);
      if (BOT)
        *reinterpret_cast<float2*>(so + (16 * m + g + 8) * 4 + 2 * t) = make_float2(o[m][2], o[m][3]);
    }
  }
  __syncwarp();
  const float4 r = reinterpret_cast<const float4*>(so)[srow];
  out[0] = r.x, out[1] = r.y, out[2] = r.z, out[3] = r.w;
}

}  // namespace nn_mma
}  // namespace mppib
