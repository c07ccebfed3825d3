/*
 * plugins/lstm_mma.cuh — the steering LSTM of RacerDubinsElevationLSTMSteering at hidden_dim = 32 evaluated by a WARP for 16
 * samples with register-level tensor-core MMAs (mma.sync m16n8k16, FP16 hi / lo operands, three products, FP32 accumulate —
 * the scheme of plugins/nn_mma.cuh). Reference: LSTMHelper::forward (include/mppi/utils/nn_helpers/lstm_helper.cu:341-463:
 * gates i, f, o = sigmoid, cell candidate = tanh, c' = i g + f c, h' = o tanh(c'), then the output FNN on [h' ; input],
 * fnn_helper.cu:419-484), called from computeLSTMSteering (racer_dubins_elevation_lstm_steering.cu:131-166).
 *
 * Why: at H = 32 one step of one sample is a [1 x 36] x [36 x 128] gate product plus a {36, L1, 1} head — 5 300 MACs. The
 * one-thread-per-sample form (plugins/dynamics.cuh: lstm_forward) walks them with h and the weights in shared memory:
 * 11.2 ms per C5-sized solve (65536 x 150), 12 % of the FP32 pipe, LSU-bound. As MMAs the same step is 177 instructions per
 * 16 samples, and — the property that makes the recurrence cheap — NOTHING of the LSTM state ever leaves registers:
 *   gate tile (q, k) = n-tile of gate q in {i, f, o, g}, hidden units 8k .. 8k+7. Lane (gg, t) of its C fragment holds
 *   units 8k + 2t, 8k + 2t + 1 of rows gg and gg + 8 — for ALL FOUR gates of the same units in the same lane, so c' and h'
 *   are lane-local; and (h'[8k + 2t], h'[8k + 2t + 1]) packed to half2 IS register (k & 1) * 2 + {0: row gg, 1: row gg + 8}
 *   of k16-tile k >> 1 of next step's A fragment (the trick nn_mma.cuh uses between layers, here between time steps).
 * Only the 4 network inputs and the 1 output per sample cross lanes (64 + 16 floats of per-warp scratch).
 *
 * Activations: sigmoid(v) = 1 / (1 + exp2(-log2(e) v)) and tanh(v) = 1 - 2 / (1 + exp2(2 log2(e) v)), the scale factors
 * folded into the gate weights and biases at load time, so a gate value costs one ex2, one add, one rcp.
 * The head's tanh is handed on as r = 1 / (1 + exp2(z)) with W2' = -2 W2, b2' = b2 + sum W2 (nn_mma.cuh).
 *
 * Restricted to H == 32 and head width L1 <= 24 (three n-tiles); other sizes keep the one-thread-per-sample form.
 */
#pragma once
#include "nn_mma.cuh"

namespace mppib
{
namespace lstm_mma
{
constexpr int H = 32, I = 4;          // hidden units, network inputs (MPPIB_RACER_LSTM_INPUT_DIM)
constexpr int kGateTiles = 16;        // 4 gates x 4 unit groups
constexpr int kKT = 3;                // k16 tiles of [h (32) ; input (4) ; padding (12)]
constexpr int kHeadTiles = 3;         // head layer 1: up to 24 neurons
// shared-memory layout in floats (uint4 fragments = 4 floats each)
constexpr int kWG = 0;                                   // gates: [tile][kt][lane] x (b0_hi, b1_hi, b0_lo, b1_lo)
constexpr int kW1 = kWG + kGateTiles * kKT * 32 * 4;     // head layer 1: [tile][kt][lane] x uint4
constexpr int kW2 = kW1 + kHeadTiles * kKT * 32 * 4;     // head layer 2: [kt 0..1][lane] x uint4 (24 -> 1, padded to k 32, n 8)
constexpr int kBG = kW2 + 2 * 32 * 4;                    // gate biases [tile][8], pre-scaled
constexpr int kB1 = kBG + kGateTiles * 8;                // head biases [24], pre-scaled
constexpr int kB2 = kB1 + kHeadTiles * 8;                // head output bias (+ row sum), 8 floats (1 real)
constexpr int kFixedFloats = kB2 + 8;
constexpr int kScratchPerWarp = 16 * 4 + 16;             // inputs [16][4], outputs [16]
__host__ __device__ constexpr int sharedFloats(int warps)
{
  return kFixedFloats + warps * kScratchPerWarp;
}
constexpr float kLog2e = 1.4426950408889634f;

// per-warp recurrent state in fragment layout (16 samples: rows gg and gg + 8 of one m16 tile)
struct State
{
  uint32_t h_hi[2][4], h_lo[2][4];  // A fragments of k16-tiles 0, 1 (hidden units 0-15, 16-31)
  float c[4][4];                    // cell: [unit group k][(row gg: 2t, 2t+1), (row gg+8: 2t, 2t+1)]
};

// Block-cooperative. g = the reference's packed blob: W_im W_fm W_om W_cm [H x H] | W_ii W_fi W_oi W_ci [H x I] |
// b_i b_f b_o b_c [H] | initial hidden, cell [H] | head W1 [L1 x (H+I)] b1 [L1] W2 [1 x L1] b2 (lstm_helper.cu:72-88,
// fnn_helper.cu:176-183). Gate order of our tiles: 0 = i, 1 = f, 2 = o, 3 = cell candidate.
__device__ __forceinline__ void load_weights(const float* __restrict__ g, int L1, float* theta_s)
{
  constexpr int HH = H * H, IH = H * I;
  const float* gb = g + 4 * HH + 4 * IH;
  const float* hd = gb + 6 * H;
  const int IN = H + I;
  uint32_t* wg = reinterpret_cast<uint32_t*>(theta_s + kWG);
  uint32_t* w1 = reinterpret_cast<uint32_t*>(theta_s + kW1);
  uint32_t* w2 = reinterpret_cast<uint32_t*>(theta_s + kW2);
  // column k (0..47) of the concatenated operand [h ; input ; 0] for gate q, unit n
  auto gate_w = [&](int q, int n, int k) -> float {
    const float sc = (q == 3) ? 2.0f * kLog2e : -kLog2e;  // tanh / sigmoid pre-scale
    if (k < H)
      return g[q * HH + n * H + k] * sc;
    if (k < H + I)
      return g[4 * HH + q * IH + n * I + (k - H)] * sc;
    return 0.0f;
  };
  for (int idx = threadIdx.x; idx < kGateTiles * kKT * 32; idx += blockDim.x)
  {
    const int lane = idx & 31, kt = (idx >> 5) % kKT, tile = (idx >> 5) / kKT;
    const int gg = lane >> 2, t = lane & 3, q = tile >> 2, grp = tile & 3;
    const int n = 8 * grp + gg, k0 = 16 * kt + 2 * t;
    uint32_t h0, l0, h1, l1;
    nn_mma::split2(gate_w(q, n, k0), gate_w(q, n, k0 + 1), h0, l0);
    nn_mma::split2(gate_w(q, n, k0 + 8), gate_w(q, n, k0 + 9), h1, l1);
    uint32_t* d = wg + idx * 4;
    d[0] = h0, d[1] = h1, d[2] = l0, d[3] = l1;
  }
  auto head_w = [&](int n, int k) -> float {  // layer 1 feeds a tanh
    return (n < L1 && k < IN) ? hd[n * IN + k] * (2.0f * kLog2e) : 0.0f;
  };
  for (int idx = threadIdx.x; idx < kHeadTiles * kKT * 32; idx += blockDim.x)
  {
    const int lane = idx & 31, kt = (idx >> 5) % kKT, tile = (idx >> 5) / kKT;
    const int gg = lane >> 2, t = lane & 3;
    const int n = 8 * tile + gg, k0 = 16 * kt + 2 * t;
    uint32_t h0, l0, h1, l1;
    nn_mma::split2(head_w(n, k0), head_w(n, k0 + 1), h0, l0);
    nn_mma::split2(head_w(n, k0 + 8), head_w(n, k0 + 9), h1, l1);
    uint32_t* d = w1 + idx * 4;
    d[0] = h0, d[1] = h1, d[2] = l0, d[3] = l1;
  }
  const float* W2 = hd + L1 * IN + L1;
  auto out_w = [&](int n, int k) -> float {  // layer 2 consumes r = (1 - tanh) / 2: W' = -2 W (output column 0 only)
    return (n == 0 && k < L1) ? -2.0f * W2[k] : 0.0f;
  };
  for (int idx = threadIdx.x; idx < 2 * 32; idx += blockDim.x)
  {
    const int lane = idx & 31, kt = idx >> 5;
    const int gg = lane >> 2, t = lane & 3, k0 = 16 * kt + 2 * t;
    uint32_t h0, l0, h1, l1;
    nn_mma::split2(out_w(gg, k0), out_w(gg, k0 + 1), h0, l0);
    nn_mma::split2(out_w(gg, k0 + 8), out_w(gg, k0 + 9), h1, l1);
    uint32_t* d = w2 + idx * 4;
    d[0] = h0, d[1] = h1, d[2] = l0, d[3] = l1;
  }
  for (int i = threadIdx.x; i < kGateTiles * 8; i += blockDim.x)
  {
    const int tile = i >> 3, q = tile >> 2, n = 8 * (tile & 3) + (i & 7);
    theta_s[kBG + i] = gb[q * H + n] * ((q == 3) ? 2.0f * kLog2e : -kLog2e);
  }
  for (int i = threadIdx.x; i < kHeadTiles * 8; i += blockDim.x)
    theta_s[kB1 + i] = (i < L1) ? hd[L1 * IN + i] * (2.0f * kLog2e) : 0.0f;
  if (threadIdx.x < 8)
  {
    double s = 0.0;
    if (threadIdx.x == 0)
    {
      s = (double)W2[L1];  // b2
      for (int k = 0; k < L1; k++)
        s += (double)W2[k];
    }
    theta_s[kB2 + threadIdx.x] = (float)s;
  }
}

// initial hidden / cell state (identical for every sample) into fragment layout
__device__ __forceinline__ void init_state(const float* __restrict__ g, State& s)
{
  const float* init = g + 4 * H * H + 4 * H * I + 4 * H;
  const int t = threadIdx.x & 3;
#pragma unroll
  for (int k = 0; k < 4; k++)
  {
    const float h0 = init[8 * k + 2 * t], h1 = init[8 * k + 2 * t + 1];
    uint32_t hi, lo;
    nn_mma::split2(h0, h1, hi, lo);
    s.h_hi[k >> 1][(k & 1) * 2] = s.h_hi[k >> 1][(k & 1) * 2 + 1] = hi;
    s.h_lo[k >> 1][(k & 1) * 2] = s.h_lo[k >> 1][(k & 1) * 2 + 1] = lo;
    const float c0 = init[H + 8 * k + 2 * t], c1 = init[H + 8 * k + 2 * t + 1];
    s.c[k][0] = s.c[k][2] = c0;
    s.c[k][1] = s.c[k][3] = c1;
  }
}

// three products of one A fragment with one B fragment pair into one accumulator
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&a_hi)[4], const uint32_t (&a_lo)[4], const uint4& w)
{
  nn_mma::mma16(c, a_hi[0], a_hi[1], a_hi[2], a_hi[3], w.x, w.y);
  nn_mma::mma16(c, a_lo[0], a_lo[1], a_lo[2], a_lo[3], w.x, w.y);
  nn_mma::mma16(c, a_hi[0], a_hi[1], a_hi[2], a_hi[3], w.z, w.w);
}

// One LSTM step + head for the calling warp's 16 samples (all 32 lanes must call it; lanes l and l + 16 carry sample l & 15
// and pass identical `in`). Returns the head's output for the lane's sample and advances `s`.
__device__ __forceinline__ float forward(const float* theta_s, float* scratch, const float (&in)[I], State& s)
{
  const int lane = threadIdx.x & 31, gg = lane >> 2, t = lane & 3, srow = lane & 15;
  // inputs to fragment layout: k16-tile 2, columns 32 + 2t, 33 + 2t (t < 2), everything else of the tile is zero padding
  float4* s4 = reinterpret_cast<float4*>(scratch);
  s4[srow] = make_float4(in[0], in[1], in[2], in[3]);
  __syncwarp();
  uint32_t x_hi[4] = { 0u, 0u, 0u, 0u }, x_lo[4] = { 0u, 0u, 0u, 0u };
  if (t < 2)
  {
    const float2 top = *reinterpret_cast<const float2*>(scratch + gg * 4 + 2 * t);
    const float2 bot = *reinterpret_cast<const float2*>(scratch + (gg + 8) * 4 + 2 * t);
    nn_mma::split2(top.x, top.y, x_hi[0], x_lo[0]);
    nn_mma::split2(bot.x, bot.y, x_hi[1], x_lo[1]);
  }
  const uint4* wg = reinterpret_cast<const uint4*>(theta_s + kWG);
  const uint4* w1 = reinterpret_cast<const uint4*>(theta_s + kW1);
  const uint4* w2 = reinterpret_cast<const uint4*>(theta_s + kW2);
  // ---- gates, one group of 8 hidden units at a time: tiles (i, f, o, g) x k-tiles (h 0-15, h 16-31, input) ------------
  uint32_t n_hi[2][4], n_lo[2][4];  // h' in next step's A-fragment layout
#pragma unroll
  for (int k = 0; k < 4; k++)
  {
    float acc[4][4];
#pragma unroll
    for (int q = 0; q < 4; q++)
    {
      const float2 b = *reinterpret_cast<const float2*>(theta_s + kBG + (q * 4 + k) * 8 + 2 * t);
      acc[q][0] = b.x, acc[q][1] = b.y, acc[q][2] = b.x, acc[q][3] = b.y;
    }
#pragma unroll
    for (int q = 0; q < 4; q++)
    {
      const int tile = q * 4 + k;
      mma3(acc[q], s.h_hi[0], s.h_lo[0], wg[(tile * kKT + 0) * 32 + lane]);
      mma3(acc[q], s.h_hi[1], s.h_lo[1], wg[(tile * kKT + 1) * 32 + lane]);
      mma3(acc[q], x_hi, x_lo, wg[(tile * kKT + 2) * 32 + lane]);
    }
    // i, f, o = sigmoid: r = 1 / (1 + exp2(z)) with z = -log2(e) v; g = tanh: 1 - 2 r with z = 2 log2(e) v
    float hn[4];
#pragma unroll
    for (int e = 0; e < 4; e += 2)
    {
      const float2 gi = nn_mma::sigmoid2_prescaled(acc[0][e], acc[0][e + 1]);
      const float2 gf = nn_mma::sigmoid2_prescaled(acc[1][e], acc[1][e + 1]);
      const float2 go = nn_mma::sigmoid2_prescaled(acc[2][e], acc[2][e + 1]);
      const float2 rc = nn_mma::sigmoid2_prescaled(acc[3][e], acc[3][e + 1]);
      const float g0 = fmaf(-2.0f, rc.x, 1.0f), g1 = fmaf(-2.0f, rc.y, 1.0f);
      const float c0 = fmaf(gi.x, g0, gf.x * s.c[k][e]), c1 = fmaf(gi.y, g1, gf.y * s.c[k][e + 1]);  // c' = i g + f c
      s.c[k][e] = c0, s.c[k][e + 1] = c1;
      const float2 rt = nn_mma::sigmoid2_prescaled(c0 * (2.0f * kLog2e), c1 * (2.0f * kLog2e));
      hn[e] = go.x * fmaf(-2.0f, rt.x, 1.0f);  // h' = o tanh(c')
      hn[e + 1] = go.y * fmaf(-2.0f, rt.y, 1.0f);
    }
    nn_mma::split2(hn[0], hn[1], n_hi[k >> 1][(k & 1) * 2], n_lo[k >> 1][(k & 1) * 2]);
    nn_mma::split2(hn[2], hn[3], n_hi[k >> 1][(k & 1) * 2 + 1], n_lo[k >> 1][(k & 1) * 2 + 1]);
  }
#pragma unroll
  for (int j = 0; j < 2; j++)
#pragma unroll
    for (int r = 0; r < 4; r++)
      s.h_hi[j][r] = n_hi[j][r], s.h_lo[j][r] = n_lo[j][r];
  // ---- head layer 1 on [h' ; input]: 3 n-tiles, activation handed on as r ----------------------------------------------
  uint32_t a_hi[2][4], a_lo[2][4];  // layer 2's A fragments: neurons 0-15 (k-tile 0), 16-23 + padding (k-tile 1)
  a_hi[1][2] = a_hi[1][3] = a_lo[1][2] = a_lo[1][3] = 0u;
#pragma unroll
  for (int i = 0; i < kHeadTiles; i++)
  {
    const float2 b = *reinterpret_cast<const float2*>(theta_s + kB1 + 8 * i + 2 * t);
    float c[4] = { b.x, b.y, b.x, b.y };
    mma3(c, s.h_hi[0], s.h_lo[0], w1[(i * kKT + 0) * 32 + lane]);
    mma3(c, s.h_hi[1], s.h_lo[1], w1[(i * kKT + 1) * 32 + lane]);
    mma3(c, x_hi, x_lo, w1[(i * kKT + 2) * 32 + lane]);
    nn_mma::activate<true>(c, a_hi[i >> 1][(i & 1) * 2], a_hi[i >> 1][(i & 1) * 2 + 1], a_lo[i >> 1][(i & 1) * 2],
                           a_lo[i >> 1][(i & 1) * 2 + 1]);
  }
  // ---- head layer 2: 24 -> 1 (output column 0 of an n8 tile) -----------------------------------------------------------
  float o[4], o2[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
  {
    const float2 b = *reinterpret_cast<const float2*>(theta_s + kB2 + 2 * t);
    o[0] = b.x, o[1] = b.y, o[2] = b.x, o[3] = b.y;
    mma3(o, a_hi[0], a_lo[0], w2[lane]);
    mma3(o2, a_hi[1], a_lo[1], w2[32 + lane]);
  }
  // padding neurons (n >= L1) contribute r = 0.5 times a zero weight; output rows gg (o[0]) and gg + 8 (o[2]) on lanes t == 0
  float* so = scratch + 16 * 4;
  if (t == 0)
  {
    so[gg] = o[0] + o2[0];
    so[gg + 8] = o[2] + o2[2];
  }
  __syncwarp();
  return so[srow];
}

}  // namespace lstm_mma
}  // namespace mppib
