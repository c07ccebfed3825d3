/*
 * noise_xorwow.cuh — K0, the engine's own control-noise draw: the SAME normals, in the SAME order, as
 * curandGenerateNormal on a CURAND_RNG_PSEUDO_DEFAULT (XORWOW) generator — what the reference's sampler calls
 * (sampling_distributions/gaussian/gaussian.cu:380-381 on the generator of controllers/controller.cu:192-207) — but
 * generated with K-fold more parallelism.
 *
 * cuRAND's host API (default ordering) interleaves 4096 XORWOW subsequences, 2^67 apart: normal PAIR p of the output
 * comes from subsequence p % 4096, Box-Muller round p / 4096 (two 32-bit draws per round). Its kernel therefore runs
 * 4096 threads that each walk N*T*C/8192 rounds serially: 14.5 us for the 0.8 M normals of Cartpole 8192 x 100 and
 * 92 us for Autorally 32768 x 100 on a B200 (tools/curand_probe.cu) — more than the rollout itself.
 *
 * Here every subsequence is additionally split in time into K chunks that run concurrently: 4096*K persistent
 * generator states live in HBM; state (k, j) produces rounds [j*Rc, (j+1)*Rc) of subsequence k for this solve and is
 * then advanced to the same position of the NEXT solve by one fixed jump, a GF(2) matrix-vector product with the
 * precomputed power A^J of the XORWOW transition matrix (4-bit lookup tables in shared memory; the Weyl counter d is
 * advanced arithmetically). The arithmetic on each draw is cuRAND's own device code (_curand_box_muller from
 * <curand_normal.h>), so the output is bit-identical to the library's — asserted in tests/test_gpu_parity.py.
 */
#pragma once
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <stdint.h>

#include <cstring>
#include <vector>

namespace mppib
{
constexpr int kXorwowStreams = 4096;  // cuRAND default-ordering interleave
constexpr int kXorwowNibbles = 40;    // 160 state bits / 4
constexpr uint32_t kXorwowWeyl = 362437u;

// ---- host: transition-matrix powers over GF(2) -------------------------------------------------------------------
struct XorwowVec
{
  uint32_t w[5];
};
inline XorwowVec xorwow_step_linear(XorwowVec s)
{  // curand_kernel.h:863-874 without the Weyl counter (that part is not linear over GF(2))
  uint32_t t = s.w[0] ^ (s.w[0] >> 2);
  XorwowVec r;
  r.w[0] = s.w[1];
  r.w[1] = s.w[2];
  r.w[2] = s.w[3];
  r.w[3] = s.w[4];
  r.w[4] = (s.w[4] ^ (s.w[4] << 4)) ^ (t ^ (t << 1));
  return r;
}
struct XorwowMatrix
{
  XorwowVec col[160];  // col[b] = image of basis vector e_b (bit b%32 of word b/32)
  XorwowVec apply(const XorwowVec& v) const
  {
    XorwowVec r = { { 0, 0, 0, 0, 0 } };
    for (int b = 0; b < 160; b++)
      if ((v.w[b >> 5] >> (b & 31)) & 1u)
        for (int i = 0; i < 5; i++)
          r.w[i] ^= col[b].w[i];
    return r;
  }
  static XorwowMatrix identity()
  {
    XorwowMatrix m;
    memset(&m, 0, sizeof(m));
    for (int b = 0; b < 160; b++)
      m.col[b].w[b >> 5] = 1u << (b & 31);
    return m;
  }
  static XorwowMatrix one_step()
  {
    XorwowMatrix m = identity();
    for (int b = 0; b < 160; b++)
      m.col[b] = xorwow_step_linear(m.col[b]);
    return m;
  }
  XorwowMatrix times(const XorwowMatrix& rhs) const
  {  // (this * rhs): apply rhs first, then this
    XorwowMatrix m;
    for (int b = 0; b < 160; b++)
      m.col[b] = apply(rhs.col[b]);
    return m;
  }
  static XorwowMatrix power(unsigned long long n)
  {
    XorwowMatrix result = identity(), base = one_step();
    while (n)
    {
      if (n & 1ULL)
        result = base.times(result);
      base = base.times(base);
      n >>= 1;
    }
    return result;
  }
};
// 4-bit lookup tables of a matrix: table[nibble][value][word]
inline void xorwow_nibble_tables(const XorwowMatrix& m, std::vector<uint32_t>& out)
{
  out.assign((size_t)kXorwowNibbles * 16 * 5, 0u);
  for (int nb = 0; nb < kXorwowNibbles; nb++)
    for (int val = 0; val < 16; val++)
    {
      uint32_t acc[5] = { 0, 0, 0, 0, 0 };
      for (int bit = 0; bit < 4; bit++)
        if (val & (1 << bit))
          for (int i = 0; i < 5; i++)
            acc[i] ^= m.col[nb * 4 + bit].w[i];
      for (int i = 0; i < 5; i++)
        out[((size_t)nb * 16 + val) * 5 + i] = acc[i];
    }
}

// ---- device ---------------------------------------------------------------------------------------------------------
// state storage: SoA, 6 words per state: v0..v4, d ; index = chunk * 4096 + subsequence
__global__ void xorwow_init_kernel(unsigned long long seed, unsigned long long first_round, int rounds_per_chunk,
                                   int nchunks, uint32_t* __restrict__ states)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nstates = nchunks * kXorwowStreams;
  if (idx >= nstates)
    return;
  const int k = idx % kXorwowStreams, j = idx / kXorwowStreams;
  curandStateXORWOW_t st;
  // subsequence k of the seed, positioned 2 draws per Box-Muller round into it
  curand_init(seed, (unsigned long long)k, 2ULL * (first_round + (unsigned long long)j * rounds_per_chunk), &st);
#pragma unroll
  for (int i = 0; i < 5; i++)
    states[(size_t)i * nstates + idx] = st.v[i];
  states[(size_t)5 * nstates + idx] = st.d;
}

// SPECTRUM = true: the draw is the ColoredNoise sampler's complex spectrum [row][F] (row = sample * C + c) and
// configureFrequencyNoise (colored_noise.cu:12-37: scale by coeffs[c][f], zero the imaginary part of DC / Nyquist) is
// applied on the way out — the separate read-modify-write pass over the 2*N*C*(T+1) floats disappears.
template <bool SPECTRUM>
__global__ void __launch_bounds__(256)
    xorwow_normal_kernel(uint32_t* __restrict__ states, const uint32_t* __restrict__ jump_tables, uint32_t jump_d,
                         int rounds_per_chunk, int nchunks, float2* __restrict__ out /* pairs of this rank's slice */,
                         const float* __restrict__ coeffs = nullptr, int C = 1, int F = 1,
                         unsigned lead = 0u /* window mode: floats of the first round that precede the slice */,
                         unsigned long long count = 0ULL /* window mode (count != 0): floats in the slice */)
{
  __shared__ uint32_t tab[kXorwowNibbles * 16 * 5];
  for (int i = threadIdx.x; i < kXorwowNibbles * 16 * 5; i += blockDim.x)
    tab[i] = jump_tables[i];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nstates = nchunks * kXorwowStreams;
  const bool active = idx < nstates;
  const int k = idx % kXorwowStreams, j = idx / kXorwowStreams;
  uint32_t v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, d = 0;
  if (active)
  {
    v0 = states[idx];
    v1 = states[(size_t)nstates + idx];
    v2 = states[(size_t)2 * nstates + idx];
    v3 = states[(size_t)3 * nstates + idx];
    v4 = states[(size_t)4 * nstates + idx];
    d = states[(size_t)5 * nstates + idx];
    float2* dst = out + ((size_t)j * rounds_per_chunk) * kXorwowStreams + k;
    // position of this thread's first complex number inside the spectrum: (row, f), c = row % C; one round later the
    // index has advanced by 4096
    int f = 0, c = 0;
    int df = 0, dc = 0;
    if (SPECTRUM)
    {
      const size_t p0 = ((size_t)j * rounds_per_chunk) * kXorwowStreams + k;
      const size_t row0 = p0 / (size_t)F;
      f = (int)(p0 - row0 * (size_t)F);
      c = (int)(row0 % (size_t)C);
      df = kXorwowStreams % F;
      dc = (kXorwowStreams / F) % C;
    }
    for (int r = 0; r < rounds_per_chunk; r++)
    {
      // two draws of curand(curandStateXORWOW_t*) (curand_kernel.h:863-874)
      uint32_t t = v0 ^ (v0 >> 2);
      v0 = v1, v1 = v2, v2 = v3, v3 = v4;
      v4 = (v4 ^ (v4 << 4)) ^ (t ^ (t << 1));
      d += kXorwowWeyl;
      const uint32_t x = v4 + d;
      t = v0 ^ (v0 >> 2);
      v0 = v1, v1 = v2, v2 = v3, v3 = v4;
      v4 = (v4 ^ (v4 << 4)) ^ (t ^ (t << 1));
      d += kXorwowWeyl;
      const uint32_t y = v4 + d;
      float2 z = _curand_box_muller(x, y);  // cuRAND's own device arithmetic
      if (SPECTRUM)
      {
        const float v = __ldg(coeffs + c * F + f);
        z.x *= v;
        z.y = (f == 0 || ((F & 1) && f == F - 1)) ? 0.0f : z.y * v;
        f += df;
        c += dc;
        if (f >= F)
        {
          f -= F;
          c += 1;
        }
        if (c >= C)
          c -= C;
      }
      if (count == 0ULL)
        dst[(size_t)r * kXorwowStreams] = z;
      else
      {
        // window mode: the states cover WHOLE 8192-normal rounds around a rank slice that starts / ends inside a round
        // (rank r of W keeps floats [lead, lead + count) of the window); `out` points at the slice, not at the window
        const unsigned long long fi = 2ULL * (((unsigned long long)j * rounds_per_chunk + r) * kXorwowStreams + k);
        float* of = reinterpret_cast<float*>(out);
        if (fi >= lead && fi < lead + count)
          of[fi - lead] = z.x;
        if (fi + 1 >= lead && fi + 1 < lead + count)
          of[fi + 1 - lead] = z.y;
      }
    }
  }
  __syncthreads();
  if (active)
  {
    // jump to the same chunk of the next solve: state <- A^J state, d <- d + 362437 * J
    const uint32_t s[5] = { v0, v1, v2, v3, v4 };
    uint32_t n0 = 0, n1 = 0, n2 = 0, n3 = 0, n4 = 0;
#pragma unroll
    for (int wd = 0; wd < 5; wd++)
#pragma unroll
      for (int q = 0; q < 8; q++)
      {
        const uint32_t* e = tab + ((wd * 8 + q) * 16 + ((s[wd] >> (4 * q)) & 15u)) * 5;
        n0 ^= e[0];
        n1 ^= e[1];
        n2 ^= e[2];
        n3 ^= e[3];
        n4 ^= e[4];
      }
    states[idx] = n0;
    states[(size_t)nstates + idx] = n1;
    states[(size_t)2 * nstates + idx] = n2;
    states[(size_t)3 * nstates + idx] = n3;
    states[(size_t)4 * nstates + idx] = n4;
    states[(size_t)5 * nstates + idx] = d + jump_d;
  }
}

}  // namespace mppib
