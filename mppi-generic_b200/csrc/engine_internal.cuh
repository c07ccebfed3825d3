/*
 * engine_internal.cuh — the engine's internal types, shared by csrc/engine.cu and by OUT-OF-TREE PLUGIN LIBRARIES.
 *
 * The reference lets a user compile any Dynamics / Cost against its templates (dynamics.cuh:67-76, cost.cuh:34-35,
 * utils/managed.cuh:109-135). Templates cannot cross a C ABI, so here a (dynamics, cost) pair is a REGISTERED kernel
 * instantiation: the built-in pairs are registered by engine.cu; a user pair is compiled into a second shared library from
 * this header — `make_entry<MyDynamics, MyCost>(dyn_id, cost_id)` instantiates K1 (resident / streaming / RMPPI variants),
 * the init-eval and sampled-trajectory kernels for it — and handed to the engine with mppib_register_pair(); see
 * plugins_example/ and INTEGRATION.md §E. The plugin library and libmppi_b200.so must be built from the same source
 * revision (kEngineAbi is checked at registration).
 */
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cufft.h>
#include <curand.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/mppi_b200.h"
#include "combine_kernel.cuh"
#include "plugins/costs.cuh"
#include "plugins/dynamics.cuh"
#include "rollout_kernel.cuh"
#include "rollout_kernel_ar_ws.cuh"
#include "rollout_kernel_nn_tc.cuh"

extern "C" int mppib_set_last_error(int status, const char* fmt, ...);
template <class... A>
static inline int fail(int status, const char* fmt, A... a)
{
  return mppib_set_last_error(status, fmt, a...);
}

#define CUDA_TRY(expr)                                                                                                 \
  do                                                                                                                   \
  {                                                                                                                    \
    cudaError_t _e = (expr);                                                                                           \
    if (_e != cudaSuccess)                                                                                             \
    {                                                                                                                  \
      cudaGetLastError();                                                                                              \
      return fail(MPPIB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);         \
    }                                                                                                                  \
  } while (0)

#define CURAND_TRY(expr)                                                                                               \
  do                                                                                                                   \
  {                                                                                                                    \
    curandStatus_t _s = (expr);                                                                                        \
    if (_s != CURAND_STATUS_SUCCESS)                                                                                   \
      return fail(MPPIB_ERR_CURAND, "%s failed: curandStatus %d (%s:%d)", #expr, (int)_s, __FILE__, __LINE__);         \
  } while (0)

typedef struct ncclComm* ncclComm_t;
using namespace mppib;

// ---- engine state -----------------------------------------------------------------------------------------------
struct mppib_engine
{
  mppib_desc desc{};
  int S = 0, C = 0, O = 0, D = 1;
  int N = 0, T = 0, TC = 0;
  int n_local = 0, n_offset = 0;
  int pstride = 0, nchunks = 0;
  int bx = 64, grid = 0;  // bx = samples (noise-tile rows) per CTA
  int spt = 1;  // samples per thread (rollout_kernel.cuh); threads per CTA = bx / spt * lps
  int lps = 1;  // lanes per sample = 32 / DYN::SAMPLES_PER_WARP (rollout_kernel.cuh: SPW)
  bool stream_k1 = false;  // streaming K1: noise slabs through a ring, controls kept in HBM (rollout_kernel.cuh: STREAM)
  int ring = 2;
  uint32_t smem_bytes = 0;
  bool use_tma = false;
  bool use_pdl = true;
  bool nn_tc = false;  // Autorally pair: NN forward pass on tcgen05 tensor cores (rollout_kernel_nn_tc.cuh)
  bool ar_ws = false;  // Autorally pair: warp-specialised K1 (rollout_kernel_ar_ws.cuh)
  int ws_pspw = 16;    // its samples per producer warp: threads per CTA = bx * (32 / ws_pspw + 1)
  bool mapped_result = true;  // K2 writes the result record straight into mapped pinned host memory
  bool spin_wait = true;      // the host waits for the solve by polling a mapped flag K2's last block sets
  unsigned* k2_counter_d = nullptr;
  volatile unsigned* done_flag_h = nullptr;
  unsigned* done_flag_dev = nullptr;
  unsigned solve_seq = 0;
  bool flag_armed = false;  // the LAST enqueued solve ends in a kernel that publishes done_flag == solve_seq
  bool writeback = false;
  bool rmppi = false;  // MPPIB_FLAG_RMPPI
  float tsallis_gamma = 0.0f, tsallis_r = 0.0f;  // both non-zero: Tsallis weights (mppib_set_tsallis)
  float value_func_threshold = 1000.0f;  // robust_mppi_controller.cuh default
  float* fb_gains_d = nullptr;           // [T][S][C] or null
  float* eval_states_d = nullptr;        // init-eval scratch: candidates, strides, costs
  int* eval_strides_d = nullptr;
  float* eval_costs_d = nullptr;
  int eval_capacity = 0;
  int (*init_eval)(mppib_engine&, const float*, const int*, int, int, const float*, int) = nullptr;
  // sampled (visualisation) trajectories scratch: picked indices, optimised sequence, outputs / costs / crash flags
  int* vis_idx_d = nullptr;
  float* vis_opt_d = nullptr;
  float* nom_d = nullptr;      // device tail: [D][T][C] smoothed controls | [D][T][S] states | [D][T][O] outputs
  float* nom_h = nullptr;      // pinned host copy of the same
  float* nom_u_d = nullptr;    // uploaded [D][T][C] when the caller passes its own U
  int (*nominal_traj)(mppib_engine&, const float*, const float*, int, const float*) = nullptr;
  float* vis_outputs_d = nullptr;
  float* vis_costs_d = nullptr;
  int* vis_crash_d = nullptr;
  int vis_capacity = 0;
  int (*sampled_traj)(mppib_engine&, const float*, const float*, int, int, bool) = nullptr;
  cudaStream_t stream = nullptr;
  bool own_stream = false;

  // solver scalars
  float dt = 0.01f, lambda = 1.0f, alpha = 0.0f;

  // parameter blobs (host copies)
  std::vector<unsigned char> dyn_blob, cost_blob;
  mppib_gaussian_params sampler{};
  bool have_dyn = false, have_cost = false, have_sampler = false;

  // aux device resources
  float* nn_theta_d = nullptr;
  float* lstm_theta_d = nullptr;  // MPPIB_BLOB_LSTM_WEIGHTS
  bool have_lstm = false;
  float* elev_d = nullptr;               // MPPIB_BLOB_ELEVATION_MAP: width * height floats, row-major
  // host copies of the weight / map blobs for mppib_compute_control's host tail (the library's host twins)
  std::vector<float> nn_theta_h, lstm_theta_h;
  std::vector<unsigned char> elev_h;
  size_t elev_capacity = 0;              // floats allocated
  mppib_elevation_map_header elev_hdr{};  // use == 0 until a map is set
  cudaArray_t costmap_array = nullptr;
  cudaTextureObject_t costmap_tex = 0;

  // RNG
  curandGenerator_t gen = nullptr;
  unsigned long long seed = 0;
  unsigned long long rng_offset = 0;  // absolute position (in normals) of the next GLOBAL draw to be CONSUMED
  static constexpr unsigned long long kNoPos = ~0ULL;
  unsigned long long curand_pos = kNoPos;  // global draw position the library generator sits at (world_size == 1 only)
  // own XORWOW draw (noise_xorwow.cuh): 4096 * xw_chunks persistent states
  bool xw_enabled = false;             // sizes allow it and MPPIB_FLAG_CURAND_HOST_API not set
  unsigned long long xw_pos = kNoPos;  // global draw position the states sit at (kNoPos = must be initialised)
  // double-buffered noise: the draw for solve s+1 runs on a side stream while K1/K2 of solve s run (it depends on
  // nothing but the RNG position)
  bool prefetch_enabled = false;
  float* noise_alloc2 = nullptr;
  float* eps_buf[2] = { nullptr, nullptr };
  CUtensorMap tmap_buf[2];
  int cur_buf = 0;
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_k1_done[2] = { nullptr, nullptr };   // K1 that read eps_buf[i] has finished
  cudaEvent_t ev_gen_done[2] = { nullptr, nullptr };  // the draw into eps_buf[i] has finished
  cudaEvent_t ev_last_gen = nullptr;                   // last draw on either stream (generator state ordering)
  bool k1_recorded[2] = { false, false };
  bool any_gen = false;
  bool prefetch_valid = false;
  unsigned long long prefetch_pos = 0;
  int prefetch_buf = 0;
  int xw_chunks = 0, xw_rounds_per_chunk = 0;
  unsigned xw_lead = 0;                     // window mode: floats of the first round before this rank's slice
  unsigned long long xw_first_round = 0;    // first round of the window, relative to the block's position
  bool xw_window = false;                   // the slice starts / ends inside an 8192-normal round
  uint32_t xw_jump_d = 0;
  // normals per generateSamples call: Gaussian N*T*C (gaussian.cu:380-381), ColoredNoise 2*N*C*(T+1) (colored_noise.cu:343)
  unsigned long long draw_global = 0;  // whole job
  unsigned long long draw_start = 0;   // this rank's first normal inside the call's block
  size_t draw_local = 0;               // this rank's normals
  // ColoredNoise sampler (noise_colored.cuh)
  bool colored = false;
  bool nln = false;              // NLN sampler: C log-normal planes + one normal block per draw (nln.cu:114-128)
  float* nln_d = nullptr;        // [C][N][T]
  int F = 0;                     // T + 1 frequencies
  float2* spec_d = nullptr;      // [n_local*C][F] complex spectrum == the raw draw
  float* spec_alloc = nullptr;   // allocation incl. the offset-alignment lead-in
  float* time_d = nullptr;       // [n_local*C][2T] cuFFT output, kept until the next draw
  float* coeffs_d = nullptr;     // [C][F]
  float* sigma_d = nullptr;      // [C]
  float* decay_pow_d = nullptr;  // [T] powf(offset_decay_rate, t)
  cufftHandle fft_plan = 0;
  bool have_plan = false;
  int colored_offset_t = 1;      // optimization_stride assumed by draws issued before a solve names its own
  int buf_offset_t[2] = { 1, 1 };  // stride the colored block in eps_buf[i] was rearranged with
  cudaEvent_t ev_rearr = nullptr;  // last re-rearrange on the main stream (time_d must outlive it)
  bool rearr_recorded = false;
  uint32_t* xw_states_d = nullptr;
  uint32_t* xw_tables_d = nullptr;

  // device buffers
  float* noise_alloc = nullptr;  // allocation incl. lead-in space for offset alignment
  float* eps_d = nullptr;        // [n_local][T][C]
  float* costs_d = nullptr;      // [D][n_local]
  float* partials_d = nullptr;   // [grid][D][pstride]
  float4* headers_d = nullptr;   // [grid][D] compact (beta, eta, sum w^2)
  float4* gather_hdr_d = nullptr;  // [world][D]
  float* controls_d = nullptr;   // optional [D][n_local][T][C]
  float* rank_rec_d = nullptr;   // [D][pstride] this rank's record (world > 1)
  float* gather_d = nullptr;     // [world][D][pstride]
  float* result_d = nullptr;     // [D][pstride] final record (device copy)
  float* result_h = nullptr;     // mapped pinned host copy K2 writes directly
  float* result_h_dev = nullptr; // device alias of result_h
  float* weights_d = nullptr;    // lazily allocated for mppib_get_weights
  unsigned char* l2_flush_d = nullptr;  // optional: buffer written between K0 and K1 to evict the noise from L2
  size_t l2_flush_bytes = 0;
  int pending = 0;               // solves enqueued and not yet waited for
  // accumulated stage timings (timing mode)
  double acc_ms[4] = { 0, 0, 0, 0 };
  long acc_n = 0;

  CUtensorMap tmap{};

  // comm
  ncclComm_t comm = nullptr;
  // peer-memory exchange (combine_kernel.cuh: exchange_merge_kernel)
  bool p2p = false;
  bool p2p_opened = false;
  float* p2p_gather_d = nullptr;   // [2][world][D][pstride] followed by the flag words [2][world]
  PeerTable peers{};
  void* peer_opened[8] = { nullptr };
  unsigned p2p_seq = 0;

  // timing
  bool timing = false;
  cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
  bool timing_valid = false;

  // registry hook
  int (*launch_rollout)(mppib_engine&, const float* x0, const float* U_in, int opt_stride, int iter) = nullptr;
  size_t dyn_param_bytes = 0, cost_param_bytes = 0;
  int dyn_shared_floats = 0;
  int (*dyn_shared_floats_fn)(const int*, int) = nullptr;
  int (*cost_shared_floats)(int) = nullptr;
  int (*prepare)(mppib_engine&) = nullptr;  // sets func attributes
  bool solved_once = false;
};

// ---- registry of (dynamics, cost) pairs compiled into this library ---------------------------------------------
template <class AUX>
struct AuxFill
{
  static void fill(AUX&, const mppib_engine&)
  {
  }
};
template <>
struct AuxFill<plugins::AutorallyNNDynamics::Aux>
{
  static void fill(plugins::AutorallyNNDynamics::Aux& a, const mppib_engine& e)
  {
    a.theta_d = e.nn_theta_d;
  }
};
template <>
struct AuxFill<plugins::RacerLSTMDynamics::Aux>
{
  static void fill(plugins::RacerLSTMDynamics::Aux& a, const mppib_engine& e)
  {
    a.theta_d = e.lstm_theta_d;
    a.H = e.desc.model_dims[0];
    a.L1 = e.desc.model_dims[1];
    a.elev.data = e.elev_d;
    a.elev.hdr = e.elev_hdr;
    if (!e.elev_d)
      a.elev.hdr.use = 0;
  }
};
template <>
struct AuxFill<plugins::ARStandardCost::Aux>
{
  static void fill(plugins::ARStandardCost::Aux& a, const mppib_engine& e)
  {
    a.costmap_tex = e.costmap_tex;
  }
};

template <class DYN, class COST>
struct Pair
{
  using Args = RolloutArgs<DYN, COST>;

  static int cost_shared(int T)
  {
    return COST::sharedFloats(T);
  }

  template <int DD, bool WB, int SPT = 1>
  static int prepare_one(mppib_engine& e)
  {
    CUDA_TRY(cudaFuncSetAttribute(rollout_kernel<DYN, COST, DD, WB, SPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)e.smem_bytes));
    return MPPIB_OK;
  }
  // resident CTAs per SM of the streaming variant (registers, threads and shared memory all count)
  static int stream_blocks_per_sm(int D, int threads, size_t smem)
  {
    int n = 0;
    cudaError_t rc = cudaErrorInvalidValue;
    if (D == 1)
    {
      cudaFuncSetAttribute(rollout_kernel<DYN, COST, 1, true, 1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem);
      rc = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, rollout_kernel<DYN, COST, 1, true, 1, false, true>, threads,
                                                         smem);
    }
    else if constexpr (DYN::MAX_DISTRIBUTIONS >= 2)
    {
      cudaFuncSetAttribute(rollout_kernel<DYN, COST, 2, true, 1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem);
      rc = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, rollout_kernel<DYN, COST, 2, true, 1, false, true>, threads,
                                                         smem);
    }
    if (rc != cudaSuccess)
    {
      cudaGetLastError();
      return 0;
    }
    return n;
  }
  static constexpr bool kHasTensorCoreVariant = std::is_same<DYN, plugins::AutorallyNNDynamics>::value &&
                                                std::is_same<COST, plugins::ARStandardCost>::value;
  static constexpr bool kHasWarpSpecVariant = std::is_same<DYN, plugins::AutorallyNNMmaDynamics<32>>::value &&
                                              std::is_same<COST, plugins::ARStandardCost>::value;
  static int prepare(mppib_engine& e)
  {
    if constexpr (kHasTensorCoreVariant)
    {
      if (e.nn_tc)
      {
        if (e.writeback)
          CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_ar_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)e.smem_bytes));
        else
          CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_ar_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)e.smem_bytes));
        return MPPIB_OK;
      }
    }
    if constexpr (kHasWarpSpecVariant)
    {
      if (e.ar_ws)
      {
        CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_ar_ws<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)e.smem_bytes));
        CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_ar_ws<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)e.smem_bytes));
        CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_ar_ws<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)e.smem_bytes));
        CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_ar_ws<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)e.smem_bytes));
        CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_ar_ws<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)e.smem_bytes));
        CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_ar_ws<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)e.smem_bytes));
        return MPPIB_OK;
      }
    }
    if (e.stream_k1)
    {
      if (e.D == 1)
      {
        CUDA_TRY(cudaFuncSetAttribute(rollout_kernel<DYN, COST, 1, true, 1, false, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e.smem_bytes));
        CUDA_TRY(cudaFuncSetAttribute(rollout_kernel<DYN, COST, 1, false, 1, false, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e.smem_bytes));
        return MPPIB_OK;
      }
      if constexpr (DYN::MAX_DISTRIBUTIONS >= 2)
      {
        CUDA_TRY(cudaFuncSetAttribute(rollout_kernel<DYN, COST, 2, true, 1, false, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e.smem_bytes));
        CUDA_TRY(cudaFuncSetAttribute(rollout_kernel<DYN, COST, 2, false, 1, false, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e.smem_bytes));
        return MPPIB_OK;
      }
      return fail(MPPIB_ERR_UNSUPPORTED, "this dynamics model is built for num_distributions == 1 only");
    }
    if (e.rmppi)
    {
      if constexpr (DYN::MAX_DISTRIBUTIONS >= 2)
      {
        CUDA_TRY(cudaFuncSetAttribute(rollout_kernel<DYN, COST, 2, true, 1, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e.smem_bytes));
        return MPPIB_OK;
      }
      return fail(MPPIB_ERR_UNSUPPORTED, "this dynamics model is built for num_distributions == 1 only");
    }
    if (e.D == 1 && e.spt == 2)
    {
      if constexpr (DYN::MAX_SPT >= 2)
        return e.writeback ? prepare_one<1, true, 2>(e) : prepare_one<1, false, 2>(e);
      return fail(MPPIB_ERR_UNSUPPORTED, "this dynamics model is built for one sample per thread only");
    }
    if (e.D == 1)
      return e.writeback ? prepare_one<1, true>(e) : prepare_one<1, false>(e);
    if constexpr (DYN::MAX_DISTRIBUTIONS >= 2)
      return e.writeback ? prepare_one<2, true>(e) : prepare_one<2, false>(e);
    return fail(MPPIB_ERR_UNSUPPORTED, "this dynamics model is built for num_distributions == 1 only");
  }

  static int launch(mppib_engine& e, const float* x0, const float* U_in, int opt_stride, int iter)
  {
    static_assert(sizeof(Args) < 30000, "kernel parameter block too large");
    Args a;
    memcpy(&a.dyn, e.dyn_blob.data(), sizeof(a.dyn));
    memcpy(&a.cost, e.cost_blob.data(), sizeof(a.cost));
    AuxFill<typename DYN::Aux>::fill(a.dyn_aux, e);
    AuxFill<typename COST::Aux>::fill(a.cost_aux, e);
    const float decay = powf(e.sampler.std_dev_decay, (float)iter);  // gaussian.cu:423
    for (int d = 0; d < MPPIB_MAX_DISTRIBUTIONS; d++)
      for (int c = 0; c < MPPIB_MAX_CONTROL_DIM; c++)
      {
        const float sd = (c < e.C) ? e.sampler.std_dev[d * e.C + c] : 1.0f;
        a.samp.std_dev[d][c] = sd;
        a.samp.std_dev_decayed[d][c] = decay * sd;  // gaussian.cu:86-90
      }
    for (int c = 0; c < MPPIB_MAX_CONTROL_DIM; c++)
      a.samp.control_cost_coeff[c] = e.sampler.control_cost_coeff[c];
    a.samp.pure_noise_threshold = (1.0f - e.sampler.pure_noise_trajectories_percentage) * e.N;  // gaussian.cu:108
    a.eps = e.eps_d;
    a.costs = e.costs_d;
    a.partials = e.partials_d;
    a.headers = e.headers_d;
    a.controls_out = e.writeback ? e.controls_d : nullptr;
    a.n_local = e.n_local;
    a.n_offset = e.n_offset;
    a.T = e.T;
    a.nchunks = e.nchunks;
    a.pstride = e.pstride;
    a.opt_stride = opt_stride;
    a.use_tma = e.use_tma ? 1 : 0;
    a.dyn_shared_floats = e.dyn_shared_floats;
    a.ring = e.stream_k1 ? e.ring : 0;
    a.stream_readback = (e.stream_k1 && e.writeback && getenv("MPPIB_STREAM_READBACK")) ? 1 : 0;
    a.fb_gains = e.fb_gains_d;
    a.value_func_threshold = e.value_func_threshold;
    a.dt = e.dt;
    a.lambda = e.lambda;
    a.alpha = e.alpha;
    a.lambda_inv = (float)(1.0 / e.lambda);  // mppi_controller.cu:201-202: 1.0 / lambda in double, narrowed
    memcpy(a.x0, x0, sizeof(float) * e.D * e.S);
    memcpy(a.means, U_in, sizeof(float) * e.D * e.TC);
    bool launched = false;
    if constexpr (kHasTensorCoreVariant)
    {
      if (e.nn_tc)
      {
        if (e.writeback)
          rollout_kernel_ar_tc<true><<<e.grid, nn_tc::kRows, e.smem_bytes, e.stream>>>(a, e.tmap);
        else
          rollout_kernel_ar_tc<false><<<e.grid, nn_tc::kRows, e.smem_bytes, e.stream>>>(a, e.tmap);
        launched = true;
      }
    }
    if constexpr (kHasWarpSpecVariant)
    {
      if (e.ar_ws)
      {
        const int ws_threads = e.bx * ar_ws::warpsPerGroup(e.ws_pspw);
        if (e.ws_pspw == 16)
        {
          if (e.writeback)
            rollout_kernel_ar_ws<true, 16><<<e.grid, ws_threads, e.smem_bytes, e.stream>>>(a, e.tmap);
          else
            rollout_kernel_ar_ws<false, 16><<<e.grid, ws_threads, e.smem_bytes, e.stream>>>(a, e.tmap);
        }
        else if (e.ws_pspw == 8)
        {
          if (e.writeback)
            rollout_kernel_ar_ws<true, 8><<<e.grid, ws_threads, e.smem_bytes, e.stream>>>(a, e.tmap);
          else
            rollout_kernel_ar_ws<false, 8><<<e.grid, ws_threads, e.smem_bytes, e.stream>>>(a, e.tmap);
        }
        else if (e.writeback)
          rollout_kernel_ar_ws<true, 32><<<e.grid, ws_threads, e.smem_bytes, e.stream>>>(a, e.tmap);
        else
          rollout_kernel_ar_ws<false, 32><<<e.grid, ws_threads, e.smem_bytes, e.stream>>>(a, e.tmap);
        launched = true;
      }
    }
    const int threads = e.bx / e.spt * e.lps;
    if (launched)
    {
    }
    else if (e.stream_k1)
    {
      if (e.D == 1)
      {
        if (e.writeback)
          rollout_kernel<DYN, COST, 1, true, 1, false, true><<<e.grid, threads, e.smem_bytes, e.stream>>>(a, e.tmap);
        else
          rollout_kernel<DYN, COST, 1, false, 1, false, true><<<e.grid, threads, e.smem_bytes, e.stream>>>(a, e.tmap);
      }
      else if constexpr (DYN::MAX_DISTRIBUTIONS >= 2)
      {
        if (e.writeback)
          rollout_kernel<DYN, COST, 2, true, 1, false, true><<<e.grid, threads, e.smem_bytes, e.stream>>>(a, e.tmap);
        else
          rollout_kernel<DYN, COST, 2, false, 1, false, true><<<e.grid, threads, e.smem_bytes, e.stream>>>(a, e.tmap);
      }
    }
    else if (e.rmppi)
    {
      if constexpr (DYN::MAX_DISTRIBUTIONS >= 2)
        rollout_kernel<DYN, COST, 2, true, 1, true><<<e.grid, threads, e.smem_bytes, e.stream>>>(a, e.tmap);
    }
    else if (e.D == 1 && e.spt == 2)
    {
      if constexpr (DYN::MAX_SPT >= 2)
      {
        if (e.writeback)
          rollout_kernel<DYN, COST, 1, true, 2><<<e.grid, threads, e.smem_bytes, e.stream>>>(a, e.tmap);
        else
          rollout_kernel<DYN, COST, 1, false, 2><<<e.grid, threads, e.smem_bytes, e.stream>>>(a, e.tmap);
      }
    }
    else if (e.D == 1)
    {
      if (e.writeback)
        rollout_kernel<DYN, COST, 1, true, 1><<<e.grid, threads, e.smem_bytes, e.stream>>>(a, e.tmap);
      else
        rollout_kernel<DYN, COST, 1, false, 1><<<e.grid, threads, e.smem_bytes, e.stream>>>(a, e.tmap);
    }
    else if constexpr (DYN::MAX_DISTRIBUTIONS >= 2)
    {
      if (e.writeback)
        rollout_kernel<DYN, COST, 2, true, 1><<<e.grid, threads, e.smem_bytes, e.stream>>>(a, e.tmap);
      else
        rollout_kernel<DYN, COST, 2, false, 1><<<e.grid, threads, e.smem_bytes, e.stream>>>(a, e.tmap);
    }
    CUDA_TRY(cudaGetLastError());
    return MPPIB_OK;
  }
};

// launchInitEvalKernel (core/rmppi_kernels.cu:912-937) for this pair
template <class DYN, class COST>
static int init_eval_launch(mppib_engine& e, const float* candidates_d, const int* strides_d, int num_candidates, int samples,
                            const float* U_nominal, int opt_stride)
{
  using Args = InitEvalArgs<DYN, COST>;
  static_assert(sizeof(Args) < 30000, "kernel parameter block too large");
  Args a;
  memcpy(&a.dyn, e.dyn_blob.data(), sizeof(a.dyn));
  memcpy(&a.cost, e.cost_blob.data(), sizeof(a.cost));
  AuxFill<typename DYN::Aux>::fill(a.dyn_aux, e);
  AuxFill<typename COST::Aux>::fill(a.cost_aux, e);
  for (int d = 0; d < MPPIB_MAX_DISTRIBUTIONS; d++)
    for (int c = 0; c < MPPIB_MAX_CONTROL_DIM; c++)
    {
      const float sd = (c < e.C) ? e.sampler.std_dev[d * e.C + c] : 1.0f;
      a.samp.std_dev[d][c] = sd;
      a.samp.std_dev_decayed[d][c] = sd;  // generateSamples(stride, 0, ...): iteration 0, decay^0 = 1
    }
  for (int c = 0; c < MPPIB_MAX_CONTROL_DIM; c++)
    a.samp.control_cost_coeff[c] = e.sampler.control_cost_coeff[c];
  a.samp.pure_noise_threshold = (1.0f - e.sampler.pure_noise_trajectories_percentage) * e.N;
  a.eps = e.eps_d;
  a.candidates = candidates_d;
  a.strides = strides_d;
  a.costs = e.eval_costs_d;
  a.num_candidates = num_candidates;
  a.samples = samples;
  a.T = e.T;
  a.opt_stride = opt_stride;
  const int threads = 64;
  a.dyn_shared_floats = DYN::sharedFloats(e.desc.model_dims, threads);
  a.dt = e.dt;
  a.lambda = e.lambda;
  a.alpha = e.alpha;
  memcpy(a.means, U_nominal, sizeof(float) * e.TC);
  const int total = num_candidates * samples;
  const size_t smem = (size_t)(((a.dyn_shared_floats + 3) / 4) * 4 + COST::sharedFloats(e.T)) * sizeof(float) + 16;
  CUDA_TRY(cudaFuncSetAttribute(init_eval_kernel<DYN, COST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  init_eval_kernel<DYN, COST><<<(total + threads - 1) / threads, threads, smem, e.stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  return MPPIB_OK;
}

// launchVisualizeKernel (core/mppi_common.cu:1376-1420) for this pair: see sampled_traj_kernel
template <class DYN, class COST>
static int sampled_traj_launch(mppib_engine& e, const float* x0, const float* U_nominal, int distribution, int n,
                               bool have_opt)
{
  using Args = SampledTrajArgs<DYN, COST>;
  static_assert(sizeof(Args) < 30000, "kernel parameter block too large");
  Args a;
  memcpy(&a.dyn, e.dyn_blob.data(), sizeof(a.dyn));
  memcpy(&a.cost, e.cost_blob.data(), sizeof(a.cost));
  AuxFill<typename DYN::Aux>::fill(a.dyn_aux, e);
  AuxFill<typename COST::Aux>::fill(a.cost_aux, e);
  for (int d = 0; d < MPPIB_MAX_DISTRIBUTIONS; d++)
    for (int c = 0; c < MPPIB_MAX_CONTROL_DIM; c++)
    {
      const float sd = (c < e.C) ? e.sampler.std_dev[d * e.C + c] : 1.0f;
      a.samp.std_dev[d][c] = sd;
      a.samp.std_dev_decayed[d][c] = sd;
    }
  for (int c = 0; c < MPPIB_MAX_CONTROL_DIM; c++)
    a.samp.control_cost_coeff[c] = e.sampler.control_cost_coeff[c];
  a.samp.pure_noise_threshold = (1.0f - e.sampler.pure_noise_trajectories_percentage) * e.N;
  a.controls = e.controls_d + (size_t)distribution * e.n_local * e.TC;
  a.opt = have_opt ? e.vis_opt_d : nullptr;
  a.sample_idx = e.vis_idx_d;
  a.outputs = e.vis_outputs_d;
  a.costs = e.vis_costs_d;
  a.crash = e.vis_crash_d;
  a.n = n;
  a.T = e.T;
  a.n_offset = e.n_offset;
  a.distribution = distribution;
  const int threads = 64;
  a.dyn_shared_floats = DYN::sharedFloats(e.desc.model_dims, threads);
  a.dt = e.dt;
  a.lambda = e.lambda;
  a.alpha = e.alpha;
  memcpy(a.x0, x0, sizeof(float) * e.S);
  memcpy(a.means, U_nominal, sizeof(float) * e.TC);
  const size_t smem = (size_t)(((a.dyn_shared_floats + 3) / 4) * 4 + COST::sharedFloats(e.T)) * sizeof(float) + 16;
  CUDA_TRY(cudaFuncSetAttribute(sampled_traj_kernel<DYN, COST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sampled_traj_kernel<DYN, COST><<<(n + threads - 1) / threads, threads, smem, e.stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  return MPPIB_OK;
}

// device-side host tail for this pair's dynamics: see nominal_traj_kernel
template <class DYN>
static int nominal_traj_launch(mppib_engine& e, const float* x0, const float* u_src, int u_stride, const float* history)
{
  using Args = NominalTrajArgs<DYN>;
  static_assert(sizeof(Args) < 30000, "kernel parameter block too large");
  Args a;
  memcpy(&a.dyn, e.dyn_blob.data(), sizeof(a.dyn));
  AuxFill<typename DYN::Aux>::fill(a.dyn_aux, e);
  a.u_src = u_src;
  a.u_stride = u_stride;
  a.u_out = e.nom_d;
  a.states = e.nom_d + (size_t)e.D * e.TC;
  a.outputs = a.states + (size_t)e.D * e.T * e.S;
  a.T = e.T;
  a.D = e.D;
  a.smooth = history != nullptr;
  a.dt = e.dt;
  const int threads = 64;
  a.dyn_shared_floats = DYN::sharedFloats(e.desc.model_dims, threads);
  memset(a.x0, 0, sizeof(a.x0));
  memset(a.history, 0, sizeof(a.history));
  for (int d = 0; d < e.D; d++)
    memcpy(a.x0[d], x0 + (size_t)d * e.S, sizeof(float) * e.S);
  if (history)
    for (int k = 0; k < 2; k++)
      memcpy(a.history[k], history + (size_t)k * e.C, sizeof(float) * e.C);
  const size_t smem = (size_t)(((a.dyn_shared_floats + 3) / 4) * 4) * sizeof(float) + 16;
  CUDA_TRY(cudaFuncSetAttribute(nominal_traj_kernel<DYN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nominal_traj_kernel<DYN><<<1, threads, smem, e.stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  return MPPIB_OK;
}

struct PairEntry
{
  int dyn_id, cost_id;
  int S, C, O;
  size_t dyn_bytes, cost_bytes;
  int (*dyn_shared_floats)(const int*, int);
  int max_block_threads;
  int max_spt;
  int spw;  // DYN::SAMPLES_PER_WARP: 32 = one sample per lane; 16 / 8 = sub-warp sample groups (plugins/nn_mma.cuh)
  int (*cost_shared_floats)(int);
  int (*launch)(mppib_engine&, const float*, const float*, int, int);
  int (*prepare)(mppib_engine&);
  int (*init_eval)(mppib_engine&, const float*, const int*, int, int, const float*, int);
  int (*stream_blocks_per_sm)(int, int, size_t);
  int (*sampled_traj)(mppib_engine&, const float*, const float*, int, int, bool);
  int (*nominal_traj)(mppib_engine&, const float*, const float*, int, const float*);
};
template <class DYN, class COST>
constexpr PairEntry make_entry(int dyn_id, int cost_id)
{
  return PairEntry{ dyn_id,
                    cost_id,
                    DYN::STATE_DIM,
                    DYN::CONTROL_DIM,
                    DYN::OUTPUT_DIM,
                    sizeof(typename DYN::Params),
                    sizeof(typename COST::Params),
                    &DYN::sharedFloats,
                    DYN::MAX_BLOCK_THREADS,
                    DYN::MAX_SPT,
                    DYN::SAMPLES_PER_WARP,
                    &Pair<DYN, COST>::cost_shared,
                    &Pair<DYN, COST>::launch,
                    &Pair<DYN, COST>::prepare,
                    &init_eval_launch<typename DYN::AuxDyn, COST>,
                    &Pair<DYN, COST>::stream_blocks_per_sm,
                    &sampled_traj_launch<typename DYN::AuxDyn, COST>,
                    &nominal_traj_launch<typename DYN::AuxDyn> };
}

// ---- registration of out-of-tree pairs -------------------------------------------------------------------------------------
// Layout fingerprint of the internal structs a plugin library shares with libmppi_b200.so: both must come from the same
// source revision.
inline unsigned engine_abi()
{
  return (unsigned)(sizeof(mppib_engine) * 131u + sizeof(PairEntry) * 7u + sizeof(SamplerArgs));
}
extern "C" int mppib_register_pair(const void* pair_entry, size_t entry_bytes, unsigned abi);
// what a plugin's mppib_plugin_init() calls, once per pair: ids >= MPPIB_USER_ID_BASE
template <class DYN, class COST>
inline int register_pair(int dyn_id, int cost_id)
{
  const PairEntry e = make_entry<DYN, COST>(dyn_id, cost_id);
  return mppib_register_pair(&e, sizeof(e), engine_abi());
}
