/*
 * engine.cu — the C-ABI (include/mppi_b200.h) and the host-side orchestration of one MPPI solve on one B200:
 *   K0 noise draw (cuRAND XORWOW, same generator / seed / offset / count as controllers/controller.cu:192-207 and
 *      sampling_distributions/gaussian/gaussian.cu:380-381, so sample indexing is bit-identical to the reference)
 *   K1 fused rollout                (rollout_kernel.cuh)
 *   K2 baseline / weights / average (combine_kernel.cuh)  [+ one NCCL all-gather and a second K2 when world_size > 1]
 * One stream, no host round trip between the kernels (the reference synchronises three times per iteration,
 * controllers/MPPI/mppi_controller.cu:187-218); x0 and the nominal controls travel in the kernel parameter bank and
 * the result record is written by K2 straight into mapped pinned host memory, so a solve issues no cudaMemcpy.
 *
 * The engine never computes on the CPU: without a CUDA device mppib_create fails with MPPIB_ERR_NO_DEVICE.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <cufft.h>
#include <curand.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <type_traits>
#include <mutex>
#include <vector>

#include "../../include/mppi_b200.h"
#include "combine_kernel.cuh"
#include "noise_colored.cuh"
#include "noise_xorwow.cuh"
#include "plugins/costs.cuh"
#include "plugins/dynamics.cuh"
#include "rollout_kernel.cuh"
#include "rollout_kernel_ar_ws.cuh"
#include "rollout_kernel_nn_tc.cuh"
#include "engine_internal.cuh"
#include "../../include/mppi_b200/host_twins.h"

namespace
{
// ---- minimal NCCL binding, resolved lazily with dlopen so single-GPU users need no NCCL at all --------------------
typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId
{
  char internal[128];
};
struct NcclApi
{
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int /*ncclDataType_t*/, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load()
  {
    if (handle)
      return true;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char* n : names)
    {
      handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (handle)
        break;
    }
    if (!handle)
      return false;
    GetUniqueId = (decltype(GetUniqueId))dlsym(handle, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(handle, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
    AllGather = (decltype(AllGather))dlsym(handle, "ncclAllGather");
    GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && CommDestroy && AllGather;
  }
};
NcclApi g_nccl;
constexpr int kNcclFloat = 7;  // ncclFloat32

}  // namespace


// The library's error channel: a thread-local message behind mppib_last_error(). Exported (not part of the public header):
// the other translation units (npz_reader.cpp) and out-of-tree plugin libraries (engine_internal.cuh: fail()) report
// through it.
static thread_local std::string g_last_error;
extern "C" int mppib_set_last_error(int status, const char* fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return status;
}

using namespace mppib;

static const PairEntry kPairs[] = {
  make_entry<plugins::CartpoleDynamics, plugins::CartpoleQuadraticCost>(MPPIB_DYN_CARTPOLE,
                                                                        MPPIB_COST_CARTPOLE_QUADRATIC),
  make_entry<plugins::DoubleIntegratorDynamics, plugins::DoubleIntegratorCircleCost>(MPPIB_DYN_DOUBLE_INTEGRATOR,
                                                                                     MPPIB_COST_DI_CIRCLE),
  make_entry<plugins::AutorallyNNDynamics, plugins::ARStandardCost>(MPPIB_DYN_AUTORALLY_NN, MPPIB_COST_AR_STANDARD),
  make_entry<plugins::RacerLSTMDynamics, plugins::RacerQuadraticCost>(MPPIB_DYN_RACER_LSTM, MPPIB_COST_RACER_QUADRATIC),
  make_entry<plugins::QuadrotorDynamics, plugins::QuadrotorQuadraticCost>(MPPIB_DYN_QUADROTOR,
                                                                          MPPIB_COST_QUADROTOR_QUADRATIC),
};

// RacerDubinsElevationLSTMSteering with the steering LSTM on tensor cores (hidden_dim 32, head width <= 24)
static const PairEntry kPairsLstmMma[] = {
  make_entry<plugins::RacerLSTMMmaDynamics, plugins::RacerQuadraticCost>(MPPIB_DYN_RACER_LSTM, MPPIB_COST_RACER_QUADRATIC),
};
constexpr int kWsPspw8MaxRollouts = 6144;   // see mppib_create (warp-specialised Autorally K1)
// the Autorally pair's default form, one entry per samples-per-warp width (chosen at create time from n_local)
static const PairEntry kPairsMma[] = {
  make_entry<plugins::AutorallyNNMmaDynamics<32>, plugins::ARStandardCost>(MPPIB_DYN_AUTORALLY_NN, MPPIB_COST_AR_STANDARD),
  make_entry<plugins::AutorallyNNMmaDynamics<16>, plugins::ARStandardCost>(MPPIB_DYN_AUTORALLY_NN, MPPIB_COST_AR_STANDARD),
  make_entry<plugins::AutorallyNNMmaDynamics<8>, plugins::ARStandardCost>(MPPIB_DYN_AUTORALLY_NN, MPPIB_COST_AR_STANDARD),
};

// pairs registered by plugin libraries (mppib_register_pair); std::vector grows, so entries are kept by pointer
static std::vector<PairEntry*>& user_pairs()
{
  static std::vector<PairEntry*> v;
  return v;
}

// ---- helpers ----------------------------------------------------------------------------------------------------
static int make_tensor_map(mppib_engine& e, float* base, CUtensorMap* out)
{
  // 2-D view of the noise buffer: rows = local rollouts, cols = T*C floats (row pitch T*C*4 B, must be 16-B multiple)
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess)
    return fail(MPPIB_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = { (cuuint64_t)e.TC, (cuuint64_t)e.n_local };
  cuuint64_t gstride[1] = { (cuuint64_t)e.TC * sizeof(float) };
  cuuint32_t box[2] = { (cuuint32_t)kChunkFloats, (cuuint32_t)e.bx };
  cuuint32_t estride[2] = { 1, 1 };
  CUresult r = ((EncodeFn)fn)(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estride,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(MPPIB_ERR_CUDA, "cuTensorMapEncodeTiled failed: CUresult %d", (int)r);
  return MPPIB_OK;
}

// ColoredNoise: rearrangeNoise (colored_noise.cu:39-56) from the retained time-domain buffer into eps_buf[buf] on `st`.
static int colored_rearrange(mppib_engine& e, int buf, cudaStream_t st, int offset_t)
{
  if (offset_t < 0 || offset_t >= 2 * e.T)
    return fail(MPPIB_ERR_INVALID_ARG, "optimization_stride %d outside the 2T = %d colored-noise samples", offset_t,
                2 * e.T);
  const size_t total = (size_t)e.n_local * e.T;
  const dim3 block(256), grid((unsigned)((total + 255) / 256));
  float* dst = e.eps_buf[buf];
  switch (e.C)
  {
    case 1:
      colored_rearrange_kernel<1><<<grid, block, 0, st>>>(e.time_d, dst, e.sigma_d, e.decay_pow_d, e.n_local, e.T, offset_t);
      break;
    case 2:
      colored_rearrange_kernel<2><<<grid, block, 0, st>>>(e.time_d, dst, e.sigma_d, e.decay_pow_d, e.n_local, e.T, offset_t);
      break;
    case 4:
      colored_rearrange_kernel<4><<<grid, block, 0, st>>>(e.time_d, dst, e.sigma_d, e.decay_pow_d, e.n_local, e.T, offset_t);
      break;
    default:
      return fail(MPPIB_ERR_UNSUPPORTED, "ColoredNoise: CONTROL_DIM %d", e.C);
  }
  CUDA_TRY(cudaGetLastError());
  e.buf_offset_t[buf] = offset_t;
  return MPPIB_OK;
}

// One generateSamples-equivalent draw (gaussian.cu:378-394 / colored_noise.cu:343-372) of the block of the global
// XORWOW stream that starts at global position `pos` (draw_global normals per block; this rank keeps elements
// [draw_start, draw_start + draw_local) of it) into eps_buf[buf], on `st`. Draws are totally ordered through
// ev_last_gen because they share the generator state (and, for ColoredNoise, the spectrum / time buffers).
static int gen_draw(mppib_engine& e, int buf, cudaStream_t st, unsigned long long pos, int offset_t)
{
  const unsigned long long global_count = e.draw_global;
  const unsigned long long start = pos + e.draw_start;
  const size_t count = e.draw_local;
  float* dst = e.colored ? reinterpret_cast<float*>(e.spec_d) : e.eps_buf[buf];
  bool scaled_in_draw = false;  // the engine's own generator applies configureFrequencyNoise on the way out
  if (e.any_gen)
    CUDA_TRY(cudaStreamWaitEvent(st, e.ev_last_gen, 0));
  if (e.colored && e.rearr_recorded)
    CUDA_TRY(cudaStreamWaitEvent(st, e.ev_rearr, 0));  // a re-rearrange may still be reading time_d
  if (e.nln)
  {
    // NLNDistribution::generateSamples (nln.cu:114-128): C curandGenerateLogNormal calls of N*T values (mean 0, std dev
    // sigma_c) into plane c, one curandGenerateNormal of N*T*C, then createNLNNoise. Library generator, call after call like
    // the reference; re-positioning (seed / burn) is exact only where cuRAND honours absolute offsets (multiples of 8192,
    // tools/curand_probe.cu) and every call then stays on such a boundary.
    const size_t plane = (size_t)e.n_local * e.T;
    if (plane & 1)
      return fail(MPPIB_ERR_UNSUPPORTED, "cuRAND draws need an even count (N * T = %zu)", plane);
    CURAND_TRY(curandSetStream(e.gen, st));
    if (e.curand_pos != pos)
    {
      if ((pos % 8192ULL) != 0 || (plane % 8192) != 0)
        return fail(MPPIB_ERR_UNSUPPORTED, "NLN draws continue the generator call after call; re-positioning it to "
                                           "offset %llu needs N * T (= %zu) to be a multiple of 8192", pos, plane);
      CURAND_TRY(curandSetGeneratorOffset(e.gen, pos));
    }
    for (int c = 0; c < e.C; c++)
      CURAND_TRY(curandGenerateLogNormal(e.gen, e.nln_d + (size_t)c * plane, plane, 0.0f, e.sampler.std_dev[c]));
    CURAND_TRY(curandGenerateNormal(e.gen, dst, plane * e.C, 0.0f, 1.0f));
    const int blocks = (int)std::min<size_t>((plane + 255) / 256, 148 * 16);
    switch (e.C)
    {
      case 1:
        nln_combine_kernel<1><<<blocks, 256, 0, st>>>(dst, e.nln_d, e.n_local, e.T);
        break;
      case 2:
        nln_combine_kernel<2><<<blocks, 256, 0, st>>>(dst, e.nln_d, e.n_local, e.T);
        break;
      default:
        nln_combine_kernel<4><<<blocks, 256, 0, st>>>(dst, e.nln_d, e.n_local, e.T);
        break;
    }
    CUDA_TRY(cudaGetLastError());
    e.curand_pos = pos + global_count;
  }
  else if (e.xw_enabled && (pos % 8192ULL) == 0)
  {
    const int nstates = e.xw_chunks * kXorwowStreams;
    if (e.xw_pos != pos)
    {
      xorwow_init_kernel<<<(nstates + 127) / 128, 128, 0, st>>>(e.seed, pos / 8192ULL + e.xw_first_round,
                                                             e.xw_rounds_per_chunk, e.xw_chunks, e.xw_states_d);
      CUDA_TRY(cudaGetLastError());
    }
    if (e.colored)
      xorwow_normal_kernel<true><<<(nstates + 255) / 256, 256, 0, st>>>(e.xw_states_d, e.xw_tables_d, e.xw_jump_d,
                                                                       e.xw_rounds_per_chunk, e.xw_chunks,
                                                                       reinterpret_cast<float2*>(dst), e.coeffs_d, e.C, e.F);
    else
      xorwow_normal_kernel<false><<<(nstates + 255) / 256, 256, 0, st>>>(
          e.xw_states_d, e.xw_tables_d, e.xw_jump_d, e.xw_rounds_per_chunk, e.xw_chunks, reinterpret_cast<float2*>(dst),
          nullptr, 1, 1, e.xw_lead, e.xw_window ? (unsigned long long)count : 0ULL);
    CUDA_TRY(cudaGetLastError());
    e.xw_pos = pos + global_count;
    scaled_in_draw = e.colored;
  }
  else
  {
    CURAND_TRY(curandSetStream(e.gen, st));
    if (e.desc.world_size == 1 && e.curand_pos == pos)
    {
      // generator already sits at `start`: plain continuation, exactly what the reference does call after call
      CURAND_TRY(curandGenerateNormal(e.gen, dst, count, 0.0f, 1.0f));
    }
    else
    {
      // XORWOW default ordering interleaves 4096 streams x 2 normals: absolute offsets are honoured at multiples of
      // 8192 (probed on B200, tools/curand_probe.cu), so start from the aligned position below and discard the lead-in.
      const unsigned long long aligned = (start / 8192ULL) * 8192ULL;
      const size_t lead = (size_t)(start - aligned);
      if (((lead + count) & 1) != 0)
        return fail(MPPIB_ERR_UNSUPPORTED, "cuRAND normal draws need an even count (lead %zu + count %zu)", lead, count);
      CURAND_TRY(curandSetGeneratorOffset(e.gen, aligned));
      CURAND_TRY(curandGenerateNormal(e.gen, dst - lead, lead + count, 0.0f, 1.0f));
    }
    e.curand_pos = (e.desc.world_size == 1) ? pos + global_count : mppib_engine::kNoPos;
  }
  if (e.colored)
  {
    if (!scaled_in_draw)
    {
      const size_t ncomplex = count / 2;
      const int blocks = (int)std::min<size_t>((ncomplex + 255) / 256, 148 * 16);
      colored_scale_kernel<<<blocks, 256, 0, st>>>(e.spec_d, e.coeffs_d, ncomplex, e.C, e.F);
      CUDA_TRY(cudaGetLastError());
    }
    if (cufftSetStream(e.fft_plan, st) != CUFFT_SUCCESS)
      return fail(MPPIB_ERR_CUDA, "cufftSetStream failed");
    const cufftResult fr = cufftExecC2R(e.fft_plan, reinterpret_cast<cufftComplex*>(e.spec_d), e.time_d);
    if (fr != CUFFT_SUCCESS)
      return fail(MPPIB_ERR_CUDA, "cufftExecC2R failed: %d", (int)fr);
    int rc = colored_rearrange(e, buf, st, offset_t);
    if (rc != MPPIB_OK)
      return rc;
  }
  CUDA_TRY(cudaEventRecord(e.ev_last_gen, st));
  e.any_gen = true;
  return MPPIB_OK;
}

// Makes eps_buf[cur_buf] hold the block at rng_offset (taking the prefetched buffer if it is the right one) and
// advances rng_offset. Everything is ordered on the main stream when this returns.
static int draw_noise(mppib_engine& e, int offset_t)
{
  if (e.prefetch_valid && e.prefetch_pos == e.rng_offset)
  {
    CUDA_TRY(cudaStreamWaitEvent(e.stream, e.ev_gen_done[e.prefetch_buf], 0));
    e.cur_buf = e.prefetch_buf;
    if (e.colored && e.buf_offset_t[e.cur_buf] != offset_t)
    {
      // the prefetch assumed another optimization_stride: redo the (cheap) rearrange from the time-domain buffer,
      // which still holds this block (no later draw has been issued)
      int rc = colored_rearrange(e, e.cur_buf, e.stream, offset_t);
      if (rc != MPPIB_OK)
        return rc;
      CUDA_TRY(cudaEventRecord(e.ev_rearr, e.stream));
      e.rearr_recorded = true;
    }
  }
  else
  {
    // the main stream orders this draw after every earlier K1 that read eps_buf[cur_buf]
    int rc = gen_draw(e, e.cur_buf, e.stream, e.rng_offset, offset_t);
    if (rc != MPPIB_OK)
      return rc;
  }
  e.prefetch_valid = false;
  e.eps_d = e.eps_buf[e.cur_buf];
  e.tmap = e.tmap_buf[e.cur_buf];
  e.rng_offset += e.draw_global;
  if (e.colored)
    e.colored_offset_t = offset_t;  // what the next prefetch assumes
  return MPPIB_OK;
}

// After K1 of the current solve has been enqueued: start the draw of the NEXT block into the other buffer on the side
// stream. It only has to wait for the K1 that last read that buffer.
static int prefetch_next(mppib_engine& e)
{
  if (!e.prefetch_enabled)
    return MPPIB_OK;
  const int nb = e.cur_buf ^ 1;
  if (e.k1_recorded[nb])
    CUDA_TRY(cudaStreamWaitEvent(e.side_stream, e.ev_k1_done[nb], 0));
  int rc = gen_draw(e, nb, e.side_stream, e.rng_offset, e.colored_offset_t);
  if (rc != MPPIB_OK)
    return rc;
  CUDA_TRY(cudaEventRecord(e.ev_gen_done[nb], e.side_stream));
  e.prefetch_valid = true;
  e.prefetch_pos = e.rng_offset;
  e.prefetch_buf = nb;
  return MPPIB_OK;
}

// K2 launch. `pdl` = programmatic dependent launch: the grid may start while the preceding kernel on the stream (K1) is
// still running and blocks at griddepcontrol.wait until that kernel has completed — hides K2's launch latency.
static int launch_combine_one(mppib_engine& e, const float* records, const float4* headers, int nrec, int normalize,
                              float* out, float* out2, bool pdl, bool final_stage = false)
{
  unsigned* counter = e.k2_counter_d;
  volatile unsigned* flag = nullptr;
  unsigned seq = 0;
  if (final_stage && e.spin_wait && out2 != nullptr)
  {
    flag = e.done_flag_dev;
    seq = ++e.solve_seq;
    e.flag_armed = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((e.TC + kCombineCols - 1) / kCombineCols, e.D, 1);
  cfg.blockDim = dim3(kCombineCols * kCombineGroups, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = e.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  const float lambda_inv = (float)(1.0 / e.lambda);
  CUDA_TRY(cudaLaunchKernelEx(&cfg, combine_kernel, records, headers, nrec, e.D, e.TC, e.pstride, lambda_inv, normalize,
                              out, out2, counter, flag, seq));
  return MPPIB_OK;
}

static int launch_combine(mppib_engine& e, bool after_k1)
{
  // only a final-stage K2 publishes the completion flag; the Tsallis reduction and the peer-memory exchange kernel end the
  // solve without it, and wait_for_stream then falls back to cudaStreamSynchronize instead of trusting a stale flag
  e.flag_armed = false;
  const bool pdl = after_k1 && e.use_pdl;
  float* host_copy = e.mapped_result ? e.result_h_dev : nullptr;
  if (e.tsallis_gamma != 0.0f && e.tsallis_r != 0.0f)
  {
    // K2 for the global baseline (device copy only), then the Tsallis-weighted reduction of the written-back controls
    int rc1 = launch_combine_one(e, e.partials_d, e.headers_d, e.grid, 1, e.result_d, nullptr, pdl, false);
    if (rc1 != MPPIB_OK)
      return rc1;
    const dim3 grid((e.TC + 31) / 32, e.D, 1);
    tsallis_reduce_kernel<<<grid, 512, 0, e.stream>>>(e.costs_d, e.controls_d, e.n_local, e.D, e.TC, e.pstride,
                                                      e.tsallis_gamma, e.tsallis_r, e.result_d, host_copy);
    CUDA_TRY(cudaGetLastError());
    if (!e.mapped_result)
      CUDA_TRY(cudaMemcpyAsync(e.result_h, e.result_d, (size_t)e.D * e.pstride * sizeof(float), cudaMemcpyDeviceToHost,
                               e.stream));
    return MPPIB_OK;
  }
  if (e.desc.world_size == 1 || !e.comm)
  {
    int rc1 = launch_combine_one(e, e.partials_d, e.headers_d, e.grid, 1, e.result_d, host_copy, pdl, true);
    if (rc1 == MPPIB_OK && !e.mapped_result)
      CUDA_TRY(cudaMemcpyAsync(e.result_h, e.result_d, (size_t)e.D * e.pstride * sizeof(float), cudaMemcpyDeviceToHost,
                               e.stream));
    return rc1;
  }
  // rank record (un-normalised) -> all-gather -> merge of the world_size records (normalised)
  int rc = launch_combine_one(e, e.partials_d, e.headers_d, e.grid, 0, e.rank_rec_d, nullptr, pdl);
  if (rc != MPPIB_OK)
    return rc;
  if (e.p2p)
  {  // KX: push to peers over NVLink, wait for theirs, merge — one launch
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(1, 1, 1);
    cfg.blockDim = dim3(512, 1, 1);
    cfg.stream = e.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = e.use_pdl ? 1 : 0;
    const unsigned seq = ++e.p2p_seq;
    const float lambda_inv = (float)(1.0 / e.lambda);
    CUDA_TRY(cudaLaunchKernelEx(&cfg, exchange_merge_kernel, (const float*)e.rank_rec_d, e.peers, e.desc.world_size,
                                e.desc.rank, e.D, e.TC, e.pstride, lambda_inv, seq, e.result_d, host_copy));
    if (!e.mapped_result)
      CUDA_TRY(cudaMemcpyAsync(e.result_h, e.result_d, (size_t)e.D * e.pstride * sizeof(float), cudaMemcpyDeviceToHost,
                               e.stream));
    return MPPIB_OK;
  }
  const size_t rec_floats = (size_t)e.D * e.pstride;
  rc = g_nccl.AllGather(e.rank_rec_d, e.gather_d, rec_floats, kNcclFloat, e.comm, e.stream);
  if (rc != 0)
    return fail(MPPIB_ERR_NCCL, "ncclAllGather failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  const int nh = e.desc.world_size * e.D;
  record_headers_kernel<<<(nh + 63) / 64, 64, 0, e.stream>>>(e.gather_d, e.desc.world_size, e.D, e.pstride,
                                                            e.gather_hdr_d);
  CUDA_TRY(cudaGetLastError());
  rc = launch_combine_one(e, e.gather_d, e.gather_hdr_d, e.desc.world_size, 1, e.result_d, host_copy, false, true);
  if (rc == MPPIB_OK && !e.mapped_result)
    CUDA_TRY(cudaMemcpyAsync(e.result_h, e.result_d, (size_t)e.D * e.pstride * sizeof(float), cudaMemcpyDeviceToHost,
                             e.stream));
  return rc;
}

static int check_ready(mppib_engine* e)
{
  if (!e)
    return fail(MPPIB_ERR_INVALID_ARG, "null engine");
  if (!e->have_dyn || !e->have_cost || !e->have_sampler)
    return fail(MPPIB_ERR_STATE, "dynamics / cost / sampler parameter blobs must be set before solving");
  if (e->desc.dynamics_id == MPPIB_DYN_AUTORALLY_NN && !e->nn_theta_d)
    return fail(MPPIB_ERR_STATE, "MPPIB_BLOB_NN_WEIGHTS not set");
  if (e->desc.cost_id == MPPIB_COST_AR_STANDARD && !e->costmap_tex)
    return fail(MPPIB_ERR_STATE, "MPPIB_BLOB_COSTMAP not set");
  if (e->desc.dynamics_id == MPPIB_DYN_RACER_LSTM && !e->have_lstm)
    return fail(MPPIB_ERR_STATE, "MPPIB_BLOB_LSTM_WEIGHTS not set");
  if (e->desc.world_size > 1 && !e->comm)
    return fail(MPPIB_ERR_STATE, "world_size > 1 but mppib_comm_init was not called");
  return MPPIB_OK;
}

static void read_result(mppib_engine& e, float* U_out, mppib_solve_stats* stats)
{
  for (int d = 0; d < e.D; d++)
  {
    const float* r = e.result_h + (size_t)d * e.pstride;
    if (stats)
    {
      stats[d].baseline = r[0];
      stats[d].normalizer = r[1];
      stats[d].sum_w2 = r[2];
      stats[d].pad = 0.0f;
    }
    if (U_out)
      memcpy(U_out + (size_t)d * e.TC, r + kPartialHeader, sizeof(float) * e.TC);
  }
}

// =================================================================================================================
extern "C" {

int mppib_version(void)
{
  return 100;
}

// A (dynamics, cost) pair compiled out of tree from csrc/engine_internal.cuh (plugins_example/): the reference lets users
// instantiate its templates with their own classes (dynamics.cuh:67-76, cost.cuh:34-35); here they instantiate OUR kernels
// with their device twins in a second shared library and register the result. ids >= MPPIB_USER_ID_BASE.
int mppib_register_pair(const void* pair_entry, size_t entry_bytes, unsigned abi)
{
  if (!pair_entry)
    return fail(MPPIB_ERR_INVALID_ARG, "null pair entry");
  if (entry_bytes != sizeof(PairEntry) || abi != engine_abi())
    return fail(MPPIB_ERR_INVALID_ARG, "plugin built against another revision of libmppi_b200 (entry %zu / %zu bytes, abi %u / %u)",
                entry_bytes, sizeof(PairEntry), abi, engine_abi());
  const PairEntry* in = static_cast<const PairEntry*>(pair_entry);
  if (in->dyn_id < MPPIB_USER_ID_BASE || in->cost_id < MPPIB_USER_ID_BASE)
    return fail(MPPIB_ERR_INVALID_ARG, "user pairs take dynamics / cost ids >= %d (got %d, %d)", MPPIB_USER_ID_BASE, in->dyn_id,
                in->cost_id);
  if (in->C != 1 && in->C != 2 && in->C != 4)
    return fail(MPPIB_ERR_UNSUPPORTED, "CONTROL_DIM %d: must divide a 16-byte noise group (1, 2 or 4)", in->C);
  for (PairEntry* p : user_pairs())
    if (p->dyn_id == in->dyn_id && p->cost_id == in->cost_id)
    {
      *p = *in;  // re-registration (plugin reloaded)
      return MPPIB_OK;
    }
  user_pairs().push_back(new PairEntry(*in));
  return MPPIB_OK;
}

// dlopen a plugin library and run its `int mppib_plugin_init(void)` (which calls register_pair<...>() for each of its pairs)
int mppib_load_plugin(const char* path)
{
  if (!path)
    return fail(MPPIB_ERR_INVALID_ARG, "null path");
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h)
    return fail(MPPIB_ERR_INVALID_ARG, "dlopen(%s) failed: %s", path, dlerror());
  // loading the same library again is a no-op (dlopen hands back the same handle; its pairs are registered already)
  static std::mutex mu;
  static std::vector<void*> loaded;
  std::lock_guard<std::mutex> lock(mu);
  for (void* l : loaded)
    if (l == h)
    {
      dlclose(h);  // drop the extra reference
      return MPPIB_OK;
    }
  typedef int (*init_fn)(void);
  init_fn init = (init_fn)dlsym(h, "mppib_plugin_init");
  if (!init)
  {
    dlclose(h);
    return fail(MPPIB_ERR_INVALID_ARG, "%s exports no mppib_plugin_init", path);
  }
  const int rc = init();
  if (rc == MPPIB_OK)
    loaded.push_back(h);
  else
    dlclose(h);  // e.g. built against another revision: a rebuilt file at the same path must be mapped afresh
  return rc;
}

const char* mppib_last_error(void)
{
  return g_last_error.c_str();
}

const char* mppib_strerror(int status)
{
  switch (status)
  {
    case MPPIB_OK:
      return "ok";
    case MPPIB_ERR_INVALID_ARG:
      return "invalid argument";
    case MPPIB_ERR_UNSUPPORTED:
      return "unsupported plugin combination or size";
    case MPPIB_ERR_CUDA:
      return "CUDA runtime error";
    case MPPIB_ERR_CURAND:
      return "cuRAND error";
    case MPPIB_ERR_NO_DEVICE:
      return "no CUDA device (the engine has no CPU fallback)";
    case MPPIB_ERR_NCCL:
      return "NCCL error";
    case MPPIB_ERR_SMEM:
      return "rollout tile does not fit in shared memory";
    case MPPIB_ERR_CUFFT:
      return "cuFFT error";
    case MPPIB_ERR_STATE:
      return "call order violated";
  }
  return "unknown status";
}

int mppib_create(mppib_engine** out, const mppib_desc* desc)
{
  if (!out || !desc)
    return fail(MPPIB_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  if (desc->num_rollouts <= 0 || desc->num_timesteps <= 0)
    return fail(MPPIB_ERR_INVALID_ARG, "num_rollouts and num_timesteps must be positive");
  if (desc->num_distributions < 1 || desc->num_distributions > MPPIB_MAX_DISTRIBUTIONS)
    return fail(MPPIB_ERR_INVALID_ARG, "num_distributions must be 1 or 2");
  const int world = desc->world_size <= 0 ? 1 : desc->world_size;
  if (desc->rank < 0 || desc->rank >= world)
    return fail(MPPIB_ERR_INVALID_ARG, "rank out of range");
  if (desc->sampler_id == MPPIB_SAMPLER_NLN && (desc->world_size != 1 || desc->num_distributions != 1))
    return fail(MPPIB_ERR_UNSUPPORTED, "the NLN sampler is built for one rank and one distribution");
  if (desc->sampler_id != MPPIB_SAMPLER_GAUSSIAN && desc->sampler_id != MPPIB_SAMPLER_COLORED_NOISE &&
      desc->sampler_id != MPPIB_SAMPLER_NLN)
    return fail(MPPIB_ERR_UNSUPPORTED, "sampler %d is not built into this library", desc->sampler_id);
  if (desc->sampler_id == MPPIB_SAMPLER_COLORED_NOISE && desc->num_distributions != 1)
    return fail(MPPIB_ERR_UNSUPPORTED,
                "ColoredNoise draws independent noise per distribution (colored_noise.cu:291); only "
                "num_distributions == 1 is built");

  if (desc->dynamics_id == MPPIB_DYN_RACER_LSTM)
  {
    const int H = desc->model_dims[0], L1 = desc->model_dims[1];
    if (H < 1 || H > plugins::RacerLSTMDynamics::MAX_HIDDEN || L1 < 1 || L1 > plugins::RacerLSTMDynamics::MAX_HEAD)
      return fail(MPPIB_ERR_INVALID_ARG, "RacerDubinsElevationLSTMSteering: model_dims = {hidden_dim %d, head width %d} "
                                         "outside [1, %d] x [1, %d]",
                  H, L1, plugins::RacerLSTMDynamics::MAX_HIDDEN, plugins::RacerLSTMDynamics::MAX_HEAD);
    if (desc->num_distributions != 1)
      return fail(MPPIB_ERR_UNSUPPORTED, "RacerDubinsElevationLSTMSteering is built for num_distributions == 1 only");
  }
  const PairEntry* entry = nullptr;
  for (const auto& p : kPairs)
    if (p.dyn_id == desc->dynamics_id && p.cost_id == desc->cost_id)
      entry = &p;
  for (const PairEntry* p : user_pairs())  // out-of-tree pairs (mppib_load_plugin / mppib_register_pair)
    if (p->dyn_id == desc->dynamics_id && p->cost_id == desc->cost_id)
      entry = p;
  // Autorally pair: the network runs on register-level mma.sync by default (plugins/nn_mma.cuh; K1 257 us against 350 us
  // for the FP32 FFMA2 form at N = 32768, T = 100). MPPIB_FLAG_NN_FFMA2 / MPPIB_NN_FFMA2 keep the FFMA2 form,
  // MPPIB_FLAG_NN_TENSOR / MPPIB_NN_TENSOR select the tcgen05 kernel (which is built on the FFMA2 entry).
  const bool nn_other = (desc->flags & (MPPIB_FLAG_NN_FFMA2 | MPPIB_FLAG_NN_TENSOR)) || getenv("MPPIB_NN_FFMA2") ||
                        getenv("MPPIB_NN_TENSOR");
  if (entry && desc->dynamics_id == MPPIB_DYN_AUTORALLY_NN && (!nn_other || (desc->flags & MPPIB_FLAG_NN_MMA)))
  {
    // samples per warp of the generic kernel's network (plugins/nn_mma.cuh): 32. Narrower sample groups (MPPIB_SPW = 16 / 8)
    // halve / quarter a warp's tensor work per step, but measured on B200 the step time of a lone warp barely moves
    // (1.70 / 1.73 / 1.45 us per step at 4096 rollouts, profiles/r02_autorally_k1_notes.md): the chain, not the work, is
    // the limit — which the warp-specialised kernel (rollout_kernel_ar_ws.cuh, the default for D == 1) removes instead.
    int spw = 32;
    if (const char* sv = getenv("MPPIB_SPW"))
    {
      const int v = atoi(sv);
      if (v == 32 || v == 16 || v == 8)
        spw = v;
    }
    for (const auto& p : kPairsMma)
      if (p.cost_id == desc->cost_id && p.spw == spw)
        entry = &p;
  }
  // steering LSTM at hidden_dim 32 (head width <= 24): gates and head as mma.sync products, hidden / cell state in fragment
  // layout in registers (plugins/lstm_mma.cuh) — 11.2 ms -> see profiles/r02_racer_h32_* per C5-sized solve.
  // MPPIB_FLAG_LSTM_SIMT / MPPIB_LSTM_SIMT keep the one-thread-per-sample network.
  if (entry && desc->dynamics_id == MPPIB_DYN_RACER_LSTM && desc->model_dims[0] == lstm_mma::H &&
      desc->model_dims[1] <= 8 * lstm_mma::kHeadTiles && !(desc->flags & MPPIB_FLAG_LSTM_SIMT) && !getenv("MPPIB_LSTM_SIMT"))
    for (const auto& p : kPairsLstmMma)
      if (p.cost_id == desc->cost_id)
        entry = &p;
  if (!entry)
    return fail(MPPIB_ERR_UNSUPPORTED, "no kernel registered for dynamics %d + cost %d", desc->dynamics_id,
                desc->cost_id);

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
  {
    cudaGetLastError();
    return fail(MPPIB_ERR_NO_DEVICE, "no CUDA device visible; libmppi_b200 has no CPU path");
  }
  if (desc->device < 0 || desc->device >= ndev)
    return fail(MPPIB_ERR_INVALID_ARG, "device %d out of range (%d devices)", desc->device, ndev);
  CUDA_TRY(cudaSetDevice(desc->device));

  mppib_engine* e = new mppib_engine();
  e->desc = *desc;
  e->desc.world_size = world;
  e->S = entry->S;
  e->C = entry->C;
  e->O = entry->O;
  e->D = desc->num_distributions;
  e->N = desc->num_rollouts;
  e->T = desc->num_timesteps;
  e->TC = e->T * e->C;
  e->launch_rollout = entry->launch;
  e->prepare = entry->prepare;
  e->init_eval = entry->init_eval;
  e->sampled_traj = entry->sampled_traj;
  e->nominal_traj = entry->nominal_traj;
  e->dyn_param_bytes = entry->dyn_bytes;
  e->cost_param_bytes = entry->cost_bytes;
  e->dyn_shared_floats_fn = entry->dyn_shared_floats;
  e->cost_shared_floats = entry->cost_shared_floats;
  e->rmppi = (desc->flags & MPPIB_FLAG_RMPPI) != 0;
  if (e->rmppi && desc->num_distributions != 2)
  {
    delete e;
    return fail(MPPIB_ERR_INVALID_ARG, "MPPIB_FLAG_RMPPI needs num_distributions == 2 (nominal, real)");
  }
  e->writeback = e->rmppi || (desc->flags & MPPIB_FLAG_WRITEBACK_CONTROLS) != 0;
  e->use_pdl = !getenv("MPPIB_NO_PDL");
  e->mapped_result = !getenv("MPPIB_NO_MAPPED_RESULT");
  // measured on B200: polling a mapped flag is not faster than cudaStreamSynchronize (39.99 vs 40.33 us per cartpole
  // solve) and costs K2 two system fences; off unless MPPIB_SPIN_WAIT is set
  e->spin_wait = e->mapped_result && getenv("MPPIB_SPIN_WAIT") != nullptr;

  auto bail = [&](int rc) {
    mppib_destroy(e);
    return rc;
  };

  if (e->D * e->TC > kMaxMeanFloats)
    return bail(fail(MPPIB_ERR_UNSUPPORTED, "D*T*C = %d exceeds %d", e->D * e->TC, kMaxMeanFloats));
  e->nchunks = (e->TC + kChunkFloats - 1) / kChunkFloats;
  if (e->nchunks > kMaxChunks)
    return bail(fail(MPPIB_ERR_UNSUPPORTED, "T*C = %d exceeds %d", e->TC, kMaxChunks * kChunkFloats));

  // rollout sharding: contiguous slices, remainder to the last rank
  const int per = e->N / world;
  e->n_offset = per * desc->rank;
  e->n_local = (desc->rank == world - 1) ? (e->N - e->n_offset) : per;
  if (e->n_local <= 0)
    return bail(fail(MPPIB_ERR_INVALID_ARG, "rank %d of %d has no rollouts (N=%d)", desc->rank, world, e->N));

  // launch geometry: one thread per sample; BX samples per CTA, whole-horizon noise tile resident in shared memory.
  // The rollout is bound by the T-step dependency chain, so a CTA takes the same time whatever its width: the block
  // width is chosen to put every CTA in ONE wave (no tail wave running at a fraction of the chip), preferring the
  // narrowest such width (more SMs busy, fewer warps contending per scheduler); 64 if several waves are unavoidable.
  int max_smem = 0, num_sms = 0, smem_per_sm = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, desc->device));
  CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, desc->device));
  CUDA_TRY(cudaDeviceGetAttribute(&smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, desc->device));
  std::function<int(int)> smem_for = [&](int b) {
    return (int)rollout_smem_layout(b, e->nchunks, e->D, e->TC, e->dyn_shared_floats_fn(e->desc.model_dims, b),
                                    e->cost_shared_floats(e->T))
        .total;
  };
  // samples per thread (rollout_kernel.cuh): 1. SPT = 2 halves the shared-memory wavefronts per sample of the NN model
  // (38 M instead of 74 M, ncu) but a lone warp per scheduler cannot overlap its own FFMA2 / MUFU / latency phases the
  // way two warps do: measured 636 us against 352 us at N = 32768 and 1854 us against 685 us at N = 65536 on B200
  // (profiles/r01_autorally_k1_notes.md). Kept selectable for experiments through MPPIB_SPT.
  int spt = 1;
  if (const char* s = getenv("MPPIB_SPT"))
  {
    const int v = atoi(s);
    if (v >= 1 && v <= entry->max_spt && (v == 1 || e->D == 1))
      spt = v;
  }
  e->spt = spt;
  const int lps = 32 / entry->spw;
  e->lps = lps;
  // Autorally pair, one system: the warp-specialised K1 (rollout_kernel_ar_ws.cuh) — a producer and a consumer warp per 32
  // samples. MPPIB_NO_WS / MPPIB_SPW / MPPIB_SPT keep the generic kernel (A/B runs, tests of the generic form).
  const bool is_mma32 = entry >= kPairsMma && entry < kPairsMma + sizeof(kPairsMma) / sizeof(kPairsMma[0]) && entry->spw == 32;
  e->ar_ws = is_mma32 && e->D == 1 && !e->rmppi && spt == 1 && !getenv("MPPIB_NO_WS") && !getenv("MPPIB_SPW") &&
             !(desc->flags & MPPIB_FLAG_NO_WARP_SPEC);
  const bool ws = e->ar_ws;
  // samples per producer warp: 16 while the GPU is throughput-bound; 8 (four producers per 32 samples, one per scheduler)
  // once it holds so few rollouts that a group's step time sets K1 (measured crossover: profiles/r02_autorally_k1_notes.md §5)
  e->ws_pspw = e->n_local <= kWsPspw8MaxRollouts ? 8 : 16;
  if (const char* sv = getenv("MPPIB_WS_PSPW"))
    if (atoi(sv) == 32 || atoi(sv) == 16 || atoi(sv) == 8)
      e->ws_pspw = atoi(sv);
  const int ws_wpg = ar_ws::warpsPerGroup(e->ws_pspw);
  if (ws)
    smem_for = [&](int b) {
      return (int)rollout_smem_layout(b, e->nchunks, 1, e->TC, ar_ws::sharedFloats(b), e->cost_shared_floats(e->T)).total;
    };
  const int unit = ws ? 32 : entry->spw * spt;  // samples per warp of threads
  const int max_bx = ws ? 32 * ar_ws::maxGroups(e->ws_pspw) : entry->max_block_threads / 32 * unit;  // samples per CTA (__launch_bounds__)
  auto threads_for = [&](int b) { return ws ? ws_wpg * b : b / spt * lps; };
  // resident CTAs per SM: shared memory (+1 KB the hardware reserves per CTA), threads, and for the warp-specialised kernel
  // its 128-register budget
  auto ctas_per_sm = [&](int b, int sm) {
    int n = std::min(std::min(smem_per_sm / (sm + 1024), 2048 / threads_for(b)), 32);
    if (ws)  // registers: __launch_bounds__(maxThreads, 1) lets ptxas use 65536 / maxThreads per thread
      n = std::min(n, ar_ws::maxThreads(e->ws_pspw) / threads_for(b));
    return n;
  };
  int bx = 0;
  if (const char* s = getenv("MPPIB_BX"))
  {
    bx = atoi(s);
    if (bx < unit || bx > 512 || (bx % unit) != 0)
      bx = 0;
  }
  if (bx == 0 && ws)
  {
    // warp-specialised kernel: up to one warp per scheduler, one pair per CTA (no two producers ever share a scheduler);
    // beyond that ONE CTA per SM, as narrow as covers n_local, so that every SM is busy and the kernel's alternating role
    // table (P C C P C P P C) balances producers over the four schedulers. Wider than fits: the wave rule below.
    const long pairs_total = (e->n_local + 31) / 32;
    if (ws_wpg * pairs_total <= 4L * num_sms)
      bx = 32;
    else
    {
      const int need = (int)(((e->n_local + num_sms - 1) / num_sms + 31) / 32) * 32;
      if (need <= max_bx && smem_for(need) <= max_smem)
        bx = need;
    }
  }
  if (bx == 0)
  {
    int best = 0;
    long best_waves = 1L << 40;
    for (int cand = ws ? unit : std::max(unit, 64 / lps); cand <= max_bx; cand += unit)
    {
      const int sm = smem_for(cand);
      if (sm > max_smem)
        break;
      const int per_sm = ctas_per_sm(cand, sm);
      if (per_sm < 1)
        break;
      const long blocks = (e->n_local + cand - 1) / cand;
      const long waves = (blocks + (long)per_sm * num_sms - 1) / ((long)per_sm * num_sms);
      if (waves < best_waves)
      {
        best_waves = waves;
        best = cand;
      }
    }
    bx = best ? best : unit;
  }
  if (bx > max_bx)
    bx = max_bx;
  bx = (bx / unit) * unit;  // whole warps of threads
  if (bx < unit)
    bx = unit;
  while (smem_for(bx) > max_smem && bx > unit)
    bx -= unit;
  e->smem_bytes = (uint32_t)smem_for(bx);
  if ((int)e->smem_bytes > max_smem)
    return bail(fail(MPPIB_ERR_SMEM, "noise tile needs %u B of shared memory, device allows %d", e->smem_bytes,
                     max_smem));
  // tensor-core variant of the Autorally pair: fixed 128-sample CTAs (one UMMA tile), streaming noise ring. Opt-in:
  // measured on B200 it does not beat the FFMA2 kernel yet (378 vs 353 us at N=32768, T=100 — three exposed MMA round
  // trips per step with only ~7 warps per SM to cover them; profiles/r01_autorally_k1_notes.md)
  if (desc->dynamics_id == MPPIB_DYN_AUTORALLY_NN && desc->cost_id == MPPIB_COST_AR_STANDARD && e->D == 1 &&
      (e->TC % 4) == 0 && !(desc->flags & MPPIB_FLAG_NO_TMA) && !getenv("MPPIB_NO_TMA") &&
      ((desc->flags & MPPIB_FLAG_NN_TENSOR) || getenv("MPPIB_NN_TENSOR")))
  {
    e->nn_tc = true;
    bx = nn_tc::kRows;
    e->smem_bytes = nn_tc::layout(e->TC, e->T).total;
    if ((int)e->smem_bytes > max_smem)
      return bail(fail(MPPIB_ERR_SMEM, "tensor-core rollout needs %u B of shared memory, device allows %d",
                       e->smem_bytes, max_smem));
  }
  // streaming K1 (rollout_kernel.cuh: STREAM): chosen when the resident whole-horizon tile forces several waves and the
  // ring variant needs fewer; MPPIB_STREAM=0/1 overrides
  {
    const bool tma_ok = !(desc->flags & MPPIB_FLAG_NO_TMA) && (e->TC % 4 == 0) && !getenv("MPPIB_NO_TMA");
    const int sm_res = smem_for(bx);
    const int per_sm_res = std::max(1, ctas_per_sm(bx, sm_res));
    const long blocks_res = (e->n_local + bx - 1) / bx;
    const long waves_res = (blocks_res + (long)per_sm_res * num_sms - 1) / ((long)per_sm_res * num_sms);
    int sbx = 64;
    if (const char* sb = getenv("MPPIB_BX"))
    {
      const int v = atoi(sb);
      if (v >= unit && v <= max_bx && (v % unit) == 0)
        sbx = v;
    }
    const int sm_str = (int)rollout_smem_layout(sbx, e->ring, e->D, e->TC, e->dyn_shared_floats_fn(e->desc.model_dims, sbx),
                                                e->cost_shared_floats(e->T))
                           .total;
    bool want = false;
    if (tma_ok && !e->rmppi && !e->nn_tc && !ws && spt == 1 && e->nchunks > e->ring && sm_str <= max_smem)
    {
      const int per_sm_str = entry->stream_blocks_per_sm(e->D, threads_for(sbx), (size_t)sm_str);
      if (per_sm_str > 0)
      {
        const long blocks_str = (e->n_local + sbx - 1) / sbx;
        const long waves_str = (blocks_str + (long)per_sm_str * num_sms - 1) / ((long)per_sm_str * num_sms);
        want = waves_res > 1 && waves_str < waves_res;
        if (const char* sv = getenv("MPPIB_STREAM"))
          want = atoi(sv) != 0;
      }
    }
    if (want)
    {
      e->stream_k1 = true;
      if (getenv("MPPIB_STREAM_READBACK"))  // A/B: the round-1 form (controls written back and re-read by the epilogue)
        e->writeback = true;
      bx = sbx;
      e->smem_bytes = (uint32_t)sm_str;
    }
  }
  e->bx = bx;
  e->dyn_shared_floats = ws ? ar_ws::sharedFloats(bx) : e->dyn_shared_floats_fn(e->desc.model_dims, bx);
  e->grid = (e->n_local + bx - 1) / bx;
  if (e->grid > kCombineMaxRecords)
    return bail(fail(MPPIB_ERR_UNSUPPORTED, "%d rollout blocks exceed the combine kernel's %d records; raise MPPIB_BX",
                     e->grid, kCombineMaxRecords));
  e->pstride = ((kPartialHeader + e->TC + 3) / 4) * 4;
  e->use_tma = !(desc->flags & MPPIB_FLAG_NO_TMA) && (e->TC % 4 == 0) && !getenv("MPPIB_NO_TMA");

  if (desc->stream)
    e->stream = (cudaStream_t)desc->stream;
  else
  {
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&e->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess)
      return bail(fail(MPPIB_ERR_CUDA, "cudaStreamCreate failed"));
    e->own_stream = true;
  }

#define CUDA_TRY_B(expr)                                                                                               \
  do                                                                                                                   \
  {                                                                                                                    \
    cudaError_t _e = (expr);                                                                                           \
    if (_e != cudaSuccess)                                                                                             \
    {                                                                                                                  \
      cudaGetLastError();                                                                                              \
      return bail(fail(MPPIB_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)));                               \
    }                                                                                                                  \
  } while (0)

  const size_t noise_floats = (size_t)e->n_local * e->TC;
  const size_t lead_floats = 8192;  // room for the offset-alignment lead-in (see draw_noise)
  e->colored = desc->sampler_id == MPPIB_SAMPLER_COLORED_NOISE;
  e->nln = desc->sampler_id == MPPIB_SAMPLER_NLN;
  e->F = e->T + 1;
  {
    // normals per generateSamples call and this rank's slice of them
    // colored_noise.cu:341-343 / gaussian.cu:380 / nln.cu:114-122 (C log-normal planes of N*T, then N*T*C normals)
    const unsigned long long per_rollout =
        e->colored ? 2ULL * e->C * e->F : (e->nln ? 2ULL * e->TC : (unsigned long long)e->TC);
    e->draw_global = per_rollout * (unsigned long long)e->N;
    e->draw_start = per_rollout * (unsigned long long)e->n_offset;
    e->draw_local = (size_t)(per_rollout * (unsigned long long)e->n_local);
  }
  CUDA_TRY_B(cudaMalloc(&e->noise_alloc, (lead_floats + noise_floats + 8) * sizeof(float)));
  e->eps_d = e->noise_alloc + lead_floats;  // cudaMalloc is 256-B aligned and 8192*4 keeps that
  CUDA_TRY_B(cudaMemsetAsync(e->noise_alloc, 0, (lead_floats + noise_floats + 8) * sizeof(float), e->stream));
  e->eps_buf[0] = e->eps_buf[1] = e->eps_d;
  e->prefetch_enabled = !(desc->flags & MPPIB_FLAG_NO_PREFETCH) && !getenv("MPPIB_NO_PREFETCH");
  CUDA_TRY_B(cudaEventCreateWithFlags(&e->ev_last_gen, cudaEventDisableTiming));
  if (e->prefetch_enabled)
  {
    CUDA_TRY_B(cudaMalloc(&e->noise_alloc2, (lead_floats + noise_floats + 8) * sizeof(float)));
    CUDA_TRY_B(cudaMemsetAsync(e->noise_alloc2, 0, (lead_floats + noise_floats + 8) * sizeof(float), e->stream));
    e->eps_buf[1] = e->noise_alloc2 + lead_floats;
    {
      // the prefetch (next solve's noise) yields to the solve in flight: lowest priority for the side stream
      int prio_lo = 0, prio_hi = 0;
      CUDA_TRY_B(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      CUDA_TRY_B(cudaStreamCreateWithPriority(&e->side_stream, cudaStreamNonBlocking, prio_lo));
    }
    for (int i = 0; i < 2; i++)
    {
      CUDA_TRY_B(cudaEventCreateWithFlags(&e->ev_k1_done[i], cudaEventDisableTiming));
      CUDA_TRY_B(cudaEventCreateWithFlags(&e->ev_gen_done[i], cudaEventDisableTiming));
    }
  }
  CUDA_TRY_B(cudaMalloc(&e->costs_d, (size_t)e->D * e->n_local * sizeof(float)));
  CUDA_TRY_B(cudaMalloc(&e->partials_d, (size_t)e->grid * e->D * e->pstride * sizeof(float)));
  CUDA_TRY_B(cudaMalloc(&e->headers_d, (size_t)e->grid * e->D * sizeof(float4)));
  CUDA_TRY_B(cudaMalloc(&e->result_d, (size_t)e->D * e->pstride * sizeof(float)));
  CUDA_TRY_B(cudaHostAlloc(&e->result_h, (size_t)e->D * e->pstride * sizeof(float), cudaHostAllocMapped));
  memset(e->result_h, 0, (size_t)e->D * e->pstride * sizeof(float));
  CUDA_TRY_B(cudaHostGetDevicePointer(&e->result_h_dev, e->result_h, 0));
  CUDA_TRY_B(cudaMalloc(&e->k2_counter_d, sizeof(unsigned)));
  CUDA_TRY_B(cudaMemsetAsync(e->k2_counter_d, 0, sizeof(unsigned), e->stream));
  {
    unsigned* f = nullptr;
    CUDA_TRY_B(cudaHostAlloc(&f, 64, cudaHostAllocMapped));
    *f = 0u;
    e->done_flag_h = f;
    CUDA_TRY_B(cudaHostGetDevicePointer(&e->done_flag_dev, f, 0));
  }
  if (e->writeback)
    CUDA_TRY_B(cudaMalloc(&e->controls_d, (size_t)e->D * noise_floats * sizeof(float)));
  if (world > 1)
  {
    CUDA_TRY_B(cudaMalloc(&e->rank_rec_d, (size_t)e->D * e->pstride * sizeof(float)));
    CUDA_TRY_B(cudaMalloc(&e->gather_d, (size_t)world * e->D * e->pstride * sizeof(float)));
    CUDA_TRY_B(cudaMalloc(&e->gather_hdr_d, (size_t)world * e->D * sizeof(float4)));
  }
  for (int i = 0; i < 4; i++)
    CUDA_TRY_B(cudaEventCreate(&e->ev[i]));

  if (e->nln)
    CUDA_TRY_B(cudaMalloc(&e->nln_d, noise_floats * sizeof(float)));
  if (e->colored)
  {
    // spectrum (the raw draw), time-domain buffer, tables and the reference's plan (colored_noise.cu:236-282)
    const size_t batch = (size_t)e->n_local * e->C;
    CUDA_TRY_B(cudaMalloc(&e->spec_alloc, (lead_floats + e->draw_local + 8) * sizeof(float)));
    CUDA_TRY_B(cudaMemsetAsync(e->spec_alloc, 0, (lead_floats + e->draw_local + 8) * sizeof(float), e->stream));
    e->spec_d = reinterpret_cast<float2*>(e->spec_alloc + lead_floats);
    CUDA_TRY_B(cudaMalloc(&e->time_d, batch * 2 * e->T * sizeof(float)));
    CUDA_TRY_B(cudaMalloc(&e->coeffs_d, (size_t)e->C * e->F * sizeof(float)));
    CUDA_TRY_B(cudaMalloc(&e->sigma_d, (size_t)e->C * sizeof(float)));
    CUDA_TRY_B(cudaMalloc(&e->decay_pow_d, (size_t)e->T * sizeof(float)));
    CUDA_TRY_B(cudaEventCreateWithFlags(&e->ev_rearr, cudaEventDisableTiming));
    const cufftResult fr = cufftPlan1d(&e->fft_plan, 2 * e->T, CUFFT_C2R, (int)batch);
    if (fr != CUFFT_SUCCESS)
      return bail(fail(MPPIB_ERR_CUDA, "cufftPlan1d(%d, C2R, %zu) failed: %d", 2 * e->T, batch, (int)fr));
    e->have_plan = true;
  }

  // own XORWOW draw: possible when a solve's block is a whole number of 8192-normal rounds (the states then sit at the
  // same place of every block and one fixed jump takes them from solve to solve). A rank slice that starts or ends
  // inside a round (e.g. 1024 x 100 normals per rank at 8 GPUs) is drawn in WINDOW mode: whole rounds around it, stores
  // predicated to the slice — the library fallback would re-seed 4096 subsequences on every solve (~200 us).
  const bool xw_aligned = (e->draw_local % 8192) == 0 && (e->draw_start % 8192ULL) == 0;
  if (!e->nln && !(desc->flags & MPPIB_FLAG_CURAND_HOST_API) && !getenv("MPPIB_CURAND_HOST_API") &&
      (e->draw_global % 8192ULL) == 0 && e->draw_local > 0 && (xw_aligned || !e->colored))
  {
    const unsigned long long w0 = e->draw_start / 8192ULL, w1 = (e->draw_start + e->draw_local + 8191ULL) / 8192ULL;
    e->xw_window = !xw_aligned;
    e->xw_first_round = w0;
    e->xw_lead = (unsigned)(e->draw_start - w0 * 8192ULL);
    const int rounds_local = (int)(w1 - w0);
    const unsigned long long rounds_global = e->draw_global / 8192ULL;
    int K = 1;
    for (int cand = 1; cand <= 64 && cand <= rounds_local; cand++)
      if (rounds_local % cand == 0 && (rounds_local / cand >= 4 || cand == 1))
        K = cand;
    e->xw_chunks = K;
    e->xw_rounds_per_chunk = rounds_local / K;
    const unsigned long long jump_draws = 2ULL * (rounds_global - (unsigned long long)e->xw_rounds_per_chunk);
    std::vector<uint32_t> tables;
    xorwow_nibble_tables(XorwowMatrix::power(jump_draws), tables);
    e->xw_jump_d = (uint32_t)(kXorwowWeyl * (uint32_t)(jump_draws & 0xffffffffULL));
    const size_t nstates = (size_t)K * kXorwowStreams;
    CUDA_TRY_B(cudaMalloc(&e->xw_states_d, nstates * 6 * sizeof(uint32_t)));
    CUDA_TRY_B(cudaMalloc(&e->xw_tables_d, tables.size() * sizeof(uint32_t)));
    CUDA_TRY_B(cudaMemcpyAsync(e->xw_tables_d, tables.data(), tables.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
                               e->stream));
    CUDA_TRY_B(cudaStreamSynchronize(e->stream));
    e->xw_enabled = true;
    if (getenv("MPPIB_DEBUG"))
      fprintf(stderr, "[mppib] create: engine %p states %p tables %p (%zu B)\n", (void*)e, (void*)e->xw_states_d,
              (void*)e->xw_tables_d, tables.size() * sizeof(uint32_t));
  }

  if (e->use_tma)
  {
    for (int i = 0; i < 2; i++)
    {
      int rc = make_tensor_map(*e, e->eps_buf[i], &e->tmap_buf[i]);
      if (rc != MPPIB_OK)
        return bail(rc);
    }
    e->tmap = e->tmap_buf[0];
  }
  {
    int rc = e->prepare(*e);
    if (rc != MPPIB_OK)
      return bail(rc);
  }

  // Controller::createAndSeedCUDARandomNumberGen (controller.cu:192-207): XORWOW, seed, offset 0
  if (curandCreateGenerator(&e->gen, CURAND_RNG_PSEUDO_DEFAULT) != CURAND_STATUS_SUCCESS)
    return bail(fail(MPPIB_ERR_CURAND, "curandCreateGenerator failed"));
  if (curandSetStream(e->gen, e->stream) != CURAND_STATUS_SUCCESS)
    return bail(fail(MPPIB_ERR_CURAND, "curandSetStream failed"));
  {
    int rc = mppib_seed(e, 0ULL, 0ULL);
    if (rc != MPPIB_OK)
      return bail(rc);
  }
  CUDA_TRY_B(cudaStreamSynchronize(e->stream));
#undef CUDA_TRY_B
  *out = e;
  return MPPIB_OK;
}

int mppib_destroy(mppib_engine* e)
{
  if (!e)
    return MPPIB_OK;
  if (getenv("MPPIB_DEBUG"))
    fprintf(stderr, "[mppib] destroy: engine %p\n", (void*)e);
  cudaSetDevice(e->desc.device);
  if (e->side_stream)
    cudaStreamSynchronize(e->side_stream);
  if (e->stream)
    cudaStreamSynchronize(e->stream);
  if (e->comm && g_nccl.CommDestroy)
    g_nccl.CommDestroy(e->comm);
  for (int r = 0; r < 8; r++)
    if (e->peer_opened[r])
      cudaIpcCloseMemHandle(e->peer_opened[r]);
  cudaFree(e->p2p_gather_d);
  if (e->gen)
    curandDestroyGenerator(e->gen);
  if (e->costmap_tex)
    cudaDestroyTextureObject(e->costmap_tex);
  if (e->costmap_array)
    cudaFreeArray(e->costmap_array);
  cudaFree(e->nn_theta_d);
  cudaFree(e->lstm_theta_d);
  cudaFree(e->elev_d);
  cudaFree(e->fb_gains_d);
  cudaFree(e->eval_states_d);
  cudaFree(e->eval_strides_d);
  cudaFree(e->eval_costs_d);
  cudaFree(e->nln_d);
  cudaFree(e->vis_idx_d);
  cudaFree(e->vis_opt_d);
  cudaFree(e->nom_d);
  cudaFree(e->nom_u_d);
  if (e->nom_h)
    cudaFreeHost(e->nom_h);
  cudaFree(e->vis_outputs_d);
  cudaFree(e->vis_costs_d);
  cudaFree(e->vis_crash_d);
  cudaFree(e->noise_alloc);
  cudaFree(e->noise_alloc2);
  cudaFree(e->costs_d);
  cudaFree(e->partials_d);
  cudaFree(e->headers_d);
  cudaFree(e->gather_hdr_d);
  cudaFree(e->controls_d);
  cudaFree(e->rank_rec_d);
  cudaFree(e->gather_d);
  cudaFree(e->result_d);
  cudaFree(e->weights_d);
  cudaFree(e->l2_flush_d);
  cudaFree(e->xw_states_d);
  cudaFree(e->xw_tables_d);
  if (e->have_plan)
    cufftDestroy(e->fft_plan);
  cudaFree(e->spec_alloc);
  cudaFree(e->time_d);
  cudaFree(e->coeffs_d);
  cudaFree(e->sigma_d);
  cudaFree(e->decay_pow_d);
  if (e->ev_rearr)
    cudaEventDestroy(e->ev_rearr);
  if (e->result_h)
    cudaFreeHost(e->result_h);
  if (e->done_flag_h)
    cudaFreeHost((void*)e->done_flag_h);
  cudaFree(e->k2_counter_d);
  for (int i = 0; i < 4; i++)
    if (e->ev[i])
      cudaEventDestroy(e->ev[i]);
  for (int i = 0; i < 2; i++)
  {
    if (e->ev_k1_done[i])
      cudaEventDestroy(e->ev_k1_done[i]);
    if (e->ev_gen_done[i])
      cudaEventDestroy(e->ev_gen_done[i]);
  }
  if (e->ev_last_gen)
    cudaEventDestroy(e->ev_last_gen);
  if (e->side_stream)
    cudaStreamDestroy(e->side_stream);
  if (e->own_stream && e->stream)
    cudaStreamDestroy(e->stream);
  cudaGetLastError();
  delete e;
  return MPPIB_OK;
}

int mppib_set_blob(mppib_engine* e, int which, const void* host, size_t nbytes)
{
  if (!e || !host)
    return fail(MPPIB_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(e->desc.device));
  // weights, the costmap texture and the LSTM blob are read by kernels of a solve still in flight (mppib_solve_async):
  // replacing them under it would be a use-after-free
  if (e->pending != 0 && (which == MPPIB_BLOB_NN_WEIGHTS || which == MPPIB_BLOB_LSTM_WEIGHTS || which == MPPIB_BLOB_COSTMAP ||
                          which == MPPIB_BLOB_ELEVATION_MAP))
    return fail(MPPIB_ERR_STATE, "mppib_set_blob(%d) while a solve is pending: call mppib_solve_wait first", which);
  switch (which)
  {
    case MPPIB_BLOB_DYN_PARAMS:
      if (nbytes != e->dyn_param_bytes)
        return fail(MPPIB_ERR_INVALID_ARG, "dynamics params: got %zu bytes, expected %zu", nbytes,
                    e->dyn_param_bytes);
      e->dyn_blob.assign((const unsigned char*)host, (const unsigned char*)host + nbytes);
      e->have_dyn = true;
      return MPPIB_OK;
    case MPPIB_BLOB_COST_PARAMS:
      if (nbytes != e->cost_param_bytes)
        return fail(MPPIB_ERR_INVALID_ARG, "cost params: got %zu bytes, expected %zu", nbytes, e->cost_param_bytes);
      e->cost_blob.assign((const unsigned char*)host, (const unsigned char*)host + nbytes);
      e->have_cost = true;
      return MPPIB_OK;
    case MPPIB_BLOB_SAMPLER_PARAMS:
    {
      if (nbytes != sizeof(mppib_gaussian_params))
        return fail(MPPIB_ERR_INVALID_ARG, "sampler params: got %zu bytes, expected %zu", nbytes,
                    sizeof(mppib_gaussian_params));
      mppib_gaussian_params sp;
      memcpy(&sp, host, sizeof(sp));
      if (e->D > 1 && !sp.use_same_noise_for_all_distributions)
        return fail(MPPIB_ERR_UNSUPPORTED,
                    "use_same_noise_for_all_distributions = false is not supported (Tube-MPPI default is true, "
                    "sampling_distribution.cuh:20)");
      for (int i = 0; i < e->D * e->C; i++)
        if (!(sp.std_dev[i] > 0.0f))
          return fail(MPPIB_ERR_INVALID_ARG, "std_dev[%d] must be positive", i);
      if (e->colored)
      {
        // frequency weights and sigma, computed like ColoredNoiseDistribution::generateSamples does on the host every
        // call (colored_noise.cu:294-338): fftfreq(2T), low-frequency cutoff, f^(-beta_c/2), theoretical std dev
        const int n2 = 2 * e->T, F = e->F, Cn = e->C;
        std::vector<float> sample_freq(F);
        for (int i = 0; i < F; i++)
          sample_freq[i] = i / (1.0f * n2);  // colored_noise.cuh:24-34
        const float cutoff_freq = fmaxf(sp.fmin, 1.0f / n2);
        std::vector<float> coeffs((size_t)Cn * F, 0.0f);  // Eigen MatrixXf(F, C) is column-major: [c][f]
        int smaller_index = 0;
        for (int i = 0; i < F; i++)
        {
          if (sample_freq[i] < cutoff_freq)
            smaller_index++;
          else if (smaller_index < F)
            for (int j = 0; j < smaller_index; j++)
            {
              sample_freq[j] = sample_freq[smaller_index];
              for (int k = 0; k < Cn; k++)
                coeffs[(size_t)k * F + j] = powf(sample_freq[smaller_index], -sp.exponents[k] / 2.0f);
            }
          for (int j = 0; j < Cn; j++)
            coeffs[(size_t)j * F + i] = powf(sample_freq[i], -sp.exponents[j] / 2.0f);
        }
        float sigma[MPPIB_MAX_CONTROL_DIM] = { 0 };
        for (int i = 0; i < Cn; i++)
        {
          for (int j = 1; j < F - 1; j++)
            sigma[i] += coeffs[(size_t)i * F + j] * coeffs[(size_t)i * F + j];
          const float last = coeffs[(size_t)i * F + F - 1] * ((1.0f + (n2 % 2)) / 2.0f);
          sigma[i] += last * last;
          sigma[i] = 2.0f * sqrtf(sigma[i]) / n2;
          if (!(sigma[i] > 0.0f) || !std::isfinite(sigma[i]))
            return fail(MPPIB_ERR_INVALID_ARG, "ColoredNoise: exponent %g gives a non-finite spectrum", sp.exponents[i]);
        }
        if (e->side_stream)
          CUDA_TRY(cudaStreamSynchronize(e->side_stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        CUDA_TRY(cudaMemcpy(e->coeffs_d, coeffs.data(), coeffs.size() * sizeof(float), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(e->sigma_d, sigma, Cn * sizeof(float), cudaMemcpyHostToDevice));
        colored_decay_table_kernel<<<(e->T + 127) / 128, 128, 0, e->stream>>>(e->decay_pow_d, e->T, sp.offset_decay_rate);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        e->prefetch_valid = false;  // a prefetched block was shaped with the old table
      }
      if (e->nln && e->have_sampler && memcmp(sp.std_dev, e->sampler.std_dev, sizeof(sp.std_dev)) != 0)
        e->prefetch_valid = false;  // a prefetched block drew its log-normal planes with the old std dev
      e->sampler = sp;
      e->have_sampler = true;
      return MPPIB_OK;
    }
    case MPPIB_BLOB_NN_WEIGHTS:
    {
      if (e->desc.dynamics_id != MPPIB_DYN_AUTORALLY_NN)
        return fail(MPPIB_ERR_INVALID_ARG, "NN weights given to a non-NN dynamics");
      if (nbytes != MPPIB_AR_NN_NUM_PARAMS * sizeof(float))
        return fail(MPPIB_ERR_INVALID_ARG, "NN weights: got %zu bytes, expected %zu", nbytes,
                    MPPIB_AR_NN_NUM_PARAMS * sizeof(float));
      const float* w = (const float*)host;
      for (int i = 0; i < MPPIB_AR_NN_NUM_PARAMS; i++)
        if (!std::isfinite(w[i]))  // fnn_helper.cu:244-247 asserts finiteness
          return fail(MPPIB_ERR_INVALID_ARG, "NN weight %d is not finite", i);
      if (!e->nn_theta_d)
        CUDA_TRY(cudaMalloc(&e->nn_theta_d, nbytes));
      CUDA_TRY(cudaMemcpyAsync(e->nn_theta_d, host, nbytes, cudaMemcpyHostToDevice, e->stream));
      CUDA_TRY(cudaStreamSynchronize(e->stream));
      e->nn_theta_h.assign(w, w + MPPIB_AR_NN_NUM_PARAMS);
      return MPPIB_OK;
    }
    case MPPIB_BLOB_LSTM_WEIGHTS:
    {
      if (e->desc.dynamics_id != MPPIB_DYN_RACER_LSTM)
        return fail(MPPIB_ERR_INVALID_ARG, "LSTM weights given to a dynamics without an LSTM");
      const int H = e->desc.model_dims[0], L1 = e->desc.model_dims[1];
      const size_t expect = (size_t)MPPIB_RACER_LSTM_NUM_PARAMS(H, L1) * sizeof(float);
      if (nbytes != expect)
        return fail(MPPIB_ERR_INVALID_ARG, "LSTM weights: got %zu bytes, expected %zu (H = %d, head width %d)", nbytes,
                    expect, H, L1);
      const float* w = (const float*)host;
      for (size_t i = 0; i < nbytes / sizeof(float); i++)
        if (!std::isfinite(w[i]))
          return fail(MPPIB_ERR_INVALID_ARG, "LSTM weight %zu is not finite", i);
      if (!e->lstm_theta_d)
        CUDA_TRY(cudaMalloc(&e->lstm_theta_d, nbytes));
      CUDA_TRY(cudaMemcpyAsync(e->lstm_theta_d, host, nbytes, cudaMemcpyHostToDevice, e->stream));
      CUDA_TRY(cudaStreamSynchronize(e->stream));
      e->lstm_theta_h.assign(w, w + nbytes / sizeof(float));
      e->have_lstm = true;
      return MPPIB_OK;
    }
    case MPPIB_BLOB_ELEVATION_MAP:
    {
      // TwoDTextureHelper<float>::updateTexture / updateOrigin / updateRotation / updateResolution / enableTexture +
      // copyToDevice of the RACER models' map 0 (racer_dubins_elevation.cuh: tex_helper_), in one blob
      if (e->desc.dynamics_id != MPPIB_DYN_RACER_LSTM)
        return fail(MPPIB_ERR_INVALID_ARG, "elevation map given to a dynamics without one");
      if (nbytes < sizeof(mppib_elevation_map_header))
        return fail(MPPIB_ERR_INVALID_ARG, "elevation map: %zu bytes is smaller than its header", nbytes);
      mppib_elevation_map_header h;
      memcpy(&h, host, sizeof(h));
      if (h.width < 2 || h.height < 2 || h.width > 16384 || h.height > 16384)
        return fail(MPPIB_ERR_INVALID_ARG, "elevation map: extent %d x %d (need 2 .. 16384 cells per side)", h.width, h.height);
      const size_t cells = (size_t)h.width * h.height;
      if (nbytes != sizeof(h) + cells * sizeof(float))
        return fail(MPPIB_ERR_INVALID_ARG, "elevation map: got %zu bytes, expected %zu (header + %d x %d floats)", nbytes,
                    sizeof(h) + cells * sizeof(float), h.width, h.height);
      for (int i = 0; i < 3; i++)
        if (!std::isfinite(h.origin[i]) || !std::isfinite(h.resolution[i]) || h.resolution[i] == 0.0f)
          return fail(MPPIB_ERR_INVALID_ARG, "elevation map: origin / resolution component %d is not usable", i);
      for (int i = 0; i < 9; i++)
        if (!std::isfinite(h.rotations[i]))
          return fail(MPPIB_ERR_INVALID_ARG, "elevation map: rotation entry %d is not finite", i);
      // the values themselves may be NaN (unobserved cells): the model's own isfinite guards handle that (racer_dubins.cu:414-425)
      if (cells > e->elev_capacity)
      {
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        cudaFree(e->elev_d);
        e->elev_d = nullptr;
        e->elev_capacity = 0;
        CUDA_TRY(cudaMalloc(&e->elev_d, cells * sizeof(float)));
        e->elev_capacity = cells;
      }
      CUDA_TRY(cudaMemcpyAsync(e->elev_d, (const char*)host + sizeof(h), cells * sizeof(float), cudaMemcpyHostToDevice,
                               e->stream));
      CUDA_TRY(cudaStreamSynchronize(e->stream));
      e->elev_hdr = h;
      e->elev_h.assign((const unsigned char*)host, (const unsigned char*)host + nbytes);
      return MPPIB_OK;
    }
    case MPPIB_BLOB_COSTMAP:
    {
      if (e->desc.cost_id != MPPIB_COST_AR_STANDARD)
        return fail(MPPIB_ERR_INVALID_ARG, "costmap given to a cost without a map");
      if (!e->have_cost)
        return fail(MPPIB_ERR_STATE, "set MPPIB_BLOB_COST_PARAMS (map_width/map_height) before the costmap");
      mppib_ar_standard_cost_params cp;
      memcpy(&cp, e->cost_blob.data(), sizeof(cp));
      const size_t expect = (size_t)cp.map_width * cp.map_height * 4 * sizeof(float);
      if (cp.map_width <= 0 || cp.map_height <= 0 || nbytes != expect)
        return fail(MPPIB_ERR_INVALID_ARG, "costmap: got %zu bytes, expected %zu (%d x %d float4)", nbytes, expect,
                    cp.map_width, cp.map_height);
      if (e->costmap_tex)
      {
        cudaDestroyTextureObject(e->costmap_tex);
        e->costmap_tex = 0;
      }
      if (e->costmap_array)
      {
        cudaFreeArray(e->costmap_array);
        e->costmap_array = nullptr;
      }
      // ar_standard_cost.cu:101-176: float4 array, clamp, point filter, element read, normalised coordinates
      cudaChannelFormatDesc ch = cudaCreateChannelDesc(32, 32, 32, 32, cudaChannelFormatKindFloat);
      CUDA_TRY(cudaMallocArray(&e->costmap_array, &ch, cp.map_width, cp.map_height));
      CUDA_TRY(cudaMemcpy2DToArrayAsync(e->costmap_array, 0, 0, host, (size_t)cp.map_width * 16,
                                        (size_t)cp.map_width * 16, cp.map_height, cudaMemcpyHostToDevice, e->stream));
      CUDA_TRY(cudaStreamSynchronize(e->stream));
      cudaResourceDesc res;
      memset(&res, 0, sizeof(res));
      res.resType = cudaResourceTypeArray;
      res.res.array.array = e->costmap_array;
      cudaTextureDesc tex;
      memset(&tex, 0, sizeof(tex));
      tex.addressMode[0] = cudaAddressModeClamp;
      tex.addressMode[1] = cudaAddressModeClamp;
      tex.filterMode = cudaFilterModePoint;
      tex.readMode = cudaReadModeElementType;
      tex.normalizedCoords = 1;
      CUDA_TRY(cudaCreateTextureObject(&e->costmap_tex, &res, &tex, nullptr));
      return MPPIB_OK;
    }
    default:
      return fail(MPPIB_ERR_INVALID_ARG, "unknown blob kind %d", which);
  }
}

int mppib_set_solver(mppib_engine* e, float dt, float lambda, float alpha)
{
  if (!e)
    return fail(MPPIB_ERR_INVALID_ARG, "null engine");
  if (!(dt > 0.0f) || !(lambda > 0.0f))
    return fail(MPPIB_ERR_INVALID_ARG, "dt and lambda must be positive");
  e->dt = dt;
  e->lambda = lambda;
  e->alpha = alpha;
  return MPPIB_OK;
}

int mppib_seed(mppib_engine* e, unsigned long long seed, unsigned long long offset)
{
  if (!e)
    return fail(MPPIB_ERR_INVALID_ARG, "null engine");
  CUDA_TRY(cudaSetDevice(e->desc.device));
  if (e->side_stream)
    CUDA_TRY(cudaStreamSynchronize(e->side_stream));  // a prefetch may still be using the generator
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  CURAND_TRY(curandSetPseudoRandomGeneratorSeed(e->gen, seed));
  CURAND_TRY(curandSetGeneratorOffset(e->gen, 0ULL));
  e->seed = seed;
  e->rng_offset = offset;
  e->xw_pos = mppib_engine::kNoPos;
  e->curand_pos = 0;  // the library generator sits at element 0 of the new stream
  e->prefetch_valid = false;
  return MPPIB_OK;
}

int mppib_get_rng_offset(mppib_engine* e, unsigned long long* offset)
{
  if (!e || !offset)
    return fail(MPPIB_ERR_INVALID_ARG, "null argument");
  *offset = e->rng_offset;
  return MPPIB_OK;
}

int mppib_burn_draws(mppib_engine* e, int n)
{
  if (!e || n < 0)
    return fail(MPPIB_ERR_INVALID_ARG, "bad argument");
  // skipping is free for a counter-positioned stream: just move the absolute offset
  e->rng_offset += (unsigned long long)n * e->draw_global;  // generators are re-positioned lazily by gen_draw
  e->prefetch_valid = false;
  return MPPIB_OK;
}

int mppib_comm_unique_id(void* unique_id_128)
{
  if (!unique_id_128)
    return fail(MPPIB_ERR_INVALID_ARG, "null argument");
  if (!g_nccl.load())
    return fail(MPPIB_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", dlerror());
  NcclUniqueId id;
  int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0)
    return fail(MPPIB_ERR_NCCL, "ncclGetUniqueId failed (%d)", rc);
  memcpy(unique_id_128, &id, sizeof(id));
  return MPPIB_OK;
}

int mppib_comm_init(mppib_engine* e, const void* unique_id_128)
{
  if (!e || !unique_id_128)
    return fail(MPPIB_ERR_INVALID_ARG, "null argument");
  if (e->desc.world_size <= 1)
    return MPPIB_OK;
  if (!g_nccl.load())
    return fail(MPPIB_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", dlerror());
  CUDA_TRY(cudaSetDevice(e->desc.device));
  NcclUniqueId id;
  memcpy(&id, unique_id_128, sizeof(id));
  int rc = g_nccl.CommInitRank(&e->comm, e->desc.world_size, id, e->desc.rank);
  if (rc != 0)
    return fail(MPPIB_ERR_NCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  return MPPIB_OK;
}

int mppib_comm_p2p_handle(mppib_engine* e, void* handle_64)
{
  if (!e || !handle_64)
    return fail(MPPIB_ERR_INVALID_ARG, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  const int world = e->desc.world_size;
  if (world < 2 || world > 8)
    return fail(MPPIB_ERR_UNSUPPORTED, "peer-memory exchange is built for 2..8 ranks");
  CUDA_TRY(cudaSetDevice(e->desc.device));
  if (!e->p2p_gather_d)
  {
    const size_t floats = (size_t)2 * world * e->D * e->pstride + 2 * world + 16;
    CUDA_TRY(cudaMalloc(&e->p2p_gather_d, floats * sizeof(float)));
    CUDA_TRY(cudaMemset(e->p2p_gather_d, 0, floats * sizeof(float)));
  }
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, e->p2p_gather_d));
  memcpy(handle_64, &h, 64);
  return MPPIB_OK;
}

int mppib_comm_p2p_open(mppib_engine* e, const void* handles)
{
  if (!e || !handles)
    return fail(MPPIB_ERR_INVALID_ARG, "null argument");
  if (!e->p2p_gather_d)
    return fail(MPPIB_ERR_STATE, "call mppib_comm_p2p_handle first");
  const int world = e->desc.world_size;
  CUDA_TRY(cudaSetDevice(e->desc.device));
  const size_t gather_floats = (size_t)2 * world * e->D * e->pstride;
  for (int r = 0; r < world; r++)
  {
    float* base = nullptr;
    if (r == e->desc.rank)
      base = e->p2p_gather_d;
    else
    {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char*)handles + (size_t)r * 64, 64);
      void* p = nullptr;
      cudaError_t rc = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
      if (rc != cudaSuccess)
      {
        cudaGetLastError();
        return fail(MPPIB_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s — keep the NCCL path", r,
                    cudaGetErrorString(rc));
      }
      e->peer_opened[r] = p;
      base = (float*)p;
    }
    e->peers.gather[r] = base;
    e->peers.flags[r] = reinterpret_cast<unsigned*>(base + gather_floats);
  }
  e->p2p = true;
  e->p2p_opened = true;
  return MPPIB_OK;
}

int mppib_set_noise(mppib_engine* e, const float* host_eps, size_t count)
{
  if (!e || !host_eps)
    return fail(MPPIB_ERR_INVALID_ARG, "null argument");
  if (count != (size_t)e->n_local * e->TC)
    return fail(MPPIB_ERR_INVALID_ARG, "noise count %zu != n_local*T*C = %zu", count, (size_t)e->n_local * e->TC);
  CUDA_TRY(cudaSetDevice(e->desc.device));
  CUDA_TRY(cudaMemcpyAsync(e->eps_d, host_eps, count * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MPPIB_OK;
}

int mppib_draw_noise(mppib_engine* e)
{
  if (!e)
    return fail(MPPIB_ERR_INVALID_ARG, "null engine");
  CUDA_TRY(cudaSetDevice(e->desc.device));
  int rc = draw_noise(*e, e->colored_offset_t);
  if (rc != MPPIB_OK)
    return rc;
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MPPIB_OK;
}

int mppib_rollout_only(mppib_engine* e, const float* x0, const float* U_in, int optimization_stride, int iteration_num)
{
  int rc = check_ready(e);
  if (rc != MPPIB_OK)
    return rc;
  if (!x0 || !U_in)
    return fail(MPPIB_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(e->desc.device));
  rc = e->launch_rollout(*e, x0, U_in, optimization_stride, iteration_num);
  if (rc != MPPIB_OK)
    return rc;
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  e->solved_once = true;
  return MPPIB_OK;
}

int mppib_reduce_only(mppib_engine* e, float* U_out, mppib_solve_stats* stats)
{
  int rc = check_ready(e);
  if (rc != MPPIB_OK)
    return rc;
  if (!e->solved_once)
    return fail(MPPIB_ERR_STATE, "no rollout has been run yet");
  CUDA_TRY(cudaSetDevice(e->desc.device));
  rc = launch_combine(*e, false);
  if (rc != MPPIB_OK)
    return rc;
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  read_result(*e, U_out, stats);
  return MPPIB_OK;
}

static int enqueue_solve(mppib_engine* e, const float* x0, const float* U_in, int optimization_stride,
                         int iteration_num)
{
  if (e->timing)
    CUDA_TRY(cudaEventRecord(e->ev[0], e->stream));
  int rc = draw_noise(*e, optimization_stride);
  if (rc != MPPIB_OK)
    return rc;
  if (e->l2_flush_d)
    CUDA_TRY(cudaMemsetAsync(e->l2_flush_d, 0, e->l2_flush_bytes, e->stream));
  if (e->timing)
    CUDA_TRY(cudaEventRecord(e->ev[1], e->stream));
  rc = e->launch_rollout(*e, x0, U_in, optimization_stride, iteration_num);
  if (rc != MPPIB_OK)
    return rc;
  rc = prefetch_next(*e);
  if (rc != MPPIB_OK)
    return rc;
  if (e->timing)
    CUDA_TRY(cudaEventRecord(e->ev[2], e->stream));
  rc = launch_combine(*e, /*after_k1=*/!e->timing);  // timing mode records an event between K1 and K2: no PDL then
  if (rc != MPPIB_OK)
    return rc;
  if (e->prefetch_enabled)
  {  // "the K1 that read eps_buf[cur_buf] is done" — recorded after K2 so nothing sits between K1 and its PDL dependent
    CUDA_TRY(cudaEventRecord(e->ev_k1_done[e->cur_buf], e->stream));
    e->k1_recorded[e->cur_buf] = true;
  }
  if (e->timing)
    CUDA_TRY(cudaEventRecord(e->ev[3], e->stream));
  e->pending++;
  return MPPIB_OK;
}

// Blocks until the last enqueued solve is complete. Fast path: poll the mapped word K2's last block writes after the
// result record (the record travels the same way, so it is visible when the flag is); every few thousand polls the
// stream is queried so that a failed launch cannot spin forever, and timing mode uses a full synchronize (events).
static int wait_for_stream(mppib_engine* e)
{
  if (e->spin_wait && !e->timing && e->flag_armed && e->solve_seq != 0)
  {
    const unsigned want = e->solve_seq;
    for (unsigned spins = 1;; spins++)
    {
      if (*e->done_flag_h == want)
        return MPPIB_OK;
      if ((spins & 0x3fff) == 0)
      {
        cudaError_t q = cudaStreamQuery(e->stream);
        if (q == cudaSuccess)
          break;  // stream drained (the flag write is then visible as well)
        if (q != cudaErrorNotReady)
        {
          cudaGetLastError();
          return fail(MPPIB_ERR_CUDA, "solve failed: %s", cudaGetErrorString(q));
        }
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
  }
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MPPIB_OK;
}

static int wait_solve(mppib_engine* e, float* U_out, mppib_solve_stats* stats)
{
  int rcw = wait_for_stream(e);
  if (rcw != MPPIB_OK)
    return rcw;
  e->pending = 0;
  e->timing_valid = e->timing;
  if (e->timing)
  {
    float ms[4];
    if (cudaEventElapsedTime(&ms[0], e->ev[0], e->ev[1]) == cudaSuccess &&
        cudaEventElapsedTime(&ms[1], e->ev[1], e->ev[2]) == cudaSuccess &&
        cudaEventElapsedTime(&ms[2], e->ev[2], e->ev[3]) == cudaSuccess &&
        cudaEventElapsedTime(&ms[3], e->ev[0], e->ev[3]) == cudaSuccess)
    {
      for (int i = 0; i < 4; i++)
        e->acc_ms[i] += ms[i];
      e->acc_n++;
    }
  }
  e->solved_once = true;
  read_result(*e, U_out, stats);
  return MPPIB_OK;
}

int mppib_solve(mppib_engine* e, const float* x0, const float* U_in, int optimization_stride, int iteration_num,
                float* U_out, mppib_solve_stats* stats)
{
  int rc = check_ready(e);
  if (rc != MPPIB_OK)
    return rc;
  if (!x0 || !U_in || !U_out)
    return fail(MPPIB_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(e->desc.device));
  rc = enqueue_solve(e, x0, U_in, optimization_stride, iteration_num);
  if (rc != MPPIB_OK)
    return rc;
  return wait_solve(e, U_out, stats);
}

// One iteration of Controller::computeControl (controllers/MPPI/mppi_controller.cu:151-241) as ONE call: the solve, then the
// host tail on its result with the parameter blobs the engine was given — smoothControlTrajectory (controller.cuh:557-586) and
// computeStateTrajectory (:643-663) through the library's host twins.
int mppib_compute_control(mppib_engine* e, const float* x0, float* U_inout, int optimization_stride, int iteration_num,
                          const float* control_history, float* states, float* outputs, mppib_solve_stats* stats)
{
  int rc = check_ready(e);
  if (rc != MPPIB_OK)
    return rc;
  if (!x0 || !U_inout || (states == nullptr) != (outputs == nullptr))
    return fail(MPPIB_ERR_INVALID_ARG, "null argument (states and outputs go together)");
  const int dyn = e->desc.dynamics_id;
  if (states && dyn >= MPPIB_USER_ID_BASE)
    return fail(MPPIB_ERR_UNSUPPORTED, "user dynamics have no host twin in the library: roll the state forward in the caller");
  CUDA_TRY(cudaSetDevice(e->desc.device));
  rc = enqueue_solve(e, x0, U_inout, optimization_stride, iteration_num);
  if (rc != MPPIB_OK)
    return rc;
  rc = wait_solve(e, U_inout, stats);
  if (rc != MPPIB_OK)
    return rc;
  for (int d = 0; d < e->D; d++)
  {
    float* u = U_inout + (size_t)d * e->TC;
    if (control_history)
      mppib_host_smooth_controls(u, control_history, e->T, e->C);
    if (!states)
      continue;
    float* st = states + (size_t)d * e->T * e->S;
    float* out = outputs + (size_t)d * e->T * e->O;
    if (dyn == MPPIB_DYN_RACER_LSTM)
    {
      mppib_host_lstm net{ e->lstm_theta_h.data(), e->desc.model_dims[0], e->desc.model_dims[1], nullptr, nullptr,
                           e->elev_h.empty() ? nullptr : reinterpret_cast<const mppib_elevation_map_header*>(e->elev_h.data()) };
      rc = mppib_host_output_trajectory_lstm(e->dyn_blob.data(), &net, x0 + (size_t)d * e->S, u, e->T, e->dt, st, out);
    }
    else
      rc = mppib_host_output_trajectory(dyn, e->dyn_blob.data(), e->nn_theta_h.empty() ? nullptr : e->nn_theta_h.data(),
                                        x0 + (size_t)d * e->S, u, e->T, e->dt, st, out);
    if (rc != MPPIB_OK)
      return fail(rc, "host roll-forward failed for dynamics %d", dyn);
  }
  return MPPIB_OK;
}

int mppib_solve_async(mppib_engine* e, const float* x0, const float* U_in, int optimization_stride, int iteration_num)
{
  int rc = check_ready(e);
  if (rc != MPPIB_OK)
    return rc;
  if (!x0 || !U_in)
    return fail(MPPIB_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(e->desc.device));
  return enqueue_solve(e, x0, U_in, optimization_stride, iteration_num);
}

int mppib_solve_wait(mppib_engine* e, float* U_out, mppib_solve_stats* stats)
{
  if (!e)
    return fail(MPPIB_ERR_INVALID_ARG, "null engine");
  if (e->pending == 0)
    return fail(MPPIB_ERR_STATE, "no solve in flight");
  CUDA_TRY(cudaSetDevice(e->desc.device));
  return wait_solve(e, U_out, stats);
}

int mppib_set_tsallis(mppib_engine* e, float gamma, float r)
{
  if (!e)
    return fail(MPPIB_ERR_INVALID_ARG, "null engine");
  if (gamma != 0.0f && r != 0.0f)
  {
    if (!e->controls_d)
      return fail(MPPIB_ERR_STATE, "Tsallis weights reduce the written-back controls: create the engine with "
                                   "MPPIB_FLAG_WRITEBACK_CONTROLS");
    if (e->desc.world_size != 1)
      return fail(MPPIB_ERR_UNSUPPORTED, "Tsallis weights are built for one rank");
    if (r == 1.0f || !(gamma > 0.0f))
      return fail(MPPIB_ERR_INVALID_ARG, "Tsallis weights need gamma > 0 and r != 1");
  }
  e->tsallis_gamma = gamma;
  e->tsallis_r = r;
  return MPPIB_OK;
}

int mppib_set_rmppi(mppib_engine* e, float value_func_threshold, const float* feedback_gains)
{
  if (!e)
    return fail(MPPIB_ERR_INVALID_ARG, "null engine");
  if (!e->rmppi)
    return fail(MPPIB_ERR_STATE, "engine was not created with MPPIB_FLAG_RMPPI");
  CUDA_TRY(cudaSetDevice(e->desc.device));
  e->value_func_threshold = value_func_threshold;
  const size_t n = (size_t)e->T * e->S * e->C;
  if (feedback_gains)
  {
    for (size_t i = 0; i < n; i++)
      if (!std::isfinite(feedback_gains[i]))
        return fail(MPPIB_ERR_INVALID_ARG, "feedback gain %zu is not finite", i);
    if (!e->fb_gains_d)
      CUDA_TRY(cudaMalloc(&e->fb_gains_d, n * sizeof(float)));
    CUDA_TRY(cudaMemcpyAsync(e->fb_gains_d, feedback_gains, n * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
  }
  else if (e->fb_gains_d)
  {
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    cudaFree(e->fb_gains_d);
    e->fb_gains_d = nullptr;
  }
  return MPPIB_OK;
}

int mppib_init_eval(mppib_engine* e, const float* candidates, const int* strides, int num_candidates,
                    int samples_per_candidate, const float* U_nominal, int optimization_stride, float* costs_out)
{
  int rc = check_ready(e);
  if (rc != MPPIB_OK)
    return rc;
  if (!candidates || !strides || !U_nominal || !costs_out || num_candidates <= 0 || samples_per_candidate <= 0)
    return fail(MPPIB_ERR_INVALID_ARG, "bad argument");
  if (e->desc.world_size != 1)
    return fail(MPPIB_ERR_UNSUPPORTED, "init-eval runs on one rank (a few hundred rollouts)");
  if (samples_per_candidate > e->n_local || (long)num_candidates * samples_per_candidate > e->N)
    return fail(MPPIB_ERR_INVALID_ARG, "(number of candidates) * (samples per candidate) cannot exceed NUM_ROLLOUTS");
  CUDA_TRY(cudaSetDevice(e->desc.device));
  const int total = num_candidates * samples_per_candidate;
  if (total > e->eval_capacity || !e->eval_states_d)
  {
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    cudaFree(e->eval_states_d);
    cudaFree(e->eval_strides_d);
    cudaFree(e->eval_costs_d);
    e->eval_states_d = nullptr;
    e->eval_capacity = 0;
    CUDA_TRY(cudaMalloc(&e->eval_states_d, (size_t)total * e->S * sizeof(float)));
    CUDA_TRY(cudaMalloc(&e->eval_strides_d, (size_t)total * sizeof(int)));
    CUDA_TRY(cudaMalloc(&e->eval_costs_d, (size_t)total * sizeof(float)));
    e->eval_capacity = total;
  }
  CUDA_TRY(cudaMemcpyAsync(e->eval_states_d, candidates, (size_t)num_candidates * e->S * sizeof(float),
                           cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->eval_strides_d, strides, (size_t)num_candidates * sizeof(int), cudaMemcpyHostToDevice,
                           e->stream));
  rc = draw_noise(*e, optimization_stride);  // sampler_->generateSamples(stride, 0, gen_) (:595)
  if (rc != MPPIB_OK)
    return rc;
  rc = e->init_eval(*e, e->eval_states_d, e->eval_strides_d, num_candidates, samples_per_candidate, U_nominal,
                    optimization_stride);
  if (rc != MPPIB_OK)
    return rc;
  if (e->prefetch_enabled)
  {  // the kernel above read eps_buf[cur_buf]: later draws into that buffer must wait for it
    CUDA_TRY(cudaEventRecord(e->ev_k1_done[e->cur_buf], e->stream));
    e->k1_recorded[e->cur_buf] = true;
  }
  CUDA_TRY(cudaMemcpyAsync(costs_out, e->eval_costs_d, (size_t)total * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MPPIB_OK;
}

int mppib_sample_trajectories(mppib_engine* e, const float* x0, const float* U_nominal, int distribution,
                              const int* sample_idx, int n, const float* U_opt, float* outputs, float* costs, int* crash)
{
  int rc = check_ready(e);
  if (rc != MPPIB_OK)
    return rc;
  if (!x0 || !U_nominal || !sample_idx || n <= 0 || !outputs || !costs || !crash)
    return fail(MPPIB_ERR_INVALID_ARG, "bad argument");
  if (distribution < 0 || distribution >= e->D)
    return fail(MPPIB_ERR_INVALID_ARG, "distribution %d out of range [0, %d)", distribution, e->D);
  if (!e->writeback || !e->controls_d)
    return fail(MPPIB_ERR_STATE, "sampled trajectories re-roll the written-back controls: create the engine with "
                                 "MPPIB_FLAG_WRITEBACK_CONTROLS");
  if (e->rmppi)
    return fail(MPPIB_ERR_UNSUPPORTED, "sampled trajectories are built for the Vanilla / Tube / Colored rollouts");
  if (!e->solved_once)
    return fail(MPPIB_ERR_STATE, "no solve has been run yet");
  if (e->pending)
    return fail(MPPIB_ERR_STATE, "a solve is in flight (mppib_solve_wait first)");
  bool have_opt = false;
  for (int i = 0; i < n; i++)
  {
    if (sample_idx[i] < -1 || sample_idx[i] >= e->n_local)
      return fail(MPPIB_ERR_INVALID_ARG, "sample index %d (entry %d) outside [-1, %d)", sample_idx[i], i, e->n_local);
    have_opt = have_opt || sample_idx[i] < 0;
  }
  if (have_opt && !U_opt)
    return fail(MPPIB_ERR_INVALID_ARG, "index -1 needs U_opt");
  for (int i = 0; i < e->S; i++)
    if (!std::isfinite(x0[i]))
      return fail(MPPIB_ERR_INVALID_ARG, "x0[%d] is not finite", i);
  CUDA_TRY(cudaSetDevice(e->desc.device));
  if (n > e->vis_capacity)
  {
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    cudaFree(e->vis_idx_d);
    cudaFree(e->vis_outputs_d);
    cudaFree(e->vis_costs_d);
    cudaFree(e->vis_crash_d);
    e->vis_idx_d = nullptr;
    e->vis_outputs_d = e->vis_costs_d = nullptr;
    e->vis_crash_d = nullptr;
    e->vis_capacity = 0;
    CUDA_TRY(cudaMalloc(&e->vis_idx_d, (size_t)n * sizeof(int)));
    CUDA_TRY(cudaMalloc(&e->vis_outputs_d, (size_t)n * e->T * e->O * sizeof(float)));
    CUDA_TRY(cudaMalloc(&e->vis_costs_d, (size_t)n * (e->T + 1) * sizeof(float)));
    CUDA_TRY(cudaMalloc(&e->vis_crash_d, (size_t)n * e->T * sizeof(int)));
    e->vis_capacity = n;
  }
  if (have_opt && !e->vis_opt_d)
    CUDA_TRY(cudaMalloc(&e->vis_opt_d, (size_t)e->TC * sizeof(float)));
  CUDA_TRY(cudaMemcpyAsync(e->vis_idx_d, sample_idx, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  if (have_opt)
    CUDA_TRY(cudaMemcpyAsync(e->vis_opt_d, U_opt, (size_t)e->TC * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  rc = e->sampled_traj(*e, x0, U_nominal, distribution, n, have_opt);
  if (rc != MPPIB_OK)
    return rc;
  CUDA_TRY(cudaMemcpyAsync(outputs, e->vis_outputs_d, (size_t)n * e->T * e->O * sizeof(float), cudaMemcpyDeviceToHost,
                           e->stream));
  CUDA_TRY(cudaMemcpyAsync(costs, e->vis_costs_d, (size_t)n * (e->T + 1) * sizeof(float), cudaMemcpyDeviceToHost,
                           e->stream));
  CUDA_TRY(cudaMemcpyAsync(crash, e->vis_crash_d, (size_t)n * e->T * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MPPIB_OK;
}

// Device-side host tail (controller.cuh:557-586, 643-663): see nominal_traj_kernel.
int mppib_nominal_trajectory(mppib_engine* e, const float* x0, const float* U, const float* control_history,
                             float* U_smoothed, float* states, float* outputs)
{
  int rc = check_ready(e);
  if (rc != MPPIB_OK)
    return rc;
  if (!x0 || !states || !outputs)
    return fail(MPPIB_ERR_INVALID_ARG, "null argument");
  if (!U && !e->solved_once && !e->pending)
    return fail(MPPIB_ERR_STATE, "U == NULL rolls out the last solve's result, and no solve has been run yet");
  if (e->T < 2)
    return fail(MPPIB_ERR_INVALID_ARG, "needs at least two time steps");
  for (int i = 0; i < e->D * e->S; i++)
    if (!std::isfinite(x0[i]))
      return fail(MPPIB_ERR_INVALID_ARG, "x0[%d] is not finite", i);
  CUDA_TRY(cudaSetDevice(e->desc.device));
  const size_t n_u = (size_t)e->D * e->TC, n_s = (size_t)e->D * e->T * e->S, n_o = (size_t)e->D * e->T * e->O;
  if (!e->nom_d)
  {
    CUDA_TRY(cudaMalloc(&e->nom_d, (n_u + n_s + n_o) * sizeof(float)));
    CUDA_TRY(cudaHostAlloc(&e->nom_h, (n_u + n_s + n_o) * sizeof(float), cudaHostAllocDefault));
  }
  const float* u_src = e->result_d + kPartialHeader;  // the optimised sequence where K2 / KX left it
  int u_stride = e->pstride;
  if (U)
  {
    if (!e->nom_u_d)
      CUDA_TRY(cudaMalloc(&e->nom_u_d, n_u * sizeof(float)));
    CUDA_TRY(cudaMemcpyAsync(e->nom_u_d, U, n_u * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    u_src = e->nom_u_d;
    u_stride = e->TC;
  }
  rc = e->nominal_traj(*e, x0, u_src, u_stride, control_history);
  if (rc != MPPIB_OK)
    return rc;
  CUDA_TRY(cudaMemcpyAsync(e->nom_h, e->nom_d, (n_u + n_s + n_o) * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  if (U_smoothed)
    memcpy(U_smoothed, e->nom_h, n_u * sizeof(float));
  memcpy(states, e->nom_h + n_u, n_s * sizeof(float));
  memcpy(outputs, e->nom_h + n_u + n_s, n_o * sizeof(float));
  return MPPIB_OK;
}

int mppib_set_option(mppib_engine* e, int option, long long value)
{
  if (e && option == MPPIB_OPT_P2P_ENABLE)
  {
    if (value != 0 && !e->p2p_opened)
      return fail(MPPIB_ERR_STATE, "mppib_comm_p2p_open has not succeeded on this rank");
    if (value == 0 && !e->comm)
      return fail(MPPIB_ERR_STATE, "no NCCL communicator to fall back to (mppib_comm_init)");
    e->p2p = value != 0;
    return MPPIB_OK;
  }
  if (e && option == MPPIB_OPT_COLORED_OFFSET_T)
  {
    if (value < 0 || value >= 2 * e->T)
      return fail(MPPIB_ERR_INVALID_ARG, "offset_t out of range");
    e->colored_offset_t = (int)value;
    return MPPIB_OK;
  }
  if (!e)
    return fail(MPPIB_ERR_INVALID_ARG, "null engine");
  CUDA_TRY(cudaSetDevice(e->desc.device));
  switch (option)
  {
    case MPPIB_OPT_L2_FLUSH_BYTES:
      CUDA_TRY(cudaStreamSynchronize(e->stream));
      if (e->l2_flush_d)
      {
        cudaFree(e->l2_flush_d);
        e->l2_flush_d = nullptr;
        e->l2_flush_bytes = 0;
      }
      if (value > 0)
      {
        CUDA_TRY(cudaMalloc(&e->l2_flush_d, (size_t)value));
        e->l2_flush_bytes = (size_t)value;
      }
      return MPPIB_OK;
  }
  return fail(MPPIB_ERR_INVALID_ARG, "unknown option %d", option);
}

int mppib_get_costs(mppib_engine* e, float* host_costs)
{
  if (!e || !host_costs)
    return fail(MPPIB_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(e->desc.device));
  CUDA_TRY(cudaMemcpyAsync(host_costs, e->costs_d, (size_t)e->D * e->n_local * sizeof(float), cudaMemcpyDeviceToHost,
                           e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MPPIB_OK;
}

int mppib_get_noise(mppib_engine* e, float* host_eps)
{
  if (!e || !host_eps)
    return fail(MPPIB_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(e->desc.device));
  CUDA_TRY(cudaMemcpyAsync(host_eps, e->eps_d, (size_t)e->n_local * e->TC * sizeof(float), cudaMemcpyDeviceToHost,
                           e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MPPIB_OK;
}

int mppib_get_samples(mppib_engine* e, float* host_samples)
{
  if (!e || !host_samples)
    return fail(MPPIB_ERR_INVALID_ARG, "null argument");
  if (!e->writeback)
    return fail(MPPIB_ERR_STATE, "engine was created without MPPIB_FLAG_WRITEBACK_CONTROLS");
  CUDA_TRY(cudaSetDevice(e->desc.device));
  CUDA_TRY(cudaMemcpyAsync(host_samples, e->controls_d, (size_t)e->D * e->n_local * e->TC * sizeof(float),
                           cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MPPIB_OK;
}

int mppib_get_weights(mppib_engine* e, float* host_weights)
{
  if (!e || !host_weights)
    return fail(MPPIB_ERR_INVALID_ARG, "null argument");
  if (!e->solved_once)
    return fail(MPPIB_ERR_STATE, "no solve has been run yet");
  CUDA_TRY(cudaSetDevice(e->desc.device));
  const size_t n = (size_t)e->D * e->n_local;
  if (!e->weights_d)
    CUDA_TRY(cudaMalloc(&e->weights_d, n * sizeof(float)));
  const dim3 grid((e->n_local + 255) / 256 > 1024 ? 1024 : (e->n_local + 255) / 256, e->D);
  weights_kernel<<<grid, 256, 0, e->stream>>>(e->costs_d, e->result_d, e->n_local, e->pstride,
                                              (float)(1.0 / e->lambda), e->weights_d);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(host_weights, e->weights_d, n * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MPPIB_OK;
}

int mppib_enable_timing(mppib_engine* e, int enable)
{
  if (!e)
    return fail(MPPIB_ERR_INVALID_ARG, "null engine");
  e->timing = enable != 0;
  e->timing_valid = false;
  for (int i = 0; i < 4; i++)
    e->acc_ms[i] = 0.0;
  e->acc_n = 0;
  return MPPIB_OK;
}

int mppib_get_timing(mppib_engine* e, mppib_timing* out)
{
  if (!e || !out)
    return fail(MPPIB_ERR_INVALID_ARG, "null argument");
  if (e->acc_n == 0)
    return fail(MPPIB_ERR_STATE, "timing not enabled or no synchronous solve since it was enabled");
  // averages over the synchronous solves since mppib_enable_timing(e, 1)
  out->noise_ms = (float)(e->acc_ms[0] / e->acc_n);
  out->rollout_ms = (float)(e->acc_ms[1] / e->acc_n);
  out->reduce_ms = (float)(e->acc_ms[2] / e->acc_n);
  out->total_ms = (float)(e->acc_ms[3] / e->acc_n);
  out->samples = (int)e->acc_n;
  return MPPIB_OK;
}

int mppib_get_launch_info(mppib_engine* e, int* grid, int* block, int* smem_bytes, int* uses_tma,
                          int* kernels_per_solve)
{
  if (!e)
    return fail(MPPIB_ERR_INVALID_ARG, "null engine");
  if (grid)
    *grid = e->grid;
  if (block)
    *block = e->nn_tc ? e->bx : (e->ar_ws ? ar_ws::warpsPerGroup(e->ws_pspw) * e->bx : e->bx / e->spt * e->lps);
  if (smem_bytes)
    *smem_bytes = (int)e->smem_bytes;
  if (uses_tma)
    *uses_tma = e->use_tma ? 1 : 0;
  if (kernels_per_solve)
    *kernels_per_solve = ((e->desc.world_size > 1) ? 3 : 2) + (e->xw_enabled ? 1 : 0);  // [K0] + K1 + K2 (+ K2'); cuRAND's
                                                                                         // own launches are not counted
  return MPPIB_OK;
}

int mppib_get_rng_info(mppib_engine* e, int* own_kernel, int* chunks, int* rounds_per_chunk)
{
  if (!e)
    return fail(MPPIB_ERR_INVALID_ARG, "null engine");
  if (own_kernel)
    *own_kernel = e->xw_enabled ? 1 : 0;
  if (chunks)
    *chunks = e->xw_chunks;
  if (rounds_per_chunk)
    *rounds_per_chunk = e->xw_rounds_per_chunk;
  return MPPIB_OK;
}

int mppib_local_rollouts(mppib_engine* e, int* n_local, int* n_offset)
{
  if (!e)
    return fail(MPPIB_ERR_INVALID_ARG, "null engine");
  if (n_local)
    *n_local = e->n_local;
  if (n_offset)
    *n_offset = e->n_offset;
  return MPPIB_OK;
}

}  // extern "C"
