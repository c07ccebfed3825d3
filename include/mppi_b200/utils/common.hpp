/*
 * mppi_b200/utils/common.hpp — small pieces every host header needs: the float2 the reference's public members use
 * (Dynamics::control_rngs_, include/mppi/dynamics/dynamics.cuh:511), and the error convention of
 * include/mppi/utils/gpu_err_chk.cuh:32-40 (print, then exit) applied to C-ABI status codes.
 */
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../../mppi_b200.h"
#include "../host_twins.h"
#include "../eigen_shim.hpp"

#if !defined(__VECTOR_TYPES_H__) && !defined(__CUDACC__)
struct float2
{
  float x, y;
};
struct dim3
{
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_)
  {
  }
};
typedef struct CUstream_st* cudaStream_t;
struct float3
{
  float x, y, z;
};
struct float4
{
  float x, y, z, w;
};
inline float3 make_float3(float x, float y, float z)
{
  return float3{ x, y, z };
}
struct cudaExtent
{
  size_t width, height, depth;
};
inline cudaExtent make_cudaExtent(size_t w, size_t h, size_t d)
{
  return cudaExtent{ w, h, d };
}
#endif

namespace mppi_b200
{
// HANDLE_ERROR equivalent (gpu_err_chk.cuh:32-40): report file/line and terminate. SMEM exhaustion keeps the
// reference's std::runtime_error (controllers/MPPI/mppi_controller.cu:64-76) so callers that catch it still can.
inline void handle_status(int status, const char* file, int line)
{
  if (status == MPPIB_OK)
    return;
  const std::string msg = std::string("MPPI-B200 error: ") + mppib_strerror(status) + ": " + mppib_last_error() + " at " +
                          file + ":" + std::to_string(line);
  if (status == MPPIB_ERR_SMEM || status == MPPIB_ERR_UNSUPPORTED)
    throw std::runtime_error(msg);
  fprintf(stderr, "%s\n", msg.c_str());
  exit(status);
}
}  // namespace mppi_b200
#define MPPIB_HANDLE(expr) (::mppi_b200::handle_status((expr), __FILE__, __LINE__))
