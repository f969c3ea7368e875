// Shared helpers for the sm_100a kernels of the SemStereo disparity hot path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "semstereo_b200.h"   // public C-ABI: declarations carry default visibility

#define SS_OK 0

void ss_set_error(const char* fmt, ...);

#define SS_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      ss_set_error(__VA_ARGS__);         \
      return SS_ERR_BAD_ARG;             \
    }                                    \
  } while (0)

#define SS_UNSUPPORTED(cond, ...)        \
  do {                                   \
    if (cond) {                          \
      ss_set_error(__VA_ARGS__);         \
      return SS_ERR_UNSUPPORTED;         \
    }                                    \
  } while (0)

// Checks the launch (not the execution: every entry point is enqueue-only).
#define SS_CHECK_LAUNCH(name)                                                        \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      ss_set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__));     \
      return SS_ERR_CUDA;                                                            \
    }                                                                                \
  } while (0)

#define SS_CUDA(call)                                                                \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      ss_set_error("%s failed: %s", #call, cudaGetErrorString(e__));                 \
      return SS_ERR_CUDA;                                                            \
    }                                                                                \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Number of SMs of the current device (cached per device id).
int ss_num_sms();
// Opt a kernel into > 48 KB dynamic shared memory (idempotent, cheap).
template <typename K>
static inline cudaError_t ss_allow_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// The five propagation taps (dy,dx) in output order: TL, C, BR, BL, TR
// (reference models/submodule.py:297-303, 366-371).
__device__ __constant__ const int kPropDy[5] = {-1, 0, 1, 1, -1};
__device__ __constant__ const int kPropDx[5] = {-1, 0, 1, -1, 1};

// Pixel-space x coordinate the reference's fp32 normalise/un-normalise round trip yields for
// (x - d) (models/submodule.py:276-279 + grid_sample align_corners=True).
__device__ __forceinline__ float warp_coord(float pos, float size_m1) {
  float half = __fdiv_rn(size_m1, 2.0f);                  // (W-1)/2
  float g = __fsub_rn(__fdiv_rn(pos, half), 1.0f);        // pos/((W-1)/2) - 1
  return __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.0f), 2.0f), size_m1);  // ((g+1)/2)*(W-1)
}

// bilinear sample with zeros padding at pixel coords (ix, iy) of one channel plane
struct Bilin {
  int o00, o01, o10, o11;     // plane offsets (clamped)
  float w00, w01, w10, w11;   // weights, zeroed for out-of-range corners
};

__device__ __forceinline__ Bilin make_bilin(float ix, float iy, int H, int W) {
  Bilin q;
  float fx0 = floorf(ix), fy0 = floorf(iy);
  float fx1 = fx0 + 1.0f, fy1 = fy0 + 1.0f;
  float wnw = (fx1 - ix) * (fy1 - iy), wne = (ix - fx0) * (fy1 - iy);
  float wsw = (fx1 - ix) * (iy - fy0), wse = (ix - fx0) * (iy - fy0);
  // compare in float first: |ix| may be huge for wild disparities
  bool vx0 = fx0 >= 0.0f && fx0 <= (float)(W - 1), vx1 = fx1 >= 0.0f && fx1 <= (float)(W - 1);
  bool vy0 = fy0 >= 0.0f && fy0 <= (float)(H - 1), vy1 = fy1 >= 0.0f && fy1 <= (float)(H - 1);
  int x0 = vx0 ? (int)fx0 : 0, x1 = vx1 ? (int)fx1 : 0, y0 = vy0 ? (int)fy0 : 0, y1 = vy1 ? (int)fy1 : 0;
  q.o00 = y0 * W + x0; q.o01 = y0 * W + x1; q.o10 = y1 * W + x0; q.o11 = y1 * W + x1;
  q.w00 = (vx0 && vy0) ? wnw : 0.0f; q.w01 = (vx1 && vy0) ? wne : 0.0f;
  q.w10 = (vx0 && vy1) ? wsw : 0.0f; q.w11 = (vx1 && vy1) ? wse : 0.0f;
  return q;
}

// A corner whose weight is exactly 0 contributes exactly 0 for finite data, so its load is skipped.  With the reference's
// fp32 coordinate round trip the y coordinate is exactly integral on ~75 % of the rows and integer disparities land on
// integral x on most pixels, so 1-2 of the 4 corner loads are the common case (the result is bit-identical either way).
__device__ __forceinline__ float bilin_fetch(const float* __restrict__ plane, const Bilin& q) {
  const float a = q.w00 != 0.0f ? __ldg(plane + q.o00) * q.w00 : 0.0f;
  const float b = q.w01 != 0.0f ? __ldg(plane + q.o01) * q.w01 : 0.0f;
  const float c = q.w10 != 0.0f ? __ldg(plane + q.o10) * q.w10 : 0.0f;
  const float d = q.w11 != 0.0f ? __ldg(plane + q.o11) * q.w11 : 0.0f;
  return ((a + b) + c) + d;
}

