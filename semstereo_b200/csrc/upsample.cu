// K11: full-resolution upsamplers, one fused pass each (all per-pixel channel math in registers).
//   ssr_upsample     : SSR_upsample.forward (models/submodule.py:421-431), BatchNorm in eval mode
//   context_upsample : models/submodule_.py:311-323
#include "common.cuh"

namespace {

template <int NC>
struct SsrPacked {
  float a0, b0;                                   // conv.0  BN(1) affine
  float wc[NC][9], bc[NC];                        // conv.1  Conv2d(1->NC, 3x3)
  float s2[NC], t2[NC];                           // conv.2  BN(NC) affine
  float w1[NC][NC], b1[NC], s1[NC], t1[NC];       // conv1.0 Conv2d 1x1 + conv1.1 BN
  float w2[NC][NC], b2[NC], sb2[NC], tb2[NC];     // conv2.0 Conv2d 1x1 + conv2.1 BN
  float w3[NC], b3;                               // conv3   Conv2d(NC->1, 1x1)
};

__device__ __forceinline__ void lin4(int dst, int n_in, int& i0, int& i1, float& l0, float& l1) {
  float src = fmaxf(((float)dst + 0.5f) * 0.25f - 0.5f, 0.0f);
  i0 = min((int)src, n_in - 1);
  i1 = min(i0 + 1, n_in - 1);
  l1 = src - (float)i0;
  l0 = 1.0f - l1;
}

template <int NC>
__global__ void __launch_bounds__(256) ssr_upsample_kernel(const float* __restrict__ depth_low, const float* __restrict__ spx,
                                                           const float* __restrict__ label, float* __restrict__ out,
                                                           const SsrPacked<NC> P, int h, int w) {
  const int H = 4 * h, W = 4 * w;
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y, b = blockIdx.z;
  if (X >= W) return;
  const float* dl = depth_low + (size_t)b * h * w;
  // bilinear x4 of the low-res disparity at the 3x3 full-res neighbourhood (zero outside: the conv pads BN output)
  float v[9], centre = 0.0f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = Y + dy;
    int y0 = 0, y1 = 0;
    float hy0 = 0.f, hy1 = 0.f;
    const bool yin = yy >= 0 && yy < H;
    if (yin) lin4(yy, h, y0, y1, hy0, hy1);
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = X + dx;
      float val = 0.0f;
      if (yin && xx >= 0 && xx < W) {
        int x0, x1;
        float wx0, wx1;
        lin4(xx, w, x0, x1, wx0, wx1);
        const float up = hy0 * (wx0 * __ldg(dl + y0 * w + x0) + wx1 * __ldg(dl + y0 * w + x1)) +
                         hy1 * (wx0 * __ldg(dl + y1 * w + x0) + wx1 * __ldg(dl + y1 * w + x1));
        if (dy == 0 && dx == 0) centre = up;
        val = fmaf(P.a0, up, P.b0);
      }
      v[(dy + 1) * 3 + dx + 1] = val;
    }
  }
  const size_t HW = (size_t)H * W, pix = (size_t)Y * W + X;
  float sp[NC], lab[NC];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    sp[i] = __ldcs(spx + ((size_t)b * NC + i) * HW + pix);
    lab[i] = __ldcs(label + ((size_t)b * NC + i) * HW + pix);
    m = fmaxf(m, lab[i]);
  }
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < NC; ++i) { lab[i] = expf(lab[i] - m); sum += lab[i]; }
  float in1[NC], g1[NC], g2[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) in1[i] = (lab[i] / sum) * sp[i];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    float a = P.b1[j];
#pragma unroll
    for (int i = 0; i < NC; ++i) a = fmaf(P.w1[j][i], in1[i], a);
    g1[j] = sigmoidf_(fmaf(P.s1[j], a, P.t1[j]));
  }
#pragma unroll
  for (int i = 0; i < NC; ++i) in1[i] = g1[i] * sp[i];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    float a = P.b2[j];
#pragma unroll
    for (int i = 0; i < NC; ++i) a = fmaf(P.w2[j][i], in1[i], a);
    g2[j] = sigmoidf_(fmaf(P.sb2[j], a, P.tb2[j]));
  }
  float res = P.b3;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    float a = P.bc[j];
#pragma unroll
    for (int t = 0; t < 9; ++t) a = fmaf(P.wc[j][t], v[t], a);
    res = fmaf(P.w3[j], fmaf(P.s2[j], a, P.t2[j]) * g2[j], res);
  }
  out[(size_t)b * HW + pix] = centre + res;
}

__global__ void __launch_bounds__(256) context_upsample_kernel(const float* __restrict__ depth_low, const float* __restrict__ upw,
                                                               float* __restrict__ out, int h, int w) {
  const int H = 4 * h, W = 4 * w;
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y, b = blockIdx.z;
  if (X >= W) return;
  const float* dl = depth_low + (size_t)b * h * w;
  const size_t HW = (size_t)H * W, pix = (size_t)Y * W + X;
  const int y = Y >> 2, x = X >> 2;
  float acc = 0.0f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = y + ky - 1, xx = x + kx - 1;
      const float dv = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(dl + yy * w + xx) : 0.0f;
      acc += dv * __ldcs(upw + ((size_t)b * 9 + ky * 3 + kx) * HW + pix);
    }
  out[(size_t)b * HW + pix] = acc;
}

}  // namespace

// `packed` is a HOST array of ss_ssr_param_count(num_classes) floats in the order of SsrPacked (see include/semstereo_b200.h).
extern "C" int ss_ssr_param_count(int num_classes) { return 2 + num_classes * (9 + 1 + 2 + 2 * (num_classes + 3) + 1) + 1; }

extern "C" int ss_ssr_upsample(const float* depth_low, const float* spx, const float* label, float* out, const float* packed_host,
                               int B, int h, int w, int num_classes, void* stream) {
  SS_REQUIRE(depth_low && spx && label && out && packed_host, "ss_ssr_upsample: null pointer");
  SS_REQUIRE(B > 0 && h > 0 && w > 0, "ss_ssr_upsample: non-positive dimension");
  SS_UNSUPPORTED(num_classes != 6, "ss_ssr_upsample: num_classes=%d unsupported (the model's spx head has 6 channels)", num_classes);
  SS_UNSUPPORTED(4 * h > 65535 || B > 65535, "ss_ssr_upsample: grid dimension exceeds 65535");
  SsrPacked<6> P;
  static_assert(sizeof(P) == 189 * sizeof(float), "packed SSR layout");
  memcpy(&P, packed_host, sizeof(P));
  ssr_upsample_kernel<6><<<dim3(ceil_div(4 * w, 256), 4 * h, B), 256, 0, (cudaStream_t)stream>>>(depth_low, spx, label, out, P, h, w);
  SS_CHECK_LAUNCH("ss_ssr_upsample");
  return SS_OK;
}

extern "C" int ss_context_upsample(const float* depth_low, const float* up_weights, float* out, int B, int h, int w, void* stream) {
  SS_REQUIRE(depth_low && up_weights && out, "ss_context_upsample: null pointer");
  SS_REQUIRE(B > 0 && h > 0 && w > 0, "ss_context_upsample: non-positive dimension");
  SS_UNSUPPORTED(4 * h > 65535 || B > 65535, "ss_context_upsample: grid dimension exceeds 65535");
  context_upsample_kernel<<<dim3(ceil_div(4 * w, 256), 4 * h, B), 256, 0, (cudaStream_t)stream>>>(depth_low, up_weights, out, h, w);
  SS_CHECK_LAUNCH("ss_context_upsample");
  return SS_OK;
}
