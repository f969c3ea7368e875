// Producers of the bf16 "blocked channels" layouts for the tensor-core stack: the memory-bound ops in front of the 3-D
// convolutions write the layout the TMA boxes want directly, so no fp32 volume is ever materialised in bf16 mode.
//   gate_sigmoid_blocked        : sigmoid(channelAtt logits) (SemStereo.py:100-102) as fp32 (B,C/8,H,W,8) for the conv epilogue
//   patch_gate_blocked          : `patch` depthwise (1,3,3) conv * gate (SemStereo.py:274,276) -> phase-split bf16
//   sparse_concat_volume_blocked: concat_volume_generator * att_topk (SemStereo.py:241-244,318) -> blocked bf16
#include "tc_common.cuh"

namespace {

__global__ void __launch_bounds__(256) gate_sigmoid_blocked_kernel(const float* __restrict__ logits, float4* __restrict__ out, int C,
                                                                   size_t P) {
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= P) return;
  const int chunk = blockIdx.y, b = blockIdx.z;
  const float* ip = logits + ((size_t)b * C + chunk * 8) * P + pix;
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = sigmoidf_(__ldg(ip + (size_t)i * P));
  float4* o = out + (((size_t)b * (C / 8) + chunk) * P + pix) * 2;
  o[0] = make_float4(f[0], f[1], f[2], f[3]);
  o[1] = make_float4(f[4], f[5], f[6], f[7]);
}

// vol (B,G,D,H,W) fp32, w (G,9), gate logits (B,G,H,W) -> phase-split bf16 (B,8,G/8,D/2,H/2,W/2,8); thread = 8 groups of one voxel
__global__ void __launch_bounds__(128) patch_gate_blocked_kernel(const float* __restrict__ vol, const float* __restrict__ w,
                                                                 const float* __restrict__ gate, uint4* __restrict__ out, int G, int D,
                                                                 int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const int y = blockIdx.y % H, d = blockIdx.y / H;
  const int G8 = G >> 3, chunk = blockIdx.z % G8, b = blockIdx.z / G8;
  const size_t HW = (size_t)H * W;
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = chunk * 8 + i;
    const float* plane = vol + (((size_t)b * G + g) * D + d) * HW;
    float acc = 0.0f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + kx - 1;
        if (xx < 0 || xx >= W) continue;
        acc = fmaf(__ldg(w + g * 9 + ky * 3 + kx), __ldg(plane + (size_t)yy * W + xx), acc);
      }
    }
    f[i] = sigmoidf_(__ldg(gate + ((size_t)b * G + g) * HW + (size_t)y * W + x)) * acc;
  }
  uint4 q;
  q.x = tc::pack_bf16x2(f[0], f[1]); q.y = tc::pack_bf16x2(f[2], f[3]);
  q.z = tc::pack_bf16x2(f[4], f[5]); q.w = tc::pack_bf16x2(f[6], f[7]);
  const int phase = ((d & 1) << 2) | ((y & 1) << 1) | (x & 1);
  const size_t S8 = (size_t)(D >> 1) * (H >> 1) * (W >> 1);
  out[(((size_t)b * 8 + phase) * G8 + chunk) * S8 + ((size_t)(d >> 1) * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)] = q;
}

// out blocked bf16 (B, 2C/8, K, H, W, 8): chunks [0, C/8) = cf_l * a_k ; [C/8, 2C/8) = bilinear(cf_r, x - d_k) * a_k.
// One thread = one pixel, all K samples: the left chunk is loaded once per 8 channels and reused by the K samples; the right
// reads of consecutive samples (ascending disparity) fall in the same few cache lines.
__global__ void __launch_bounds__(128) sparse_concat_blocked_kernel(const float* __restrict__ cf_l, const float* __restrict__ cf_r,
                                                                    const float* __restrict__ disp, const float* __restrict__ att,
                                                                    uint4* __restrict__ out, int C, int K, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const int y = blockIdx.y, b = blockIdx.z;
  const size_t HW = (size_t)H * W, pix = (size_t)y * W + x;
  const float iy = warp_coord((float)y, (float)(H - 1));
  const float* lp = cf_l + (size_t)b * C * HW + pix;
  const float* rp = cf_r + (size_t)b * C * HW;
  const float* dp = disp + (size_t)b * K * HW + pix;
  const float* ap = att ? att + (size_t)b * K * HW + pix : nullptr;
  const int C8 = C >> 3;
  const size_t cs = (size_t)K * HW;                                   // chunk stride (uint4)
  uint4* ob = out + ((size_t)b * 2 * C8) * cs + pix;
  for (int c8 = 0; c8 < C8; ++c8) {
    float l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) l[i] = __ldg(lp + (size_t)(c8 * 8 + i) * HW);
    const float* rc = rp + (size_t)c8 * 8 * HW;
    for (int k = 0; k < K; ++k) {
      const float d = __ldg(dp + (size_t)k * HW);
      const float a = ap ? __ldg(ap + (size_t)k * HW) : 1.0f;
      const Bilin q = make_bilin(warp_coord((float)x - d, (float)(W - 1)), iy, H, W);
      float r[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = a * bilin_fetch(rc + (size_t)i * HW, q);
      uint4 ql, qr;
      ql.x = tc::pack_bf16x2(a * l[0], a * l[1]); ql.y = tc::pack_bf16x2(a * l[2], a * l[3]);
      ql.z = tc::pack_bf16x2(a * l[4], a * l[5]); ql.w = tc::pack_bf16x2(a * l[6], a * l[7]);
      qr.x = tc::pack_bf16x2(r[0], r[1]); qr.y = tc::pack_bf16x2(r[2], r[3]); qr.z = tc::pack_bf16x2(r[4], r[5]); qr.w = tc::pack_bf16x2(r[6], r[7]);
      __stcs(ob + (size_t)c8 * cs + (size_t)k * HW, ql);
      __stcs(ob + (size_t)(C8 + c8) * cs + (size_t)k * HW, qr);
    }
  }
}

}  // namespace

extern "C" int ss_gate_sigmoid_blocked(const float* gate_logits, float* out_blocked, int B, int C, int H, int W, void* stream) {
  SS_REQUIRE(gate_logits && out_blocked && B > 0 && C > 0 && H > 0 && W > 0, "ss_gate_sigmoid_blocked: bad argument");
  SS_REQUIRE(C % 8 == 0, "ss_gate_sigmoid_blocked: C=%d must be a multiple of 8", C);
  SS_UNSUPPORTED(C / 8 > 65535 || B > 65535, "ss_gate_sigmoid_blocked: grid dimension exceeds 65535");
  const size_t P = (size_t)H * W;
  gate_sigmoid_blocked_kernel<<<dim3((unsigned)ceil_div64(P, 256), C / 8, B), 256, 0, (cudaStream_t)stream>>>(
      gate_logits, reinterpret_cast<float4*>(out_blocked), C, P);
  SS_CHECK_LAUNCH("ss_gate_sigmoid_blocked");
  return SS_OK;
}

extern "C" int ss_patch_gate_blocked(const float* volume, const float* patch_w, const float* gate_logits, void* out_s2d, int B, int G,
                                     int D, int H, int W, void* stream) {
  SS_REQUIRE(volume && patch_w && gate_logits && out_s2d, "ss_patch_gate_blocked: null pointer");
  SS_REQUIRE(B > 0 && G > 0 && D > 0 && H > 0 && W > 0, "ss_patch_gate_blocked: non-positive dimension");
  SS_REQUIRE(G % 8 == 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "ss_patch_gate_blocked: G %% 8 == 0 and even D,H,W required");
  SS_UNSUPPORTED((int64_t)D * H > 65535 || (int64_t)B * (G / 8) > 65535, "ss_patch_gate_blocked: grid dimension exceeds 65535");
  patch_gate_blocked_kernel<<<dim3(ceil_div(W, 128), D * H, B * (G / 8)), 128, 0, (cudaStream_t)stream>>>(
      volume, patch_w, gate_logits, reinterpret_cast<uint4*>(out_s2d), G, D, H, W);
  SS_CHECK_LAUNCH("ss_patch_gate_blocked");
  return SS_OK;
}

extern "C" int ss_sparse_concat_volume_blocked(const float* cf_l, const float* cf_r, const float* disp_topk, const float* att_topk_or_null,
                                               void* volume_blocked, int B, int C, int K, int H, int W, void* stream) {
  SS_REQUIRE(cf_l && cf_r && disp_topk && volume_blocked, "ss_sparse_concat_volume_blocked: null pointer");
  SS_REQUIRE(B > 0 && C > 0 && K > 0 && H > 1 && W > 1, "ss_sparse_concat_volume_blocked: bad dimension");
  SS_REQUIRE(C % 8 == 0, "ss_sparse_concat_volume_blocked: C=%d must be a multiple of 8", C);
  SS_UNSUPPORTED(H > 65535 || B > 65535, "ss_sparse_concat_volume_blocked: grid dimension exceeds 65535");
  sparse_concat_blocked_kernel<<<dim3(ceil_div(W, 128), H, B), 128, 0, (cudaStream_t)stream>>>(
      cf_l, cf_r, disp_topk, att_topk_or_null, reinterpret_cast<uint4*>(volume_blocked), C, K, H, W);
  SS_CHECK_LAUNCH("ss_sparse_concat_volume_blocked");
  return SS_OK;
}
