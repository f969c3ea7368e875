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
__device__ __forceinline__ uint4 bf16_lo8(const float* v, const uint4 q) {      // bf16(v - hi) of 8 packed values
  uint4 l;
  l.x = tc::pack_bf16x2(v[0] - __uint_as_float(q.x << 16), v[1] - __uint_as_float(q.x & 0xffff0000u));
  l.y = tc::pack_bf16x2(v[2] - __uint_as_float(q.y << 16), v[3] - __uint_as_float(q.y & 0xffff0000u));
  l.z = tc::pack_bf16x2(v[4] - __uint_as_float(q.z << 16), v[5] - __uint_as_float(q.z & 0xffff0000u));
  l.w = tc::pack_bf16x2(v[6] - __uint_as_float(q.w << 16), v[7] - __uint_as_float(q.w & 0xffff0000u));
  return l;
}

// split_off != 0 (both kernels): also write the lo half of the bf16x3 split, bf16(value - hi), split_off uint4 further.
__global__ void __launch_bounds__(128) patch_gate_blocked_kernel(const float* __restrict__ vol, const float* __restrict__ w,
                                                                 const float* __restrict__ gate, uint4* __restrict__ out, int G, int D,
                                                                 int H, int W, size_t split_off) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const int y = blockIdx.y % H, d = blockIdx.y / H;
  const int G8 = G >> 3, chunk = blockIdx.z % G8, b = blockIdx.z / G8;
  const size_t HW = (size_t)H * W;
  // branch-free 3x3: out-of-range taps read a clamped address and get weight 0, so all 72 loads of the 8 groups can be in flight
  int yo[3], xo[3];
  float ym[3], xm[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int yy = y + k - 1, xx = x + k - 1;
    ym[k] = (yy >= 0 && yy < H) ? 1.0f : 0.0f;
    xm[k] = (xx >= 0 && xx < W) ? 1.0f : 0.0f;
    yo[k] = min(max(yy, 0), H - 1) * W;
    xo[k] = min(max(xx, 0), W - 1);
  }
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = chunk * 8 + i;
    const float* plane = vol + (((size_t)b * G + g) * D + d) * HW;
    float v[9];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) v[ky * 3 + kx] = __ldg(plane + yo[ky] + xo[kx]);
    float acc = 0.0f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) acc = fmaf(__ldg(w + g * 9 + ky * 3 + kx) * (ym[ky] * xm[kx]), v[ky * 3 + kx], acc);
    f[i] = sigmoidf_(__ldg(gate + ((size_t)b * G + g) * HW + (size_t)y * W + x)) * acc;
  }
  uint4 q;
  q.x = tc::pack_bf16x2(f[0], f[1]); q.y = tc::pack_bf16x2(f[2], f[3]);
  q.z = tc::pack_bf16x2(f[4], f[5]); q.w = tc::pack_bf16x2(f[6], f[7]);
  const int phase = ((d & 1) << 2) | ((y & 1) << 1) | (x & 1);
  const size_t S8 = (size_t)(D >> 1) * (H >> 1) * (W >> 1);
  const size_t o = (((size_t)b * 8 + phase) * G8 + chunk) * S8 + ((size_t)(d >> 1) * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1);
  out[o] = q;
  if (split_off) out[o + split_off] = bf16_lo8(f, q);
}

// Same op, 4 consecutive x per thread (W % 4 == 0): each of the 3 rows is one aligned float4 plus its two neighbours, so a
// thread issues 9 loads per group for 4 outputs instead of 36; the two x-parities of its outputs are two 32-byte runs.
__global__ void __launch_bounds__(128) patch_gate_blocked_x4_kernel(const float* __restrict__ vol, const float* __restrict__ w,
                                                                    const float* __restrict__ gate, uint4* __restrict__ out, int G, int D,
                                                                    int H, int W, size_t split_off) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (x >= W) return;
  const int y = blockIdx.y % H, d = blockIdx.y / H;
  const int G8 = G >> 3, chunk = blockIdx.z % G8, b = blockIdx.z / G8;
  const size_t HW = (size_t)H * W;
  int yo[3];
  float ym[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int yy = y + k - 1;
    ym[k] = (yy >= 0 && yy < H) ? 1.0f : 0.0f;
    yo[k] = min(max(yy, 0), H - 1) * W;
  }
  const float lm = x > 0 ? 1.0f : 0.0f, rm = x + 4 < W ? 1.0f : 0.0f;
  const int xl = max(x - 1, 0), xr = min(x + 4, W - 1);
  float f[4][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = chunk * 8 + i;
    const float* plane = vol + (((size_t)b * G + g) * D + d) * HW;
    float wk[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wk[t] = __ldg(w + g * 9 + t);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const float4 c = __ldg(reinterpret_cast<const float4*>(plane + yo[ky] + x));
      const float v[6] = {__ldg(plane + yo[ky] + xl) * lm, c.x, c.y, c.z, c.w, __ldg(plane + yo[ky] + xr) * rm};
      const float w0 = wk[ky * 3] * ym[ky], w1 = wk[ky * 3 + 1] * ym[ky], w2 = wk[ky * 3 + 2] * ym[ky];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(w2, v[j + 2], fmaf(w1, v[j + 1], fmaf(w0, v[j], acc[j])));
    }
    const float4 gl = __ldg(reinterpret_cast<const float4*>(gate + ((size_t)b * G + g) * HW + (size_t)y * W + x));
    f[0][i] = sigmoidf_(gl.x) * acc[0]; f[1][i] = sigmoidf_(gl.y) * acc[1];
    f[2][i] = sigmoidf_(gl.z) * acc[2]; f[3][i] = sigmoidf_(gl.w) * acc[3];
  }
  const size_t S8 = (size_t)(D >> 1) * (H >> 1) * (W >> 1);
  const size_t sp = ((size_t)(d >> 1) * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1);
  const int ph0 = ((d & 1) << 2) | ((y & 1) << 1);
#pragma unroll
  for (int pw = 0; pw < 2; ++pw) {
    uint4* o = out + (((size_t)b * 8 + ph0 + pw) * G8 + chunk) * S8 + sp;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float* v = f[2 * j + pw];
      uint4 q;
      q.x = tc::pack_bf16x2(v[0], v[1]); q.y = tc::pack_bf16x2(v[2], v[3]);
      q.z = tc::pack_bf16x2(v[4], v[5]); q.w = tc::pack_bf16x2(v[6], v[7]);
      o[j] = q;
      if (split_off) o[j + split_off] = bf16_lo8(v, q);
    }
  }
}

// out blocked bf16 (B, 2C/8, K, H, W, 8): chunks [0, C/8) = cf_l * a_k ; [C/8, 2C/8) = bilinear(cf_r, x - d_k) * a_k.
// One CTA = one image row x 64 pixels, thread (pixel, half of the samples).  The left tile and the two right rows the bilinear
// taps can touch (floor(iy), floor(iy)+1: the same for the whole row) are staged once in shared memory with a +-PAD column
// window (128-bit loads); every (sample, channel) corner read then hits shared memory.  Samples whose corners leave the
// window (|disparity| > PAD-2) read global memory instead.
constexpr int SC_TX = 64, SC_PAD = 24, SC_WW = SC_TX + 2 * SC_PAD, SC_C = 32;

__global__ void __launch_bounds__(128) sparse_concat_blocked_kernel(const float* __restrict__ cf_l, const float* __restrict__ cf_r,
                                                                    const float* __restrict__ disp, const float* __restrict__ att,
                                                                    uint4* __restrict__ out, int C, int K, int H, int W) {
  __shared__ __align__(16) float Ls[SC_C][SC_TX];
  __shared__ __align__(16) float Rs[SC_C][2][SC_WW];
  const int tx = threadIdx.x & (SC_TX - 1), half = threadIdx.x >> 6, tid = threadIdx.x;
  const int x0 = blockIdx.x * SC_TX, x = x0 + tx, y = blockIdx.y, b = blockIdx.z;
  const size_t HW = (size_t)H * W, pix = (size_t)y * W + x;
  const float iy = warp_coord((float)y, (float)(H - 1));
  const int y0 = (int)floorf(iy);
  const float* lb = cf_l + (size_t)b * C * HW + (size_t)y * W;
  const float* rb = cf_r + (size_t)b * C * HW;
  const bool vec4 = (W & 3) == 0 && ((reinterpret_cast<uintptr_t>(cf_l) | reinterpret_cast<uintptr_t>(cf_r)) & 15) == 0;
  const int C8 = C >> 3;
  const size_t cs = (size_t)K * HW;                                   // chunk stride (uint4)
  for (int c0 = 0; c0 < C; c0 += SC_C) {
    if (vec4) {
      for (int i = tid; i < SC_C * (SC_TX / 4); i += 128) {
        const int c = i / (SC_TX / 4), j = (i - c * (SC_TX / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + c < C && x0 + j < W) v = __ldg(reinterpret_cast<const float4*>(lb + (size_t)(c0 + c) * HW + x0 + j));
        *reinterpret_cast<float4*>(&Ls[c][j]) = v;
      }
      for (int i = tid; i < SC_C * 2 * (SC_WW / 4); i += 128) {
        const int c = i / (2 * (SC_WW / 4)), r = (i / (SC_WW / 4)) & 1, j = (i % (SC_WW / 4)) * 4;
        const int xx = x0 - SC_PAD + j, yy = y0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + c < C && xx >= 0 && xx < W && yy >= 0 && yy < H)
          v = __ldg(reinterpret_cast<const float4*>(rb + (size_t)(c0 + c) * HW + (size_t)yy * W + xx));
        *reinterpret_cast<float4*>(&Rs[c][r][j]) = v;
      }
    } else {
      for (int i = tid; i < SC_C * SC_TX; i += 128) {
        const int c = i / SC_TX, j = i - c * SC_TX;
        Ls[c][j] = (c0 + c < C && x0 + j < W) ? __ldg(lb + (size_t)(c0 + c) * HW + x0 + j) : 0.0f;
      }
      for (int i = tid; i < SC_C * 2 * SC_WW; i += 128) {
        const int c = i / (2 * SC_WW), r = (i / SC_WW) & 1, j = i % SC_WW;
        const int xx = x0 - SC_PAD + j, yy = y0 + r;
        Rs[c][r][j] = (c0 + c < C && xx >= 0 && xx < W && yy >= 0 && yy < H) ? __ldg(rb + (size_t)(c0 + c) * HW + (size_t)yy * W + xx) : 0.0f;
      }
    }
    __syncthreads();
    if (x < W) {
      const int nch = min(SC_C, C - c0) >> 3;                         // 8-channel chunks staged
      for (int k = half; k < K; k += 2) {
        const float d = __ldg(disp + ((size_t)b * K + k) * HW + pix);
        const float a = att ? __ldg(att + ((size_t)b * K + k) * HW + pix) : 1.0f;
        const float ix = warp_coord((float)x - d, (float)(W - 1));
        const Bilin q = make_bilin(ix, iy, H, W);
        const float jf = floorf(ix) - (float)(x0 - SC_PAD);
        const bool in_win = jf >= 0.0f && jf <= (float)(SC_WW - 2);
        const int j0 = in_win ? (int)jf : 0;
        uint4* ob = out + ((size_t)b * 2 * C8 + (c0 >> 3)) * cs + (size_t)k * HW + pix;
        for (int c8 = 0; c8 < nch; ++c8) {
          float l[8], r[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = c8 * 8 + i;
            l[i] = a * Ls[c][tx];
            const float rv = in_win ? (((Rs[c][0][j0] * q.w00 + Rs[c][0][j0 + 1] * q.w01) + Rs[c][1][j0] * q.w10) + Rs[c][1][j0 + 1] * q.w11)
                                    : bilin_fetch(rb + (size_t)(c0 + c) * HW, q);
            r[i] = a * rv;
          }
          uint4 ql, qr;
          ql.x = tc::pack_bf16x2(l[0], l[1]); ql.y = tc::pack_bf16x2(l[2], l[3]); ql.z = tc::pack_bf16x2(l[4], l[5]); ql.w = tc::pack_bf16x2(l[6], l[7]);
          qr.x = tc::pack_bf16x2(r[0], r[1]); qr.y = tc::pack_bf16x2(r[2], r[3]); qr.z = tc::pack_bf16x2(r[4], r[5]); qr.w = tc::pack_bf16x2(r[6], r[7]);
          __stcs(ob + (size_t)c8 * cs, ql);
          __stcs(ob + (size_t)(C8 + c8) * cs, qr);
        }
      }
    }
    __syncthreads();
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
  return ss_patch_gate_blocked_ex(volume, patch_w, gate_logits, out_s2d, B, G, D, H, W, 0, stream);
}

extern "C" int ss_patch_gate_blocked_ex(const float* volume, const float* patch_w, const float* gate_logits, void* out_s2d, int B, int G,
                                        int D, int H, int W, int split, void* stream) {
  const size_t split_off = split ? (size_t)B * G * (size_t)D * H * W / 8 : (size_t)0;     // one [B][8][G/8][D/2][H/2][W/2] stack, in uint4
  SS_REQUIRE(volume && patch_w && gate_logits && out_s2d, "ss_patch_gate_blocked: null pointer");
  SS_REQUIRE(B > 0 && G > 0 && D > 0 && H > 0 && W > 0, "ss_patch_gate_blocked: non-positive dimension");
  SS_REQUIRE(G % 8 == 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "ss_patch_gate_blocked: G %% 8 == 0 and even D,H,W required");
  SS_UNSUPPORTED((int64_t)D * H > 65535 || (int64_t)B * (G / 8) > 65535, "ss_patch_gate_blocked: grid dimension exceeds 65535");
  const bool x4 = W % 4 == 0 && ((reinterpret_cast<uintptr_t>(volume) | reinterpret_cast<uintptr_t>(gate_logits)) & 15) == 0;
  if (x4)
    patch_gate_blocked_x4_kernel<<<dim3(ceil_div(W / 4, 32), D * H, B * (G / 8)), 32, 0, (cudaStream_t)stream>>>(
        volume, patch_w, gate_logits, reinterpret_cast<uint4*>(out_s2d), G, D, H, W, split_off);
  else
    patch_gate_blocked_kernel<<<dim3(ceil_div(W, 128), D * H, B * (G / 8)), 128, 0, (cudaStream_t)stream>>>(
        volume, patch_w, gate_logits, reinterpret_cast<uint4*>(out_s2d), G, D, H, W, split_off);
  SS_CHECK_LAUNCH("ss_patch_gate_blocked");
  return SS_OK;
}

extern "C" int ss_sparse_concat_volume_blocked(const float* cf_l, const float* cf_r, const float* disp_topk, const float* att_topk_or_null,
                                               void* volume_blocked, int B, int C, int K, int H, int W, void* stream) {
  SS_REQUIRE(cf_l && cf_r && disp_topk && volume_blocked, "ss_sparse_concat_volume_blocked: null pointer");
  SS_REQUIRE(B > 0 && C > 0 && K > 0 && H > 1 && W > 1, "ss_sparse_concat_volume_blocked: bad dimension");
  SS_REQUIRE(C % 8 == 0, "ss_sparse_concat_volume_blocked: C=%d must be a multiple of 8", C);
  SS_UNSUPPORTED(H > 65535 || B > 65535, "ss_sparse_concat_volume_blocked: grid dimension exceeds 65535");
  sparse_concat_blocked_kernel<<<dim3(ceil_div(W, SC_TX), H, B), 128, 0, (cudaStream_t)stream>>>(
      cf_l, cf_r, disp_topk, att_topk_or_null, reinterpret_cast<uint4*>(volume_blocked), C, K, H, W);
  SS_CHECK_LAUNCH("ss_sparse_concat_volume_blocked");
  return SS_OK;
}

// segmenthead.conv2 (models/submodule.py:36,44): Conv2d 1x1 with bias from the blocked bf16 activation (B,C/8,H,W,8) to a few
// fp32 NCHW channels (Cout <= 8).  One thread per pixel; weights in shared memory; 16-byte channel-chunk loads.
namespace {
template <int CO>
__global__ void __launch_bounds__(256) pointwise_blocked_small_kernel(const uint4* __restrict__ in, const float* __restrict__ w,
                                                                      const float* __restrict__ bias, float* __restrict__ out, int C,
                                                                      int cout, size_t P) {
  extern __shared__ float sw[];                 // [CO][C]
  for (int i = threadIdx.x; i < CO * C; i += blockDim.x) sw[i] = (i / C) < cout ? __ldg(w + i) : 0.0f;
  __syncthreads();
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (pix >= P) return;
  float acc[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) acc[o] = (bias && o < cout) ? __ldg(bias + o) : 0.0f;
  for (int c8 = 0; c8 < C / 8; ++c8) {
    const uint4 q = __ldg(in + ((size_t)b * (C / 8) + c8) * P + pix);
    const uint32_t u[4] = {q.x, q.y, q.z, q.w};
    float x[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[2 * i] = __uint_as_float(u[i] << 16); x[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u); }
#pragma unroll
    for (int o = 0; o < CO; ++o)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[o] = fmaf(sw[o * C + c8 * 8 + i], x[i], acc[o]);
  }
  for (int o = 0; o < cout; ++o) out[((size_t)b * cout + o) * P + pix] = acc[o];
}
}  // namespace

extern "C" int ss_pointwise_blocked_small(const void* in_blocked, const float* weight, const float* bias_or_null, float* out, int B, int C,
                                          int Cout, int H, int W, void* stream) {
  SS_REQUIRE(in_blocked && weight && out, "ss_pointwise_blocked_small: null pointer");
  SS_REQUIRE(B > 0 && C > 0 && Cout > 0 && H > 0 && W > 0 && C % 8 == 0, "ss_pointwise_blocked_small: bad dimension");
  SS_UNSUPPORTED(Cout > 8 || C > 512 || B > 65535, "ss_pointwise_blocked_small: Cout=%d (max 8) / C=%d (max 512) unsupported", Cout, C);
  const size_t P = (size_t)H * W;
  pointwise_blocked_small_kernel<8><<<dim3((unsigned)ceil_div64(P, 256), B), 256, 8 * C * sizeof(float), (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(in_blocked), weight, bias_or_null, out, C, Cout, P);
  SS_CHECK_LAUNCH("ss_pointwise_blocked_small");
  return SS_OK;
}
