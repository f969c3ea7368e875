// Memory-bound kernels of the MobileViTv2-1.0 backbone (`Feature`, models/SemStereo.py:33-56, is
// timm's mobilevitv2_100; its architecture is restated in semstereo_b200/backbone.py).  The 1x1 convolutions (94 % of the FLOPs)
// run on the tensor cores through ss_conv2d_tc_ex; here is everything else, on the bf16 blocked layout [B][C/8][H][W][8]:
//   stem_conv3x3_s2   : Conv2d(3, 32, 3, s2, p1) + BN + SiLU from the fp32 NCHW image, output zero-padded to 64 channels
//   dwconv3x3_blocked : depthwise Conv2d 3x3 (stride 1 / 2) + BN + SiLU
//   groupnorm1_blocked: GroupNorm(1, C) -- two deterministic passes (partial sums per CTA, then normalise)
//   linear_attention  : the separable self-attention of MobileViTv2 (softmax over the patches of a pixel-parity class, context
//                       vector, relu(v) * context), computed WITHOUT unfolding: unfold(2x2) only renames pixel (y, x) to
//                       (patch position p = 2*(y&1) + (x&1), patch index), every other op of the block is position-wise.
#include "tc_common.cuh"

namespace {

__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

__device__ __forceinline__ void unpack8(const uint4 q, float* f) {
  f[0] = __uint_as_float(q.x << 16); f[1] = __uint_as_float(q.x & 0xffff0000u);
  f[2] = __uint_as_float(q.y << 16); f[3] = __uint_as_float(q.y & 0xffff0000u);
  f[4] = __uint_as_float(q.z << 16); f[5] = __uint_as_float(q.z & 0xffff0000u);
  f[6] = __uint_as_float(q.w << 16); f[7] = __uint_as_float(q.w & 0xffff0000u);
}

__device__ __forceinline__ uint4 pack8f(const float* v) {
  uint4 q;
  q.x = tc::pack_bf16x2(v[0], v[1]); q.y = tc::pack_bf16x2(v[2], v[3]);
  q.z = tc::pack_bf16x2(v[4], v[5]); q.w = tc::pack_bf16x2(v[6], v[7]);
  return q;
}

// ---- stem ---------------------------------------------------------------------------------------------------------------
constexpr int STEM_C = 32;
__global__ void __launch_bounds__(128) stem_conv_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, uint4* __restrict__ out, int H, int W, int chunks_out) {
  __shared__ float sw[STEM_C * 27], ss[STEM_C], st[STEM_C];
  for (int i = threadIdx.x; i < STEM_C * 27; i += blockDim.x) sw[i] = __ldg(w + i);
  for (int i = threadIdx.x; i < STEM_C; i += blockDim.x) { ss[i] = __ldg(scale + i); st[i] = __ldg(shift + i); }
  __syncthreads();
  const int Ho = H / 2, Wo = W / 2;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= Wo) return;
  float in[27];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = 2 * y + ky - 1, xx = 2 * x + kx - 1;
        in[c * 9 + ky * 3 + kx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + ((size_t)(b * 3 + c) * H + yy) * W + xx) : 0.0f;
      }
  const size_t P = (size_t)Ho * Wo, pix = (size_t)y * Wo + x;
#pragma unroll 1
  for (int c8 = 0; c8 < STEM_C / 8; ++c8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int co = c8 * 8 + i;
      float a = 0.0f;
#pragma unroll
      for (int t = 0; t < 27; ++t) a = fmaf(sw[co * 27 + t], in[t], a);
      v[i] = silu(fmaf(a, ss[co], st[co]));
    }
    out[((size_t)b * chunks_out + c8) * P + pix] = pack8f(v);
  }
  for (int c8 = STEM_C / 8; c8 < chunks_out; ++c8) out[((size_t)b * chunks_out + c8) * P + pix] = make_uint4(0u, 0u, 0u, 0u);
}

// ---- depthwise 3x3 ------------------------------------------------------------------------------------------------------
template <int STRIDE>
__global__ void __launch_bounds__(128) dwconv_kernel(const uint4* __restrict__ in, const float* __restrict__ w, const float* __restrict__ scale,
                                                     const float* __restrict__ shift, uint4* __restrict__ out, int C8, int H, int W, int act) {
  const int Ho = H / STRIDE, Wo = W / STRIDE;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  const int chunk = blockIdx.z % C8, b = blockIdx.z / C8;
  if (x >= Wo) return;
  const uint4* ip = in + ((size_t)b * C8 + chunk) * H * W;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = STRIDE * y + ky - 1;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = STRIDE * x + kx - 1;
      if (xx < 0 || xx >= W) continue;
      float f[8];
      unpack8(__ldg(ip + (size_t)yy * W + xx), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(__ldg(w + (chunk * 8 + i) * 9 + ky * 3 + kx), f[i], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = chunk * 8 + i;
    acc[i] = fmaf(acc[i], __ldg(scale + c), __ldg(shift + c));
    if (act == 2) acc[i] = silu(acc[i]);
    else if (act == 1) acc[i] = fmaxf(acc[i], 0.0f);
  }
  out[((size_t)b * C8 + chunk) * Ho * Wo + (size_t)y * Wo + x] = pack8f(acc);
}

// Same op, 4 consecutive output pixels per thread (Wo % 4 == 0): the (4*STRIDE + 2)-column input window of a row is loaded and
// unpacked once for the 4 outputs (18 loads per 4 outputs at stride 1 instead of 36), and the 72 weights of the channel chunk are
// staged in shared memory as [tap][8 channels] so that a tap costs two broadcast 128-bit loads instead of 8 scalar global loads per
// output.  The one-pixel kernel above was instruction-bound at ~1/4 of the HBM rate on the backbone's 512^2 layers (round 2).
template <int STRIDE>
__global__ void __launch_bounds__(128) dwconv4_kernel(const uint4* __restrict__ in, const float* __restrict__ w, const float* __restrict__ scale,
                                                      const float* __restrict__ shift, uint4* __restrict__ out, int C8, int H, int W, int act) {
  constexpr int NC = 4 * STRIDE + 2 - (STRIDE - 1);          // input columns a thread touches: 6 (stride 1), 9 (stride 2)
  __shared__ __align__(16) float sw[9][8];
  __shared__ float ssc[8], ssh[8];
  const int Ho = H / STRIDE, Wo = W / STRIDE;
  const int chunk = blockIdx.z % C8, b = blockIdx.z / C8;
  if (threadIdx.x < 72) sw[threadIdx.x % 9][threadIdx.x / 9] = __ldg(w + (chunk * 8 + threadIdx.x / 9) * 9 + threadIdx.x % 9);
  if (threadIdx.x >= 96 && threadIdx.x < 104) ssc[threadIdx.x - 96] = __ldg(scale + chunk * 8 + threadIdx.x - 96);
  if (threadIdx.x >= 104 && threadIdx.x < 112) ssh[threadIdx.x - 104] = __ldg(shift + chunk * 8 + threadIdx.x - 104);
  __syncthreads();
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
  if (x0 >= Wo) return;
  const uint4* ip = in + ((size_t)b * C8 + chunk) * H * W;
  float acc[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
  const int xin0 = STRIDE * x0 - 1;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = STRIDE * y + ky - 1;
    if (yy < 0 || yy >= H) continue;
    float f[NC][8];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int xx = xin0 + c;
      if (xx >= 0 && xx < W) unpack8(__ldg(ip + (size_t)yy * W + xx), f[c]);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[c][i] = 0.f;
      }
    }
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const float4 wa = *reinterpret_cast<const float4*>(&sw[ky * 3 + kx][0]), wb = *reinterpret_cast<const float4*>(&sw[ky * 3 + kx][4]);
      const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(wv[i], f[STRIDE * j + kx][i], acc[j][i]);
    }
  }
  uint4* op = out + ((size_t)b * C8 + chunk) * Ho * Wo + (size_t)y * Wo + x0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = fmaf(acc[j][i], ssc[i], ssh[i]);
      if (act == 2) v = silu(v);
      else if (act == 1) v = fmaxf(v, 0.0f);
      acc[j][i] = v;
    }
    op[j] = pack8f(acc[j]);
  }
}

// ---- GroupNorm(1, C) ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum of two values (256 threads); result valid in every thread
__device__ __forceinline__ void block_sum2(float& a, float& b) {
  __shared__ float ra[8], rb[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  a = warp_sum(a); b = warp_sum(b);
  __syncthreads();
  if (lane == 0) { ra[wid] = a; rb[wid] = b; }
  __syncthreads();
  a = 0.f; b = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { a += ra[i]; b += rb[i]; }
}

// partial[b][blk] = (sum, sum of squares) over this CTA's slice of sample b (n16 = C/8 * H * W 16-byte voxels, contiguous)
__global__ void __launch_bounds__(256) gn_stats_kernel(const uint4* __restrict__ x, float2* __restrict__ partial, size_t n16) {
  const int b = blockIdx.y;
  const uint4* xp = x + (size_t)b * n16;
  float s = 0.f, q = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    float f[8];
    unpack8(__ldg(xp + i), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { s += f[k]; q = fmaf(f[k], f[k], q); }
  }
  block_sum2(s, q);
  if (threadIdx.x == 0) partial[(size_t)b * gridDim.x + blockIdx.x] = make_float2(s, q);
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const uint4* __restrict__ x, const float2* __restrict__ partial, int nblk,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta, uint4* __restrict__ out,
                                                       size_t HW, size_t n16, float eps) {
  const int b = blockIdx.y;
  float s = 0.f, q = 0.f;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) { const float2 p = __ldg(partial + (size_t)b * nblk + i); s += p.x; q += p.y; }
  block_sum2(s, q);
  const double n = (double)n16 * 8.0;
  const double mean_d = (double)s / n;
  const float mean = (float)mean_d, rstd = rsqrtf(fmaxf((float)((double)q / n - mean_d * mean_d), 0.0f) + eps);
  const uint4* xp = x + (size_t)b * n16;
  uint4* op = out + (size_t)b * n16;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const int chunk = (int)(i / HW);
    float f[8];
    unpack8(__ldg(xp + i), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = fmaf((f[k] - mean) * rstd, __ldg(gamma + chunk * 8 + k), __ldg(beta + chunk * 8 + k));
    op[i] = pack8f(f);
  }
}

// ---- separable (linear) self-attention -------------------------------------------------------------------------------------
// qkv blocked: chunks [0, d/8) = key, [d/8, 2d/8) = value, chunk 2d/8 lane 0 = query (the caller orders the projection so).
// (1) per (b, parity p): max and 1/sum(exp) of the query over the pixels of that parity class
__global__ void __launch_bounds__(256) la_stats_kernel(const __nv_bfloat16* __restrict__ qkv, float2* __restrict__ stats, int CH, int d8, int H, int W) {
  const int p = blockIdx.x, b = blockIdx.y;
  const int py = p >> 1, px = p & 1, h2 = H / 2, w2 = W / 2, n = h2 * w2;
  const __nv_bfloat16* q = qkv + ((size_t)b * CH + 2 * d8) * H * W * 8;
  float m = -INFINITY, s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int y = 2 * (i / w2) + py, x = 2 * (i % w2) + px;
    const float v = __bfloat162float(q[((size_t)y * W + x) * 8]);
    const float mn = fmaxf(m, v);
    s = s * __expf(m - mn) + __expf(v - mn);
    m = mn;
  }
  __shared__ float rm[8], rs[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float wm = warp_max(m);
  s = warp_sum(m == -INFINITY ? 0.0f : s * __expf(m - wm));       // lanes (or whole warps) that saw no pixel contribute nothing
  if (lane == 0) { rm[wid] = wm; rs[wid] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = -INFINITY;
    for (int i = 0; i < 8; ++i) M = fmaxf(M, rm[i]);
    float S = 0.f;
    for (int i = 0; i < 8; ++i) S += rm[i] == -INFINITY ? 0.0f : rs[i] * __expf(rm[i] - M);
    stats[b * 4 + p] = make_float2(M, 1.0f / S);
  }
}
// (2) context[b][p][c] = sum over the pixels of parity p of key[c] * softmax(query); CTA = (key chunk, p, b)
__global__ void __launch_bounds__(256) la_context_kernel(const __nv_bfloat16* __restrict__ qkv, const float2* __restrict__ stats,
                                                         float* __restrict__ context, int CH, int d8, int H, int W) {
  const int chunk = blockIdx.x, p = blockIdx.y, b = blockIdx.z;
  const int py = p >> 1, px = p & 1, h2 = H / 2, w2 = W / 2, n = h2 * w2;
  const size_t HW = (size_t)H * W;
  const __nv_bfloat16* q = qkv + ((size_t)b * CH + 2 * d8) * HW * 8;
  const uint4* k = reinterpret_cast<const uint4*>(qkv) + ((size_t)b * CH + chunk) * HW;
  const float2 st = __ldg(stats + b * 4 + p);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const size_t pix = (size_t)(2 * (i / w2) + py) * W + 2 * (i % w2) + px;
    const float e = __expf(__bfloat162float(q[pix * 8]) - st.x) * st.y;
    float f[8];
    unpack8(__ldg(k + pix), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[j], e, acc[j]);
  }
  __shared__ float red[8][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float v = warp_sum(acc[j]);
    if (lane == 0) red[wid][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += red[i][threadIdx.x];
    context[((size_t)b * 4 + p) * d8 * 8 + chunk * 8 + threadIdx.x] = v;
  }
}
// (3) out = relu(value) * context[parity of the pixel]
__global__ void __launch_bounds__(256) la_apply_kernel(const uint4* __restrict__ qkv, const float* __restrict__ context, uint4* __restrict__ out,
                                                       int CH, int d8, int H, int W) {
  const size_t HW = (size_t)H * W;
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y, b = blockIdx.z;
  if (pix >= HW) return;
  const int y = (int)(pix / W), x = (int)(pix % W), p = ((y & 1) << 1) | (x & 1);
  float f[8];
  unpack8(__ldg(qkv + ((size_t)b * CH + d8 + chunk) * HW + pix), f);
  const float* c = context + ((size_t)b * 4 + p) * d8 * 8 + chunk * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.0f) * __ldg(c + j);
  out[((size_t)b * d8 + chunk) * HW + pix] = pack8f(f);
}

}  // namespace

extern "C" int ss_stem_conv3x3_s2(const float* image, const float* weight, const float* scale, const float* shift, void* out_blocked, int B,
                                  int H, int W, int Cout_padded, void* stream) {
  SS_REQUIRE(image && weight && scale && shift && out_blocked, "ss_stem_conv3x3_s2: null pointer");
  SS_REQUIRE(B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "ss_stem_conv3x3_s2: even positive dims required");
  SS_REQUIRE(Cout_padded >= STEM_C && Cout_padded % 8 == 0 && (reinterpret_cast<uintptr_t>(out_blocked) & 15) == 0,
             "ss_stem_conv3x3_s2: Cout_padded must be a multiple of 8 >= 32 and the output 16-byte aligned");
  SS_UNSUPPORTED(H / 2 > 65535 || B > 65535, "ss_stem_conv3x3_s2: grid dimension exceeds 65535");
  stem_conv_kernel<<<dim3(ceil_div(W / 2, 128), H / 2, B), 128, 0, (cudaStream_t)stream>>>(image, weight, scale, shift,
                                                                                          reinterpret_cast<uint4*>(out_blocked), H, W, Cout_padded / 8);
  SS_CHECK_LAUNCH("ss_stem_conv3x3_s2");
  return SS_OK;
}

extern "C" int ss_dwconv3x3_blocked(const void* in_blocked, const float* weight, const float* scale, const float* shift, void* out_blocked,
                                    int B, int C, int H, int W, int stride, int act, void* stream) {
  SS_REQUIRE(in_blocked && weight && scale && shift && out_blocked, "ss_dwconv3x3_blocked: null pointer");
  SS_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && H > 0 && W > 0, "ss_dwconv3x3_blocked: bad dimension");
  SS_REQUIRE(stride == 1 || (stride == 2 && H % 2 == 0 && W % 2 == 0), "ss_dwconv3x3_blocked: stride 1, or 2 with even dims");
  SS_REQUIRE(act >= 0 && act <= 2, "ss_dwconv3x3_blocked: act must be 0, 1 (ReLU) or 2 (SiLU)");
  SS_REQUIRE(((reinterpret_cast<uintptr_t>(in_blocked) | reinterpret_cast<uintptr_t>(out_blocked)) & 15) == 0, "ss_dwconv3x3_blocked: 16-byte alignment");
  SS_UNSUPPORTED(H / stride > 65535 || (long long)B * (C / 8) > 65535, "ss_dwconv3x3_blocked: grid dimension exceeds 65535");
  if (stride == 1 && W % 4 == 0) {          // (the 4-pixel variant measured slower at stride 2: 9 input columns per row in registers)
    const dim3 grid4(ceil_div(W / 4, 128), H, B * (C / 8));
    dwconv4_kernel<1><<<grid4, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(in_blocked), weight, scale, shift,
                                                              reinterpret_cast<uint4*>(out_blocked), C / 8, H, W, act);
    SS_CHECK_LAUNCH("ss_dwconv3x3_blocked");
    return SS_OK;
  }
  const dim3 grid(ceil_div(W / stride, 128), H / stride, B * (C / 8));
  if (stride == 1)
    dwconv_kernel<1><<<grid, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(in_blocked), weight, scale, shift,
                                                            reinterpret_cast<uint4*>(out_blocked), C / 8, H, W, act);
  else
    dwconv_kernel<2><<<grid, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(in_blocked), weight, scale, shift,
                                                            reinterpret_cast<uint4*>(out_blocked), C / 8, H, W, act);
  SS_CHECK_LAUNCH("ss_dwconv3x3_blocked");
  return SS_OK;
}

extern "C" int ss_groupnorm1_workspace_floats(int B) { return B * 256 * 2; }

extern "C" int ss_groupnorm1_blocked(const void* in_blocked, const float* gamma, const float* beta, void* out_blocked, float* workspace, int B,
                                     int C, int H, int W, float eps, void* stream) {
  SS_REQUIRE(in_blocked && gamma && beta && out_blocked && workspace, "ss_groupnorm1_blocked: null pointer");
  SS_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && H > 0 && W > 0, "ss_groupnorm1_blocked: bad dimension");
  SS_REQUIRE(((reinterpret_cast<uintptr_t>(in_blocked) | reinterpret_cast<uintptr_t>(out_blocked) | reinterpret_cast<uintptr_t>(workspace)) & 15) == 0,
             "ss_groupnorm1_blocked: 16-byte alignment");
  SS_UNSUPPORTED(B > 65535, "ss_groupnorm1_blocked: grid dimension exceeds 65535");
  const size_t HW = (size_t)H * W, n16 = (size_t)(C / 8) * HW;
  const int nblk = (int)(n16 / 1024 < 1 ? 1 : (n16 / 1024 > 256 ? 256 : n16 / 1024));
  gn_stats_kernel<<<dim3(nblk, B), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(in_blocked), reinterpret_cast<float2*>(workspace), n16);
  SS_CHECK_LAUNCH("ss_groupnorm1_blocked(stats)");
  const int nblk2 = (int)(n16 / 2048 < 1 ? 1 : (n16 / 2048 > 1024 ? 1024 : n16 / 2048));
  gn_apply_kernel<<<dim3(nblk2, B), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(in_blocked), reinterpret_cast<const float2*>(workspace),
                                                                    nblk, gamma, beta, reinterpret_cast<uint4*>(out_blocked), HW, n16, eps);
  SS_CHECK_LAUNCH("ss_groupnorm1_blocked(apply)");
  return SS_OK;
}

extern "C" int ss_linear_attention_workspace_floats(int B, int d) { return B * 4 * 2 + B * 4 * d; }

extern "C" int ss_linear_attention_blocked(const void* qkv_blocked, void* out_blocked, float* workspace, int B, int d, int H, int W, void* stream) {
  SS_REQUIRE(qkv_blocked && out_blocked && workspace, "ss_linear_attention_blocked: null pointer");
  SS_REQUIRE(B > 0 && d > 0 && d % 8 == 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "ss_linear_attention_blocked: d %% 8 == 0 and even H, W required");
  SS_REQUIRE(((reinterpret_cast<uintptr_t>(qkv_blocked) | reinterpret_cast<uintptr_t>(out_blocked) | reinterpret_cast<uintptr_t>(workspace)) & 15) == 0,
             "ss_linear_attention_blocked: 16-byte alignment");
  SS_UNSUPPORTED(B > 65535 || d / 8 > 65535, "ss_linear_attention_blocked: grid dimension exceeds 65535");
  const int d8 = d / 8, CH = 2 * d8 + 1;
  float2* stats = reinterpret_cast<float2*>(workspace);
  float* context = workspace + (size_t)B * 4 * 2;
  cudaStream_t st = (cudaStream_t)stream;
  la_stats_kernel<<<dim3(4, B), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv_blocked), stats, CH, d8, H, W);
  SS_CHECK_LAUNCH("ss_linear_attention_blocked(stats)");
  la_context_kernel<<<dim3(d8, 4, B), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv_blocked), stats, context, CH, d8, H, W);
  SS_CHECK_LAUNCH("ss_linear_attention_blocked(context)");
  la_apply_kernel<<<dim3((unsigned)ceil_div64((int64_t)H * W, 256), d8, B), 256, 0, st>>>(reinterpret_cast<const uint4*>(qkv_blocked), context,
                                                                                         reinterpret_cast<uint4*>(out_blocked), CH, d8, H, W);
  SS_CHECK_LAUNCH("ss_linear_attention_blocked(apply)");
  return SS_OK;
}
