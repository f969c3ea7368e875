// K3, K9 and the 1x1 convolutions of channelAtt.
//   patch_gate           : `patch` depthwise Conv3d (1,3,3) (SemStereo.py:219,274) fused with the channelAtt gate
//                          sigmoid(logits)[:, :, None] * cv (SemStereo.py:98-103)
//   sparse_concat_volume : concat_volume_generator (SemStereo.py:241-244) * att_topk (:318)
//   pointwise_conv2d     : BasicConv 1x1 (+BN+ReLU) / Conv2d 1x1 (+bias) of channelAtt.im_att (SemStereo.py:93-95)
#include "common.cuh"

namespace {

// out[b,g,d,y,x] = sigmoid(gate[b,g,y,x]) * sum_{ky,kx} w[g,ky,kx] * vol[b,g,d,y+ky-1,x+kx-1]
__global__ void __launch_bounds__(256) patch_gate_kernel(const float* __restrict__ vol, const float* __restrict__ w,
                                                         const float* __restrict__ gate, float* __restrict__ out,
                                                         int G, int D, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  int r = blockIdx.y;                 // (d, y)
  const int y = r % H, d = r / H;
  const int g = blockIdx.z % G, b = blockIdx.z / G;
  const size_t HW = (size_t)H * W;
  const float* plane = vol + (((size_t)b * G + g) * D + d) * HW;
  float acc;
  if (w) {
    acc = 0.0f;
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
  } else {
    acc = __ldg(plane + (size_t)y * W + x);
  }
  if (gate) acc = sigmoidf_(__ldg(gate + ((size_t)b * G + g) * HW + (size_t)y * W + x)) * acc;
  out[(((size_t)b * G + g) * D + d) * HW + (size_t)y * W + x] = acc;
}

// out (B, 2C, K, H, W): [:C] = cf_l * a_k ; [C:] = bilinear(cf_r, x - d_k) * a_k
__global__ void __launch_bounds__(128) sparse_concat_kernel(const float* __restrict__ cf_l, const float* __restrict__ cf_r,
                                                            const float* __restrict__ disp, const float* __restrict__ att,
                                                            float* __restrict__ out, int C, int K, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const int y = blockIdx.y % H, k = blockIdx.y / H, b = blockIdx.z;
  const size_t HW = (size_t)H * W, pix = (size_t)y * W + x;
  const float d = __ldg(disp + ((size_t)b * K + k) * HW + pix);
  const float a = att ? __ldg(att + ((size_t)b * K + k) * HW + pix) : 1.0f;
  const Bilin q = make_bilin(warp_coord((float)x - d, (float)(W - 1)), warp_coord((float)y, (float)(H - 1)), H, W);
  const float* lp = cf_l + (size_t)b * C * HW + pix;
  const float* rp = cf_r + (size_t)b * C * HW;
  float* ol = out + ((size_t)b * 2 * C * K + k) * HW + pix;
  float* orr = ol + (size_t)C * K * HW;
  for (int c = 0; c < C; ++c) {
    const float r = bilin_fetch(rp + (size_t)c * HW, q);
    __stcs(ol + (size_t)c * K * HW, a * __ldg(lp + (size_t)c * HW));
    __stcs(orr + (size_t)c * K * HW, a * r);
  }
}

// out[b,co,p] = act(scale[co] * sum_ci W[co,ci] * in[b,ci,p] + shift[co]).  Tile = 128 pixels x BN couts, K chunk 16, register
// double buffering of the global loads; thread tile = 4 pixels x (BN/8) couts (warp-uniform couts -> weight reads broadcast).
template <int BN>
__global__ void __launch_bounds__(256) pointwise_conv2d_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                               const float* __restrict__ scale, const float* __restrict__ shift,
                                                               float* __restrict__ out, int Cin, int Cout, int P, int relu) {
  constexpr int TN = BN / 8;
  __shared__ __align__(16) float As[2][16][128];
  __shared__ __align__(16) float Ws[2][16][BN];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 128, co0 = blockIdx.y * BN;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* ib = in + (size_t)b * Cin * P;
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
  float a_reg[8], w_reg[(16 * BN) / 256];
  const int ap = threadIdx.x & 127, ak0 = threadIdx.x >> 7;           // A gather: pixel, first k row (step 2)
  auto load = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + ak0 + 2 * j;
      a_reg[j] = (k < Cin && p0 + ap < P) ? __ldg(ib + (size_t)k * P + p0 + ap) : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < (16 * BN) / 256; ++j) {
      const int i = threadIdx.x + j * 256, c = i >> 4, k = i & 15;
      w_reg[j] = (k0 + k < Cin && co0 + c < Cout) ? __ldg(w + (size_t)(co0 + c) * Cin + k0 + k) : 0.0f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) As[buf][ak0 + 2 * j][ap] = a_reg[j];
#pragma unroll
    for (int j = 0; j < (16 * BN) / 256; ++j) {
      const int i = threadIdx.x + j * 256, c = i >> 4, k = i & 15;
      Ws[buf][k][c] = w_reg[j];
    }
  };
  load(0);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < Cin; k0 += 16) {
    const bool more = k0 + 16 < Cin;
    if (more) load(k0 + 16);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][4 * lane]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      float wv[TN];
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 ww = *reinterpret_cast<const float4*>(&Ws[buf][k][warp * TN + j]);
        wv[j] = ww.x; wv[j + 1] = ww.y; wv[j + 2] = ww.z; wv[j + 3] = ww.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int co = co0 + warp * TN + j;
    if (co >= Cout) continue;
    const float s = scale ? __ldg(scale + co) : 1.0f, t = shift ? __ldg(shift + co) : 0.0f;
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = fmaf(acc[i][j], s, t); if (relu) v[i] = fmaxf(v[i], 0.0f); }
    float* o = out + ((size_t)b * Cout + co) * P + p0 + 4 * lane;
    if ((P & 3) == 0 && p0 + 4 * lane + 3 < P) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    else
      for (int i = 0; i < 4; ++i)
        if (p0 + 4 * lane + i < P) o[i] = v[i];
  }
}

}  // namespace

extern "C" int ss_patch_gate(const float* volume, const float* patch_w_or_null, const float* gate_logits_or_null, float* out,
                             int B, int G, int D, int H, int W, void* stream) {
  SS_REQUIRE(volume && out, "ss_patch_gate: null pointer");
  SS_REQUIRE(B > 0 && G > 0 && D > 0 && H > 0 && W > 0, "ss_patch_gate: non-positive dimension");
  SS_UNSUPPORTED((int64_t)D * H > 65535 || (int64_t)B * G > 65535, "ss_patch_gate: grid dimension exceeds 65535");
  patch_gate_kernel<<<dim3(ceil_div(W, 256), D * H, B * G), 256, 0, (cudaStream_t)stream>>>(volume, patch_w_or_null,
                                                                                         gate_logits_or_null, out, G, D, H, W);
  SS_CHECK_LAUNCH("ss_patch_gate");
  return SS_OK;
}

extern "C" int ss_sparse_concat_volume(const float* cf_l, const float* cf_r, const float* disp_topk, const float* att_topk_or_null,
                                       float* volume, int B, int C, int K, int H, int W, void* stream) {
  SS_REQUIRE(cf_l && cf_r && disp_topk && volume, "ss_sparse_concat_volume: null pointer");
  SS_REQUIRE(B > 0 && C > 0 && K > 0 && H > 1 && W > 1, "ss_sparse_concat_volume: bad dimension");
  SS_UNSUPPORTED((int64_t)K * H > 65535 || B > 65535, "ss_sparse_concat_volume: grid dimension exceeds 65535");
  sparse_concat_kernel<<<dim3(ceil_div(W, 128), K * H, B), 128, 0, (cudaStream_t)stream>>>(cf_l, cf_r, disp_topk, att_topk_or_null,
                                                                                        volume, C, K, H, W);
  SS_CHECK_LAUNCH("ss_sparse_concat_volume");
  return SS_OK;
}

extern "C" int ss_pointwise_conv2d(const float* in, const float* weight, const float* scale_or_null, const float* shift_or_null,
                                   float* out, int B, int Cin, int Cout, int P, int relu, void* stream) {
  SS_REQUIRE(in && weight && out, "ss_pointwise_conv2d: null pointer");
  SS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && P > 0, "ss_pointwise_conv2d: non-positive dimension");
  SS_UNSUPPORTED(B > 65535 || ceil_div(Cout, 32) > 65535, "ss_pointwise_conv2d: grid dimension exceeds 65535");
  if (Cout >= 64)
    pointwise_conv2d_kernel<64><<<dim3(ceil_div(P, 128), ceil_div(Cout, 64), B), 256, 0, (cudaStream_t)stream>>>(
        in, weight, scale_or_null, shift_or_null, out, Cin, Cout, P, relu);
  else
    pointwise_conv2d_kernel<32><<<dim3(ceil_div(P, 128), ceil_div(Cout, 32), B), 256, 0, (cudaStream_t)stream>>>(
        in, weight, scale_or_null, shift_or_null, out, Cin, Cout, P, relu);
  SS_CHECK_LAUNCH("ss_pointwise_conv2d");
  return SS_OK;
}
