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
    const float* plane = rp + (size_t)c * HW;
    float r = __ldg(plane + q.o00) * q.w00 + __ldg(plane + q.o01) * q.w01 + __ldg(plane + q.o10) * q.w10 +
              __ldg(plane + q.o11) * q.w11;
    __stcs(ol + (size_t)c * K * HW, a * __ldg(lp + (size_t)c * HW));
    __stcs(orr + (size_t)c * K * HW, a * r);
  }
}

// out[b,co,p] = act(scale[co] * sum_ci W[co,ci] * in[b,ci,p] + shift[co]);  tile 128 pixels x 32 couts, K chunk 16
__global__ void __launch_bounds__(256) pointwise_conv2d_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                               const float* __restrict__ scale, const float* __restrict__ shift,
                                                               float* __restrict__ out, int Cin, int Cout, int P, int relu) {
  __shared__ __align__(16) float As[16][128];
  __shared__ __align__(16) float Ws[16][32];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 128, co0 = blockIdx.y * 32;
  const int tp = threadIdx.x & 31, tc = threadIdx.x >> 5;   // 32 pixel-quads x 8 cout-quads
  const float* ib = in + (size_t)b * Cin * P;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  for (int k0 = 0; k0 < Cin; k0 += 16) {
    for (int i = threadIdx.x; i < 16 * 128; i += 256) {
      const int k = i >> 7, p = i & 127;
      As[k][p] = (k0 + k < Cin && p0 + p < P) ? __ldg(ib + (size_t)(k0 + k) * P + p0 + p) : 0.0f;
    }
    for (int i = threadIdx.x; i < 16 * 32; i += 256) {
      const int c = i >> 4, k = i & 15;
      Ws[k][c] = (k0 + k < Cin && co0 + c < Cout) ? __ldg(w + (size_t)(co0 + c) * Cin + k0 + k) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][4 * tp]);
      const float4 ww = *reinterpret_cast<const float4*>(&Ws[k][4 * tc]);
      const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int co = co0 + 4 * tc + j;
    if (co >= Cout) continue;
    const float s = scale ? __ldg(scale + co) : 1.0f, t = shift ? __ldg(shift + co) : 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = p0 + 4 * tp + i;
      if (p >= P) continue;
      float v = fmaf(acc[i][j], s, t);
      out[((size_t)b * Cout + co) * P + p] = relu ? fmaxf(v, 0.0f) : v;
    }
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
  pointwise_conv2d_kernel<<<dim3(ceil_div(P, 128), ceil_div(Cout, 32), B), 256, 0, (cudaStream_t)stream>>>(
      in, weight, scale_or_null, shift_or_null, out, Cin, Cout, P, relu);
  SS_CHECK_LAUNCH("ss_pointwise_conv2d");
  return SS_OK;
}
