// Training closure of the 3-D convolution stack (BASELINE config #5): what convbn_3d /
// BasicConv(is_3d) (models/submodule_other.py:845-848, models/submodule.py:89-116) need beyond the inference kernels when the
// model trains (main_us3d.py:186-222): the weight gradient of Conv3d, and BatchNorm3d with BATCH statistics, forward and backward.
//   * input gradient of Conv3d: no new kernel -- dX of a k3 s1 conv is the k3 s1 conv of dY with the flipped / transposed weight,
//     dX of a k3 s2 p1 conv is ConvTranspose3d(k3, s2, p1, op1) of dY with the same weight, dX of a k1 conv is the k1 conv with
//     W^T: all three are launches of the forward kernels with a re-packed weight (semstereo_b200/train_ops.py: the bf16x3
//     tensor-core kernels where the layer has such a configuration, else ss_conv3d_f32 of csrc/conv3d_f32.cu);
//   * conv3d_wgrad: dW[tap][ci][co] = sum over (b, output voxel o) of dY[b,co,o] * X[b,ci, o*stride - pad + tap]
//     -- a GEMM with K = B * output voxels, split over CTAs along K, fp32 FFMA, atomic accumulation into a zeroed dW;
//   * bn_stats / bn_apply / bn_backward: per-channel batch mean and biased variance over (B, D, H, W), normalise + affine
//     (+ ReLU), and the standard BatchNorm VJP  dx = w * rstd / N * (N dy - sum(dy) - xhat * sum(dy xhat)).
// fp32 throughout (the reference trains in fp32); deterministic except for the order of the wgrad atomics.
#include "common.cuh"

namespace {

// ---- Conv3d weight gradient -----------------------------------------------------------------------------------------------
constexpr int WG_T = 32, WG_V = 32, WG_CHUNK = 4096;      // 32 ci x 32 co tile, 32 voxels per smem step, voxels per CTA

struct WgP {
  const float* x;      // (B,Cin,Di,Hi,Wi)
  const float* dy;     // (B,Cout,Do,Ho,Wo)
  float* dw;           // [K^3][Cin][Cout], zeroed by the caller
  int B, Cin, Cout, Di, Hi, Wi, Do, Ho, Wo, K, stride, pad;
  long long M;         // B*Do*Ho*Wo
  int tiles_co;
};

// 64 threads per CTA, each owns a 4 (ci) x 4 (co) register tile of the 32 x 32 block: per staged voxel a thread reads 4 + 4 shared
// values (broadcast within the warp: 4 / 8 distinct addresses) for 16 FMAs.  (Round 1's version gave every one of 256 threads a 2 x 2
// tile -- one FMA per shared load, 8 TFLOP/s on the 32 -> 32 layer.)  The staged tiles are [voxel][channel], so the 4 + 4 values are two
// 128-bit loads for 16 FMAs.
__global__ void __launch_bounds__(64) conv3d_wgrad_kernel(const WgP p) {
  __shared__ __align__(16) float Xs[WG_V][WG_T + 4];                   // [voxel][channel]: a thread's 4 channels are one 128-bit load
  __shared__ __align__(16) float Ys[WG_V][WG_T + 4];
  const int tap = blockIdx.y;
  const int kd = tap / (p.K * p.K), kh = (tap / p.K) % p.K, kw = tap % p.K;
  const int ci0 = (blockIdx.z / p.tiles_co) * WG_T, co0 = (blockIdx.z % p.tiles_co) * WG_T;
  const int tid = threadIdx.x, ti = tid >> 3, tj = tid & 7;            // thread owns ci 4*ti .. +3  x  co 4*tj .. +3
  const long long m0 = (long long)blockIdx.x * WG_CHUNK, m1 = min(p.M, m0 + WG_CHUNK);
  const size_t in_cs = (size_t)p.Di * p.Hi * p.Wi, out_cs = (size_t)p.Do * p.Ho * p.Wo;
  const int lv = tid & 31, lc = tid >> 5;                              // loader: voxel lv, channels lc, lc+2, ..., lc+30
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
  for (long long mb = m0; mb < m1; mb += WG_V) {
    const long long m = mb + lv;
    bool ok = m < m1;
    long long r = ok ? m : 0;
    const int ow = (int)(r % p.Wo); r /= p.Wo;
    const int oh = (int)(r % p.Ho); r /= p.Ho;
    const int od = (int)(r % p.Do);
    const int b = (int)(r / p.Do);
    const int di = od * p.stride - p.pad + kd, hi = oh * p.stride - p.pad + kh, wi = ow * p.stride - p.pad + kw;
    const bool xin = ok && di >= 0 && di < p.Di && hi >= 0 && hi < p.Hi && wi >= 0 && wi < p.Wi;
    const size_t xo = ((size_t)(xin ? di : 0) * p.Hi + (xin ? hi : 0)) * p.Wi + (xin ? wi : 0);
    const size_t yo = ((size_t)od * p.Ho + oh) * p.Wo + ow;
    const float* xp = p.x + ((size_t)b * p.Cin + ci0 + lc) * in_cs + xo;
    const float* yp = p.dy + ((size_t)b * p.Cout + co0 + lc) * out_cs + yo;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int c = lc + 2 * q;
      Xs[lv][c] = (xin && ci0 + c < p.Cin) ? __ldg(xp + (size_t)(2 * q) * in_cs) : 0.0f;
      Ys[lv][c] = (ok && co0 + c < p.Cout) ? __ldg(yp + (size_t)(2 * q) * out_cs) : 0.0f;
    }
    __syncthreads();
#pragma unroll 8
    for (int v = 0; v < WG_V; ++v) {
      const float4 x4 = *reinterpret_cast<const float4*>(&Xs[v][4 * ti]), y4 = *reinterpret_cast<const float4*>(&Ys[v][4 * tj]);
      const float xv[4] = {x4.x, x4.y, x4.z, x4.w}, yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(xv[a], yv[c], acc[a][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int ci = ci0 + 4 * ti + a, co = co0 + 4 * tj + c;
      if (ci < p.Cin && co < p.Cout) atomicAdd(p.dw + ((size_t)tap * p.Cin + ci) * p.Cout + co, acc[a][c]);
    }
}

// ---- BatchNorm with batch statistics ------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_d(double v, double* red) {      // 256 threads
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < 8; ++i) t += red[i];
  return t;
}

// Per-channel reductions over the B*S elements of a channel, split over BN_SPLIT CTAs per channel (a full-resolution SSR map has
// 1-6 channels of 2M elements: one CTA per channel left the GPU empty).  Partials are doubles; a finalize kernel combines them.
constexpr int BN_SPLIT = 64;

// partial[c][s] = (sum a, sum b) with (a, b) = (x, x^2) [MODE 0]  or  (dy, dy * xhat) [MODE 1]
template <int MODE>
__global__ void __launch_bounds__(256) bn_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean,
                                                         const float* __restrict__ var, double2* __restrict__ partial, int B, int C, size_t S,
                                                         float eps) {
  __shared__ double red[8];
  const int c = blockIdx.y, sp = blockIdx.x;
  const size_t n = (size_t)B * S;
  const size_t per = (n + BN_SPLIT - 1) / BN_SPLIT, lo = sp * per, hi = min(n, lo + per);
  float mu = 0.f, rstd = 1.f;
  if (MODE == 1) { mu = __ldg(mean + c); rstd = rsqrtf(__ldg(var + c) + eps); }
  double a = 0.0, b2 = 0.0;
  float fa = 0.f, fb = 0.f;
  int cnt = 0;
  for (size_t i = lo + threadIdx.x; i < hi; i += 256) {
    const size_t o = ((i / S) * C + c) * S + i % S;
    const float xv = __ldg(x + o);
    if (MODE == 0) { fa += xv; fb = fmaf(xv, xv, fb); }
    else { const float g = __ldg(dy + o); fa += g; fb = fmaf(g, (xv - mu) * rstd, fb); }
    if (++cnt == 64) { a += fa; b2 += fb; fa = fb = 0.f; cnt = 0; }      // flush the fp32 partials into doubles every 64 elements
  }
  a += fa; b2 += fb;
  const double sa = block_sum_d(a, red);
  const double sb = block_sum_d(b2, red);
  if (threadIdx.x == 0) partial[(size_t)c * BN_SPLIT + sp] = make_double2(sa, sb);
}

// MODE 0: mean, biased variance.  MODE 1: out_a = sum dy (grad_bias), out_b = sum dy * xhat (grad_weight)
template <int MODE>
__global__ void __launch_bounds__(64) bn_finalize_kernel(const double2* __restrict__ partial, float* __restrict__ out_a, float* __restrict__ out_b,
                                                         double n) {
  const int c = blockIdx.x;
  double2 v = threadIdx.x < BN_SPLIT ? partial[(size_t)c * BN_SPLIT + threadIdx.x] : make_double2(0.0, 0.0);
  for (int o = 16; o > 0; o >>= 1) { v.x += __shfl_xor_sync(0xffffffffu, v.x, o); v.y += __shfl_xor_sync(0xffffffffu, v.y, o); }
  __shared__ double2 r[2];
  if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    const double a = r[0].x + r[1].x, b = r[0].y + r[1].y;
    if (MODE == 0) { const double mu = a / n; out_a[c] = (float)mu; out_b[c] = (float)fmax(b / n - mu * mu, 0.0); }
    else { out_a[c] = (float)a; out_b[c] = (float)b; }
  }
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ var,
                                                       const float* __restrict__ weight, const float* __restrict__ bias, float* __restrict__ out,
                                                       int C, size_t S, size_t total, float eps, int relu) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)((i / S) % C);
  const float rstd = rsqrtf(__ldg(var + c) + eps);
  float y = (x[i] - __ldg(mean + c)) * rstd * (weight ? __ldg(weight + c) : 1.0f) + (bias ? __ldg(bias + c) : 0.0f);
  out[i] = relu ? fmaxf(y, 0.0f) : y;
}

__global__ void __launch_bounds__(256) bn_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean,
                                                        const float* __restrict__ var, const float* __restrict__ weight,
                                                        const float* __restrict__ dweight, const float* __restrict__ dbias, float* __restrict__ dx,
                                                        int C, size_t S, size_t total, float inv_n, float eps) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)((i / S) % C);
  const float rstd = rsqrtf(__ldg(var + c) + eps), xhat = (x[i] - __ldg(mean + c)) * rstd;
  const float w = weight ? __ldg(weight + c) : 1.0f;
  dx[i] = w * rstd * (dy[i] - inv_n * (__ldg(dbias + c) + xhat * __ldg(dweight + c)));
}

}  // namespace

extern "C" int ss_conv3d_wgrad_f32(const float* x, const float* grad_out, float* grad_weight_packed, int B, int Cin, int Cout, int Di, int Hi,
                                   int Wi, int K, int stride, void* stream) {
  SS_REQUIRE(x && grad_out && grad_weight_packed, "ss_conv3d_wgrad_f32: null pointer");
  SS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && Di > 0 && Hi > 0 && Wi > 0, "ss_conv3d_wgrad_f32: non-positive dimension");
  SS_UNSUPPORTED(!(K == 1 || K == 3) || !(stride == 1 || stride == 2), "ss_conv3d_wgrad_f32: K in {1,3}, stride in {1,2} (got K=%d, stride=%d)", K, stride);
  WgP p;
  p.x = x; p.dy = grad_out; p.dw = grad_weight_packed;
  p.B = B; p.Cin = Cin; p.Cout = Cout; p.Di = Di; p.Hi = Hi; p.Wi = Wi; p.K = K; p.stride = stride; p.pad = K / 2;
  p.Do = (Di + 2 * p.pad - K) / stride + 1; p.Ho = (Hi + 2 * p.pad - K) / stride + 1; p.Wo = (Wi + 2 * p.pad - K) / stride + 1;
  p.M = (long long)B * p.Do * p.Ho * p.Wo;
  p.tiles_co = ceil_div(Cout, WG_T);
  const long long chunks = ceil_div64(p.M, WG_CHUNK);
  const int tiles = ceil_div(Cin, WG_T) * p.tiles_co;
  SS_UNSUPPORTED(chunks > 0x7fffffffLL || tiles > 65535, "ss_conv3d_wgrad_f32: grid dimension too large");
  conv3d_wgrad_kernel<<<dim3((unsigned)chunks, K * K * K, tiles), 64, 0, (cudaStream_t)stream>>>(p);
  SS_CHECK_LAUNCH("ss_conv3d_wgrad_f32");
  return SS_OK;
}

extern "C" int ss_bn_workspace_bytes(int C) { return C * BN_SPLIT * (int)sizeof(double2); }

extern "C" int ss_bn_train_forward(const float* x, const float* weight_or_null, const float* bias_or_null, float* out, float* batch_mean,
                                   float* batch_var, void* workspace, int B, int C, long long S, float eps, int relu, void* stream) {
  SS_REQUIRE(x && out && batch_mean && batch_var && B > 0 && C > 0 && S > 0, "ss_bn_train_forward: bad argument");
  SS_UNSUPPORTED((long long)B * S < 2, "ss_bn_train_forward: batch statistics need more than one value per channel");
  cudaStream_t st = (cudaStream_t)stream;
  SS_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "ss_bn_train_forward: 16-byte aligned workspace required");
  double2* part = reinterpret_cast<double2*>(workspace);
  bn_partial_kernel<0><<<dim3(BN_SPLIT, C), 256, 0, st>>>(x, nullptr, nullptr, nullptr, part, B, C, (size_t)S, eps);
  SS_CHECK_LAUNCH("ss_bn_train_forward(partial)");
  bn_finalize_kernel<0><<<C, 64, 0, st>>>(part, batch_mean, batch_var, (double)B * (double)S);
  SS_CHECK_LAUNCH("ss_bn_train_forward(stats)");
  const size_t total = (size_t)B * C * S;
  bn_apply_kernel<<<(unsigned)ceil_div64((int64_t)total, 256), 256, 0, st>>>(x, batch_mean, batch_var, weight_or_null, bias_or_null, out, C, (size_t)S,
                                                                           total, eps, relu);
  SS_CHECK_LAUNCH("ss_bn_train_forward(apply)");
  return SS_OK;
}

extern "C" int ss_bn_train_backward(const float* x, const float* grad_out, const float* batch_mean, const float* batch_var,
                                    const float* weight_or_null, float* grad_x, float* grad_weight, float* grad_bias, void* workspace, int B, int C,
                                    long long S, float eps, void* stream) {
  SS_REQUIRE(x && grad_out && batch_mean && batch_var && grad_x && grad_weight && grad_bias && B > 0 && C > 0 && S > 0,
             "ss_bn_train_backward: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  SS_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "ss_bn_train_backward: 16-byte aligned workspace required");
  double2* part = reinterpret_cast<double2*>(workspace);
  bn_partial_kernel<1><<<dim3(BN_SPLIT, C), 256, 0, st>>>(x, grad_out, batch_mean, batch_var, part, B, C, (size_t)S, eps);
  SS_CHECK_LAUNCH("ss_bn_train_backward(partial)");
  bn_finalize_kernel<1><<<C, 64, 0, st>>>(part, grad_bias, grad_weight, 1.0);
  SS_CHECK_LAUNCH("ss_bn_train_backward(reduce)");
  const size_t total = (size_t)B * C * S;
  bn_bwd_dx_kernel<<<(unsigned)ceil_div64((int64_t)total, 256), 256, 0, st>>>(x, grad_out, batch_mean, batch_var, weight_or_null, grad_weight, grad_bias,
                                                                            grad_x, C, (size_t)S, total, 1.0f / (float)((double)B * (double)S), eps);
  SS_CHECK_LAUNCH("ss_bn_train_backward(dx)");
  return SS_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Small-channel 2-D convolutions (Cin, Cout <= 8; k in {1, 3}, stride 1, zero padding k/2): the four convs of SSR_upsample
// (models/submodule.py:394-431: 1 -> 6 3x3, 6 -> 6 and 6 -> 1 1x1) at FULL image resolution in training mode.  Through the generic
// implicit-GEMM kernels (a depth-1 volume, channels padded to the GEMM tile) they cost 2.3 ms per 3x3 launch and 7.6 ms for its
// weight gradient at two 1024^2 images -- 20 % of the config-5 training step for 0.1 GFLOP.  Here: one thread per pixel, every output
// channel in registers; dX is the same kernel on dY with flipped / transposed weights (done by the caller); dW accumulates all
// Cout*Cin*k*k sums per thread over a strided set of pixels, then warp shuffles, shared memory and one atomicAdd per block and weight.
// ---------------------------------------------------------------------------------------------------------------------
namespace {

template <int K>
__global__ void __launch_bounds__(256) small_conv2d_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out,
                                                           int Cin, int Cout, int H, int W) {
  __shared__ float sw[8 * 8 * K * K];
  for (int i = threadIdx.x; i < Cout * Cin * K * K; i += blockDim.x) sw[i] = __ldg(w + i);
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  const size_t HW = (size_t)H * W;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int ci = 0; ci < Cin; ++ci) {
    const float* ip = in + ((size_t)b * Cin + ci) * HW;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const int yy = y + ky - K / 2;
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int xx = x + kx - K / 2;
        const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(ip + (size_t)yy * W + xx) : 0.0f;
#pragma unroll
        for (int co = 0; co < 8; ++co)
          if (co < Cout) acc[co] = fmaf(sw[((co * Cin + ci) * K + ky) * K + kx], v, acc[co]);
      }
    }
  }
  for (int co = 0; co < Cout; ++co) out[((size_t)b * Cout + co) * HW + (size_t)y * W + x] = acc[co];
}

template <int CIN, int COUT, int K>
__global__ void __launch_bounds__(256) small_conv2d_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                                                                 int B, int H, int W) {
  constexpr int NW = COUT * CIN * K * K;
  __shared__ float red[8][NW];
  const size_t HW = (size_t)H * W, total = (size_t)B * HW;
  float acc[NW];
#pragma unroll
  for (int i = 0; i < NW; ++i) acc[i] = 0.f;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(p / HW);
    const size_t pix = p - (size_t)b * HW;
    const int yy0 = (int)(pix / W), xx0 = (int)(pix - (size_t)yy0 * W);
    float g[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) g[co] = __ldg(dy + ((size_t)b * COUT + co) * HW + pix);
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float* ip = x + ((size_t)b * CIN + ci) * HW;
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int yy = yy0 + ky - K / 2;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const int xx = xx0 + kx - K / 2;
          const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(ip + (size_t)yy * W + xx) : 0.0f;
#pragma unroll
          for (int co = 0; co < COUT; ++co) acc[((co * CIN + ci) * K + ky) * K + kx] = fmaf(g[co], v, acc[((co * CIN + ci) * K + ky) * K + kx]);
        }
      }
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid][i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NW; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[k][i];
    atomicAdd(dw + i, v);
  }
}

}  // namespace

extern "C" int ss_conv2d_small_f32(const float* in, const float* weight, float* out, int B, int Cin, int Cout, int H, int W, int k, void* stream) {
  SS_REQUIRE(in && weight && out && B > 0 && H > 0 && W > 0, "ss_conv2d_small_f32: bad argument");
  SS_UNSUPPORTED(Cin < 1 || Cin > 8 || Cout < 1 || Cout > 8 || (k != 1 && k != 3), "ss_conv2d_small_f32: Cin, Cout in [1, 8] and k in {1, 3} only");
  SS_UNSUPPORTED(H > 65535 || B > 65535, "ss_conv2d_small_f32: grid dimension exceeds 65535");
  const dim3 grid(ceil_div(W, 256), H, B);
  if (k == 3) small_conv2d_kernel<3><<<grid, 256, 0, (cudaStream_t)stream>>>(in, weight, out, Cin, Cout, H, W);
  else small_conv2d_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(in, weight, out, Cin, Cout, H, W);
  SS_CHECK_LAUNCH("ss_conv2d_small_f32");
  return SS_OK;
}

// 1 when ss_conv2d_small_wgrad_f32 has an instantiation for this layer (the SSR_upsample shapes), else 0
extern "C" int ss_conv2d_small_wgrad_supported(int Cin, int Cout, int k) {
  return (Cin == 1 && Cout == 6 && k == 3) || (Cin == 6 && Cout == 6 && k == 1) || (Cin == 6 && Cout == 1 && k == 1) ? 1 : 0;
}

// grad_weight (Cout, Cin, k, k) is ACCUMULATED into (atomicAdd): zero it first
extern "C" int ss_conv2d_small_wgrad_f32(const float* x, const float* grad_out, float* grad_weight, int B, int Cin, int Cout, int H, int W, int k,
                                         void* stream) {
  SS_REQUIRE(x && grad_out && grad_weight && B > 0 && H > 0 && W > 0, "ss_conv2d_small_wgrad_f32: bad argument");
  SS_UNSUPPORTED(!ss_conv2d_small_wgrad_supported(Cin, Cout, k), "ss_conv2d_small_wgrad_f32: no instantiation for (Cin=%d, Cout=%d, k=%d)", Cin, Cout, k);
  const int grid = ss_num_sms() * 4;
  cudaStream_t st = (cudaStream_t)stream;
  if (k == 3) small_conv2d_wgrad_kernel<1, 6, 3><<<grid, 256, 0, st>>>(x, grad_out, grad_weight, B, H, W);
  else if (Cout == 6) small_conv2d_wgrad_kernel<6, 6, 1><<<grid, 256, 0, st>>>(x, grad_out, grad_weight, B, H, W);
  else small_conv2d_wgrad_kernel<6, 1, 1><<<grid, 256, 0, st>>>(x, grad_out, grad_weight, B, H, W);
  SS_CHECK_LAUNCH("ss_conv2d_small_wgrad_f32");
  return SS_OK;
}
