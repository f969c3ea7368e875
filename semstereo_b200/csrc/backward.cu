// Backward (vector-Jacobian) kernels of the cost-volume / regression operators of the reference surface -- the first step of
// BASELINE config #5 (build-plan item 10): the operators of models/submodule.py that SemStereo.py calls as
// free functions become differentiable through torch.library (semstereo_b200/torch_ops.py), so a caller that trains the reference
// model can keep them on the sm_100a path.  All kernels are gather-form (every gradient element is written exactly once: no
// atomics, deterministic), fp32, and HBM-bound like their forward twins.
//   gwc_volume_backward          build_gwc_volume / build_gwc_volume_norm (submodule.py:198-238, submodule_.py:188-221)
//   concat_volume_backward       build_concat_volume (submodule.py:173-187, submodule_.py:166-178)
//   disparity_regression_backward  disparity_regression (submodule.py:164-170)
//   regression_topk_backward     regression_topk (submodule.py:434-442); the sort indices carry no gradient
//   context_upsample_backward    context_upsample (submodule_.py:311-323)
#include "common.cuh"

namespace {

constexpr float kEpsNorm = 1e-5f;   // groupwise_correlation_norm adds it to the L2 norm (submodule.py:218)

// One CTA = one (b, group, image row).  Shared memory: Lh, Rh [cg][W] (normalised when `norm`), nL, nR [W] (the norms), dV [D][W].
// dLh[c,x]  = 1/cg * sum_k dV[k,x]      * Rh[c, x - d_k]      (0 <= x - d_k < W)
// dRh[c,x'] = 1/cg * sum_k dV[k,x'+d_k] * Lh[c, x' + d_k]     (0 <= x' + d_k < W)
// norm: with n = ||L_g|| and Lh = L / (n + eps):  dL = dLh / (n + eps) - L * <dLh, L> / (n * (n + eps)^2)   (0 where n == 0)
__global__ void __launch_bounds__(128) gwc_volume_backward_kernel(const float* __restrict__ left, const float* __restrict__ right,
                                                                  const float* __restrict__ gvol, float* __restrict__ gleft,
                                                                  float* __restrict__ gright, int C, int H, int W, int G, int D,
                                                                  int dmin, int norm) {
  extern __shared__ float sm[];
  const int cg = C / G;
  float* Lh = sm;
  float* Rh = Lh + (size_t)cg * W;
  float* nL = Rh + (size_t)cg * W;
  float* nR = nL + W;
  float* dV = nR + W;
  const int y = blockIdx.x, g = blockIdx.y % G, b = blockIdx.y / G;
  const size_t HW = (size_t)H * W;
  const float* lb = left + ((size_t)b * C + (size_t)g * cg) * HW + (size_t)y * W;
  const float* rb = right + ((size_t)b * C + (size_t)g * cg) * HW + (size_t)y * W;
  for (int i = threadIdx.x; i < cg * W; i += blockDim.x) {
    const int c = i / W, x = i - c * W;
    Lh[i] = __ldg(lb + (size_t)c * HW + x);
    Rh[i] = __ldg(rb + (size_t)c * HW + x);
  }
  const float* gv = gvol + ((size_t)b * G + g) * D * HW + (size_t)y * W;
  for (int i = threadIdx.x; i < D * W; i += blockDim.x) {
    const int k = i / W, x = i - k * W;
    dV[i] = __ldg(gv + (size_t)k * HW + x);
  }
  __syncthreads();
  if (norm) {
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
      float sl = 0.0f, sr = 0.0f;
      for (int c = 0; c < cg; ++c) { sl = fmaf(Lh[c * W + x], Lh[c * W + x], sl); sr = fmaf(Rh[c * W + x], Rh[c * W + x], sr); }
      nL[x] = sqrtf(sl); nR[x] = sqrtf(sr);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cg * W; i += blockDim.x) {
      const int x = i % W;
      Lh[i] /= (nL[x] + kEpsNorm);
      Rh[i] /= (nR[x] + kEpsNorm);
    }
    __syncthreads();
  }
  const float inv_cg = 1.0f / (float)cg;
  float* glb = gleft + ((size_t)b * C + (size_t)g * cg) * HW + (size_t)y * W;
  float* grb = gright + ((size_t)b * C + (size_t)g * cg) * HW + (size_t)y * W;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    float sL = 0.0f, sR = 0.0f;                       // <dLh, Lh>, <dRh, Rh> over the group's channels (norm only)
    for (int c = 0; c < cg; ++c) {
      float dl = 0.0f, dr = 0.0f;
      for (int k = 0; k < D; ++k) {
        const int d = dmin + k;
        const int xr = x - d, xl = x + d;
        if (xr >= 0 && xr < W) dl = fmaf(dV[k * W + x], Rh[c * W + xr], dl);
        if (xl >= 0 && xl < W) dr = fmaf(dV[k * W + xl], Lh[c * W + xl], dr);
      }
      dl *= inv_cg; dr *= inv_cg;
      if (norm) {
        sL = fmaf(dl, Lh[c * W + x], sL);
        sR = fmaf(dr, Rh[c * W + x], sR);
      }
      glb[(size_t)c * HW + x] = dl;                   // provisional (final when !norm); re-read by this same thread below
      grb[(size_t)c * HW + x] = dr;
    }
    if (norm) {
      // L = Lh*(n+eps):  dL = (dLh - Lh * <dLh,Lh> * (n+eps)/n) / (n+eps)
      const float nl = nL[x], nr = nR[x];
      const float il = 1.0f / (nl + kEpsNorm), ir = 1.0f / (nr + kEpsNorm);
      const float fl = nl > 0.0f ? sL * (nl + kEpsNorm) / nl : 0.0f, fr = nr > 0.0f ? sR * (nr + kEpsNorm) / nr : 0.0f;
      for (int c = 0; c < cg; ++c) {
        glb[(size_t)c * HW + x] = (glb[(size_t)c * HW + x] - Lh[c * W + x] * fl) * il;
        grb[(size_t)c * HW + x] = (grb[(size_t)c * HW + x] - Rh[c * W + x] * fr) * ir;
      }
    }
  }
}

// volume (B,2C,D,H,W): [:C] = left * valid(x - d) (signed) or left (unsigned: the left half is not masked), [C:] = right[x - d].
__global__ void __launch_bounds__(256) concat_volume_backward_kernel(const float* __restrict__ gvol, float* __restrict__ gleft,
                                                                     float* __restrict__ gright, int C, int H, int W, int D, int dmin,
                                                                     int mask_left) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const int y = blockIdx.y % H, c = blockIdx.y / H, b = blockIdx.z;
  const size_t HW = (size_t)H * W;
  const float* gl = gvol + ((size_t)b * 2 * C + c) * D * HW + (size_t)y * W;
  const float* gr = gvol + ((size_t)b * 2 * C + C + c) * D * HW + (size_t)y * W;
  float al = 0.0f, ar = 0.0f;
  for (int k = 0; k < D; ++k) {
    const int d = dmin + k;
    const int xs = x - d;                             // the column of `right` that lands at x; valid(x - d) for the left mask
    if (!mask_left || (xs >= 0 && xs < W)) al += __ldg(gl + (size_t)k * HW + x);
    const int xt = x + d;                             // where right[x] lands
    if (xt >= 0 && xt < W) ar += __ldg(gr + (size_t)k * HW + xt);
  }
  gleft[((size_t)b * C + c) * HW + (size_t)y * W + x] = al;
  gright[((size_t)b * C + c) * HW + (size_t)y * W + x] = ar;
}

__global__ void __launch_bounds__(256) disparity_regression_backward_kernel(const float* __restrict__ gout, float* __restrict__ gprob,
                                                                            int D, size_t HW, float dmin) {
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (pix >= HW) return;
  const float g = __ldg(gout + (size_t)b * HW + pix);
  for (int k = 0; k < D; ++k) gprob[((size_t)b * D + k) * HW + pix] = g * (dmin + (float)k);
}

// pred = sum_{i in top-K} p_i s_i, p = softmax(cost over the top-K): dcost_i = g p_i (s_i - pred), dsample_i = g p_i, 0 elsewhere.
template <int NB>
__global__ void __launch_bounds__(128) regression_topk_backward_kernel(const float* __restrict__ cost, const float* __restrict__ samples,
                                                                       const float* __restrict__ gpred, float* __restrict__ gcost,
                                                                       float* __restrict__ gsamp, int D, int K, size_t HW) {
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (pix >= HW) return;
  float c[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) c[k] = (k < D) ? __ldg(cost + ((size_t)b * D + k) * HW + pix) : -INFINITY;
  unsigned long long taken = 0ull;
  float top = -INFINITY;
  for (int r = 0; r < K; ++r) {                       // same selection (and tie rule) as the forward kernel
    float best = -INFINITY;
    int bi = 0;
    bool found = false;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const bool free_ = k < D && !((taken >> k) & 1ull);
      if (free_ && (!found || c[k] > best)) { best = c[k]; bi = k; found = true; }
    }
    taken |= 1ull << bi;
    if (r == 0) top = best;
  }
  float sum = 0.0f, acc = 0.0f;
#pragma unroll
  for (int k = 0; k < NB; ++k)
    if ((taken >> k) & 1ull) {
      const float e = expf(c[k] - top);
      sum += e;
      acc += e * __ldg(samples + ((size_t)b * D + k) * HW + pix);
    }
  const float pred = acc / sum, g = __ldg(gpred + (size_t)b * HW + pix);
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    if (k >= D) continue;
    float gc = 0.0f, gs = 0.0f;
    if ((taken >> k) & 1ull) {
      const float pk = expf(c[k] - top) / sum;
      gs = g * pk;
      gc = gs * (__ldg(samples + ((size_t)b * D + k) * HW + pix) - pred);
    }
    gcost[((size_t)b * D + k) * HW + pix] = gc;
    gsamp[((size_t)b * D + k) * HW + pix] = gs;
  }
}

// out[Y,X] = sum_t w[t,Y,X] * depth[Y/4 + ky - 1, X/4 + kx - 1]   (t = ky*3 + kx, zero padded)
__global__ void __launch_bounds__(256) context_upsample_backward_w_kernel(const float* __restrict__ depth_low, const float* __restrict__ gout,
                                                                          float* __restrict__ gw, int h, int w) {
  const int H = 4 * h, W = 4 * w;
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y, b = blockIdx.z;
  if (X >= W) return;
  const float* dl = depth_low + (size_t)b * h * w;
  const size_t HW = (size_t)H * W, pix = (size_t)Y * W + X;
  const float g = __ldg(gout + (size_t)b * HW + pix);
  const int y = Y >> 2, x = X >> 2;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = y + ky - 1, xx = x + kx - 1;
      const float dv = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(dl + yy * w + xx) : 0.0f;
      gw[((size_t)b * 9 + ky * 3 + kx) * HW + pix] = g * dv;
    }
}
__global__ void __launch_bounds__(128) context_upsample_backward_d_kernel(const float* __restrict__ upw, const float* __restrict__ gout,
                                                                          float* __restrict__ gdepth, int h, int w) {
  const int H = 4 * h, W = 4 * w;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= w) return;
  const size_t HW = (size_t)H * W;
  float acc = 0.0f;
  // depth[y,x] is read by the 4x4 output block of low-res pixel (y - ky + 1, x - kx + 1) through tap (ky,kx)
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yo = y - ky + 1, xo = x - kx + 1;
      if (yo < 0 || yo >= h || xo < 0 || xo >= w) continue;
      const float* wp = upw + ((size_t)b * 9 + ky * 3 + kx) * HW;
      const float* gp = gout + (size_t)b * HW;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const size_t row = (size_t)(4 * yo + i) * W + 4 * xo;
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wp + row)), gv = __ldg(reinterpret_cast<const float4*>(gp + row));
        acc += wv.x * gv.x + wv.y * gv.y + wv.z * gv.z + wv.w * gv.w;
      }
    }
  gdepth[((size_t)b * h + y) * w + x] = acc;
}

// Propagation / Propagation_prob (submodule.py:290-307, 361-377): out_s[y,x] = in[clamp(y+dy_s), clamp(x+dx_s)].
// Gather form: pixel (y',x') collects g_s[y,x] from every (y,x) of its 3x3 neighbourhood whose clamped tap lands on it.
__global__ void __launch_bounds__(256) propagation_backward_kernel(const float* __restrict__ gout, float* __restrict__ gin, int D, int H,
                                                                   int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const int y = blockIdx.y % H, d = blockIdx.y / H, b = blockIdx.z;
  const size_t HW = (size_t)H * W, S = (size_t)D * HW;
  float acc = 0.0f;
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const float* g = gout + ((size_t)b * 5 + s) * S + (size_t)d * HW;
#pragma unroll
    for (int oy = -1; oy <= 1; ++oy)
#pragma unroll
      for (int ox = -1; ox <= 1; ++ox) {
        const int yy = y + oy, xx = x + ox;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        if (min(max(yy + kPropDy[s], 0), H - 1) == y && min(max(xx + kPropDx[s], 0), W - 1) == x) acc += __ldg(g + (size_t)yy * W + xx);
      }
  }
  gin[(size_t)b * S + (size_t)d * HW + (size_t)y * W + x] = acc;
}

// disparity_variance (submodule.py:257-263): var = sum_k p_k (d_k - mu)^2.
__global__ void __launch_bounds__(256) disparity_variance_backward_kernel(const float* __restrict__ p, const float* __restrict__ mu,
                                                                          const float* __restrict__ gout, float* __restrict__ gp,
                                                                          float* __restrict__ gmu, int D, size_t HW, float dmin) {
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (pix >= HW) return;
  const float m = __ldg(mu + (size_t)b * HW + pix), g = __ldg(gout + (size_t)b * HW + pix);
  float acc = 0.0f;
  for (int k = 0; k < D; ++k) {
    const float e = dmin + (float)k - m;
    gp[((size_t)b * D + k) * HW + pix] = g * e * e;
    acc = fmaf(__ldg(p + ((size_t)b * D + k) * HW + pix), e, acc);
  }
  gmu[(size_t)b * HW + pix] = -2.0f * g * acc;
}

// SpatialTransformer_grid (submodule.py:265-288).  One thread per (b, k, y, x):
//   grad_disp = -sum_c g_yw[c] * d(bilinear)/d(ix)   (ix = x - d up to the fp32 round trip: d ix / d d = -1)
//   grad_src  : the four corners receive g_yw[c] * weight   (scatter: atomicAdd, like torch's grid_sampler backward)
// grad_x (the repeated left features) is the sum of g_xrep over k and is done by the caller-facing wrapper kernel below.
__global__ void __launch_bounds__(128) stn_backward_kernel(const float* __restrict__ src, const float* __restrict__ disp,
                                                           const float* __restrict__ gyw, float* __restrict__ gsrc,
                                                           float* __restrict__ gdisp, int C, int K, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const int y = blockIdx.y % H, k = blockIdx.y / H, b = blockIdx.z;
  const size_t HW = (size_t)H * W;
  const float d = __ldg(disp + ((size_t)b * K + k) * HW + (size_t)y * W + x);
  const float ix = warp_coord((float)x - d, (float)(W - 1)), iy = warp_coord((float)y, (float)(H - 1));
  const Bilin q = make_bilin(ix, iy, H, W);
  const float fy0 = floorf(iy), wy0 = fy0 + 1.0f - iy, wy1 = iy - fy0;      // the y factors of the four weights
  float gd = 0.0f;
  for (int c = 0; c < C; ++c) {
    const float g = __ldg(gyw + (((size_t)b * C + c) * K + k) * HW + (size_t)y * W + x);
    const float* plane = src + ((size_t)b * C + c) * HW;
    float* gplane = gsrc + ((size_t)b * C + c) * HW;
    const float a00 = __ldg(plane + q.o00), a01 = __ldg(plane + q.o01), a10 = __ldg(plane + q.o10), a11 = __ldg(plane + q.o11);
    if (q.w00 != 0.0f) atomicAdd(gplane + q.o00, g * q.w00);
    if (q.w01 != 0.0f) atomicAdd(gplane + q.o01, g * q.w01);
    if (q.w10 != 0.0f) atomicAdd(gplane + q.o10, g * q.w10);
    if (q.w11 != 0.0f) atomicAdd(gplane + q.o11, g * q.w11);
    // d out / d ix = (ne - nw) * wy0 + (se - sw) * wy1 with out-of-range corners contributing 0
    const float fx0 = floorf(ix), fx1 = fx0 + 1.0f;
    const bool x0ok = fx0 >= 0.0f && fx0 <= (float)(W - 1), x1ok = fx1 >= 0.0f && fx1 <= (float)(W - 1);
    const bool y0ok = fy0 >= 0.0f && fy0 <= (float)(H - 1), y1ok = fy0 + 1.0f >= 0.0f && fy0 + 1.0f <= (float)(H - 1);
    const float nw = (x0ok && y0ok) ? a00 : 0.0f, ne = (x1ok && y0ok) ? a01 : 0.0f;
    const float sw = (x0ok && y1ok) ? a10 : 0.0f, se = (x1ok && y1ok) ? a11 : 0.0f;
    gd = fmaf(g, (ne - nw) * wy0 + (se - sw) * wy1, gd);
  }
  gdisp[((size_t)b * K + k) * HW + (size_t)y * W + x] = -gd;
}

__global__ void __launch_bounds__(256) sum_over_k_kernel(const float* __restrict__ gxr, float* __restrict__ gx, int K, size_t HW) {
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t bc = blockIdx.y;
  if (pix >= HW) return;
  float acc = 0.0f;
  for (int k = 0; k < K; ++k) acc += __ldg(gxr + (bc * K + k) * HW + pix);
  gx[bc * HW + pix] = acc;
}

}  // namespace

extern "C" int ss_gwc_volume_backward(const float* left, const float* right, const float* grad_volume, float* grad_left, float* grad_right,
                                      int B, int C, int H, int W, int maxdisp, int num_groups, int flags, void* stream) {
  SS_REQUIRE(left && right && grad_volume && grad_left && grad_right, "ss_gwc_volume_backward: null pointer");
  SS_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && maxdisp > 0 && num_groups > 0, "ss_gwc_volume_backward: non-positive dimension");
  SS_REQUIRE(C % num_groups == 0, "ss_gwc_volume_backward: C=%d is not a multiple of num_groups=%d", C, num_groups);
  const int D = (flags & SS_SIGNED) ? 2 * maxdisp : maxdisp, dmin = (flags & SS_SIGNED) ? -maxdisp : 0, cg = C / num_groups;
  const size_t smem = ((size_t)2 * cg * W + 2 * (size_t)W + (size_t)D * W) * sizeof(float);
  SS_UNSUPPORTED(smem > 200 * 1024, "ss_gwc_volume_backward: row working set of %zu bytes exceeds shared memory", smem);
  SS_UNSUPPORTED(H > 65535 * 32 || (long long)B * num_groups > 65535, "ss_gwc_volume_backward: grid dimension exceeds the limit");
  SS_CUDA(ss_allow_smem(gwc_volume_backward_kernel, smem));
  gwc_volume_backward_kernel<<<dim3(H, B * num_groups), 128, smem, (cudaStream_t)stream>>>(left, right, grad_volume, grad_left, grad_right,
                                                                                         C, H, W, num_groups, D, dmin, (flags & SS_NORM) ? 1 : 0);
  SS_CHECK_LAUNCH("ss_gwc_volume_backward");
  return SS_OK;
}

extern "C" int ss_concat_volume_backward(const float* grad_volume, float* grad_left, float* grad_right, int B, int C, int H, int W,
                                         int maxdisp, int flags, void* stream) {
  SS_REQUIRE(grad_volume && grad_left && grad_right, "ss_concat_volume_backward: null pointer");
  SS_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && maxdisp > 0, "ss_concat_volume_backward: non-positive dimension");
  SS_UNSUPPORTED((long long)C * H > 65535 || B > 65535, "ss_concat_volume_backward: grid dimension exceeds 65535");
  const int sg = (flags & SS_SIGNED) ? 1 : 0;
  concat_volume_backward_kernel<<<dim3(ceil_div(W, 256), C * H, B), 256, 0, (cudaStream_t)stream>>>(
      grad_volume, grad_left, grad_right, C, H, W, sg ? 2 * maxdisp : maxdisp, sg ? -maxdisp : 0, sg);
  SS_CHECK_LAUNCH("ss_concat_volume_backward");
  return SS_OK;
}

extern "C" int ss_disparity_regression_backward(const float* grad_out, float* grad_prob, int B, int D, int H, int W, float dmin, void* stream) {
  SS_REQUIRE(grad_out && grad_prob && B > 0 && D > 0 && H > 0 && W > 0, "ss_disparity_regression_backward: bad argument");
  SS_UNSUPPORTED(B > 65535, "ss_disparity_regression_backward: grid dimension exceeds 65535");
  const size_t HW = (size_t)H * W;
  disparity_regression_backward_kernel<<<dim3((unsigned)ceil_div64(HW, 256), B), 256, 0, (cudaStream_t)stream>>>(grad_out, grad_prob, D, HW, dmin);
  SS_CHECK_LAUNCH("ss_disparity_regression_backward");
  return SS_OK;
}

extern "C" int ss_regression_topk_backward(const float* cost, const float* disp_samples, const float* grad_pred, float* grad_cost,
                                           float* grad_samples, int B, int D, int K, int H, int W, void* stream) {
  SS_REQUIRE(cost && disp_samples && grad_pred && grad_cost && grad_samples, "ss_regression_topk_backward: null pointer");
  SS_REQUIRE(B > 0 && D > 0 && K > 0 && K <= D && H > 0 && W > 0, "ss_regression_topk_backward: need 0 < k <= D");
  SS_UNSUPPORTED(D > 64, "ss_regression_topk_backward: more than 64 samples (%d) unsupported", D);
  SS_UNSUPPORTED(B > 65535, "ss_regression_topk_backward: grid dimension exceeds 65535");
  const size_t HW = (size_t)H * W;
  const dim3 grid((unsigned)ceil_div64(HW, 128), B);
  if (D <= 32)
    regression_topk_backward_kernel<32><<<grid, 128, 0, (cudaStream_t)stream>>>(cost, disp_samples, grad_pred, grad_cost, grad_samples, D, K, HW);
  else
    regression_topk_backward_kernel<64><<<grid, 128, 0, (cudaStream_t)stream>>>(cost, disp_samples, grad_pred, grad_cost, grad_samples, D, K, HW);
  SS_CHECK_LAUNCH("ss_regression_topk_backward");
  return SS_OK;
}

extern "C" int ss_context_upsample_backward(const float* depth_low, const float* up_weights, const float* grad_out, float* grad_depth,
                                            float* grad_weights, int B, int h, int w, void* stream) {
  SS_REQUIRE(depth_low && up_weights && grad_out && grad_depth && grad_weights, "ss_context_upsample_backward: null pointer");
  SS_REQUIRE(B > 0 && h > 0 && w > 0, "ss_context_upsample_backward: non-positive dimension");
  SS_UNSUPPORTED(4 * h > 65535 || B > 65535, "ss_context_upsample_backward: grid dimension exceeds 65535");
  SS_REQUIRE(((reinterpret_cast<uintptr_t>(up_weights) | reinterpret_cast<uintptr_t>(grad_out)) & 15) == 0,
             "ss_context_upsample_backward: up_weights and grad_out must be 16-byte aligned");
  context_upsample_backward_w_kernel<<<dim3(ceil_div(4 * w, 256), 4 * h, B), 256, 0, (cudaStream_t)stream>>>(depth_low, grad_out, grad_weights, h, w);
  SS_CHECK_LAUNCH("ss_context_upsample_backward(weights)");
  context_upsample_backward_d_kernel<<<dim3(ceil_div(w, 128), h, B), 128, 0, (cudaStream_t)stream>>>(up_weights, grad_out, grad_depth, h, w);
  SS_CHECK_LAUNCH("ss_context_upsample_backward(depth)");
  return SS_OK;
}

extern "C" int ss_propagation_backward(const float* grad_out, float* grad_in, int B, int D, int H, int W, void* stream) {
  SS_REQUIRE(grad_out && grad_in && B > 0 && D > 0 && H > 0 && W > 0, "ss_propagation_backward: bad argument");
  SS_UNSUPPORTED((long long)D * H > 65535 || B > 65535, "ss_propagation_backward: grid dimension exceeds 65535");
  propagation_backward_kernel<<<dim3(ceil_div(W, 256), D * H, B), 256, 0, (cudaStream_t)stream>>>(grad_out, grad_in, D, H, W);
  SS_CHECK_LAUNCH("ss_propagation_backward");
  return SS_OK;
}

extern "C" int ss_disparity_variance_backward(const float* prob, const float* disparity, const float* grad_out, float* grad_prob,
                                              float* grad_disparity, int B, int D, int H, int W, float dmin, void* stream) {
  SS_REQUIRE(prob && disparity && grad_out && grad_prob && grad_disparity && B > 0 && D > 0 && H > 0 && W > 0,
             "ss_disparity_variance_backward: bad argument");
  SS_UNSUPPORTED(B > 65535, "ss_disparity_variance_backward: grid dimension exceeds 65535");
  const size_t HW = (size_t)H * W;
  disparity_variance_backward_kernel<<<dim3((unsigned)ceil_div64(HW, 256), B), 256, 0, (cudaStream_t)stream>>>(
      prob, disparity, grad_out, grad_prob, grad_disparity, D, HW, dmin);
  SS_CHECK_LAUNCH("ss_disparity_variance_backward");
  return SS_OK;
}

// grad_y (B,C,H,W) is ACCUMULATED with atomics: it must be zero-filled by the caller.  grad_x_rep / grad_x may both be NULL.
extern "C" int ss_spatial_transformer_grid_backward(const float* y, const float* disp_samples, const float* grad_y_warped,
                                                    const float* grad_x_rep_or_null, float* grad_x_or_null, float* grad_y, float* grad_disp,
                                                    int B, int C, int K, int H, int W, void* stream) {
  SS_REQUIRE(y && disp_samples && grad_y_warped && grad_y && grad_disp, "ss_spatial_transformer_grid_backward: null pointer");
  SS_REQUIRE((grad_x_rep_or_null != nullptr) == (grad_x_or_null != nullptr), "ss_spatial_transformer_grid_backward: grad_x_rep and grad_x go together");
  SS_REQUIRE(B > 0 && C > 0 && K > 0 && H > 1 && W > 1, "ss_spatial_transformer_grid_backward: bad dimension");
  SS_UNSUPPORTED((long long)K * H > 65535 || (long long)B * C > 65535 || B > 65535, "ss_spatial_transformer_grid_backward: grid dimension exceeds 65535");
  stn_backward_kernel<<<dim3(ceil_div(W, 128), K * H, B), 128, 0, (cudaStream_t)stream>>>(y, disp_samples, grad_y_warped, grad_y, grad_disp, C, K, H, W);
  SS_CHECK_LAUNCH("ss_spatial_transformer_grid_backward");
  if (grad_x_or_null) {
    const size_t HW = (size_t)H * W;
    sum_over_k_kernel<<<dim3((unsigned)ceil_div64(HW, 256), B * C), 256, 0, (cudaStream_t)stream>>>(grad_x_rep_or_null, grad_x_or_null, K, HW);
    SS_CHECK_LAUNCH("ss_spatial_transformer_grid_backward(x)");
  }
  return SS_OK;
}
