// K6, K7, K8, K10, K12 and the standalone warps/propagations: single-pass, HBM-bound kernels that keep
// the whole disparity column of a pixel in registers.
//   att_stats            : trilinear x2 -> softmax -> mean -> variance -> sigmoid gate   (SemStereo.py:279-287)
//   sample_strength      : Propagation x2 -> SpatialTransformer_grid -> corr -> softmax   (SemStereo.py:288-293)
//   topk_select          : Propagation_prob -> mix -> softmax -> top-k -> gathers          (SemStereo.py:295-310)
//   regression_topk      : models/submodule.py:434-442
//   disparity_regression / disparity_variance : models/submodule.py:164-170, 257-263
//   propagation / propagation_prob / spatial_transformer_grid : models/submodule.py:265-307, 361-377
#include <type_traits>

#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// linear-resize source index, PyTorch align_corners=False: src = max(0,(dst+.5)*scale-.5)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void lin_src(int dst, float scale, int n_in, int& i0, int& i1, float& l0, float& l1) {
  float src = fmaxf(((float)dst + 0.5f) * scale - 0.5f, 0.0f);
  i0 = min((int)src, n_in - 1);
  i1 = min(i0 + 1, n_in - 1);
  l1 = src - (float)i0;
  l0 = 1.0f - l1;
}

// ---------------------------------------------------------------------------------------------
// K6
// ---------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(128) att_stats_kernel(const float* __restrict__ cost, const float* __restrict__ beta,
                                                        const float* __restrict__ gamma, float* __restrict__ att_up,
                                                        float* __restrict__ mu_out, float* __restrict__ gate_out,
                                                        int B, int D8, int H8, int W8, float dmin) {
  const int H4 = 2 * H8, W4 = 2 * W8, nb = 2 * D8;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= W4) return;
  int y0, y1, x0, x1;
  float hy0, hy1, wx0, wx1;
  lin_src(y, 0.5f, H8, y0, y1, hy0, hy1);
  lin_src(x, 0.5f, W8, x0, x1, wx0, wx1);
  const float* cb = cost + (size_t)b * D8 * H8 * W8;
  // v[d] for d < D8, and v[D8] = v[D8 - 1]: the x2 depth upsample (align_corners=False) of bin k reads the fixed pair
  // (k/2 - 1, k/2) with weights (1/4, 3/4) for even k and (k/2, k/2 + 1) with (3/4, 1/4) for odd k -- exactly lin_src()'s values --
  // except bin 0 (source clamped to 0: weight 1 on slice 0) and the last bin, whose upper neighbour is clamped to D8 - 1: the
  // duplicated last slice reproduces that clamp, so no run-time index selection is needed (round 2: it was ~half the instructions).
  const unsigned HW8 = (unsigned)H8 * (unsigned)W8, o00 = y0 * W8 + x0, o01 = y0 * W8 + x1, o10 = y1 * W8 + x0, o11 = y1 * W8 + x1;
  float v[NB / 2 + 1];
#pragma unroll
  for (int d = 0; d <= NB / 2; ++d) {
    const int ds = min(d, D8 - 1);
    if (d <= D8) {
      const float* s = cb + (size_t)ds * HW8;
      const float a = __ldg(s + o00), bb = __ldg(s + o01), c = __ldg(s + o10), dd = __ldg(s + o11);
      v[d] = hy0 * (wx0 * a + wx1 * bb) + hy1 * (wx0 * c + wx1 * dd);
    } else v[d] = 0.0f;
  }
  float u[NB];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    if (k < nb) {
      if (k == 0) u[k] = 1.0f * v[0] + 0.0f * v[D8 > 1 ? 1 : 0];
      else if (k & 1) u[k] = 0.75f * v[k / 2] + 0.25f * v[k / 2 + 1];
      else u[k] = 0.25f * v[k / 2 - 1] + 0.75f * v[k / 2];
      m = fmaxf(m, u[k]);
    } else u[k] = -INFINITY;
  }
  const size_t pix = (size_t)y * W4 + x;
  float* ao = att_up + (size_t)b * nb * H4 * W4 + pix;
  float s = 0.0f;
  float e[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    if (k < nb) {
      ao[(size_t)k * H4 * W4] = u[k];
      e[k] = expf(u[k] - m);
      s += e[k];
    } else e[k] = 0.0f;
  }
  float mu = 0.0f;
#pragma unroll
  for (int k = 0; k < NB; ++k)
    if (k < nb) { e[k] = e[k] / s; mu += e[k] * (dmin + (float)k); }
  float var = 0.0f;
#pragma unroll
  for (int k = 0; k < NB; ++k)
    if (k < nb) { float t = (dmin + (float)k) - mu; var += e[k] * (t * t); }
  mu_out[(size_t)b * H4 * W4 + pix] = mu;
  gate_out[(size_t)b * H4 * W4 + pix] = sigmoidf_(__ldg(beta) + __ldg(gamma) * var);
}

// ---------------------------------------------------------------------------------------------
// K7: block = 128 pixels of one image row, one thread per pixel and all 5 hypotheses.  Per 32-channel chunk the left row tile
// and the right row are staged in shared memory (128-bit loads) with a +-PAD column window.  The bilinear weights factor into
// an x and a y term and the y term is the same for the whole row (iy depends on y only), so the two right rows floor(iy),
// floor(iy)+1 are blended while staging -- with the reference's fp32 coordinate round trip the second weight is exactly 0 on
// ~75 % of the rows and that row is then not even loaded.  A hypothesis then costs 2 shared loads per channel instead of 4,
// and the left value is read once for all 5.  Hypotheses whose corners leave the window (|disparity| > PAD-2) read global memory.
// ---------------------------------------------------------------------------------------------
// Round 2 (ncu r02_three: 160 M instructions for 42 M FP32 + 21 M shared loads -- index arithmetic and predicates of the staging
// loops dominated, at 23 % occupancy): 256 threads per CTA, the two halves of the CTA take the two halves of every staged
// 32-channel chunk for the same 128 pixels (partial sums combined through shared memory at the end).  The left features are NOT
// staged: a thread is the only reader of its pixel's left value, so it loads its 16 channels straight into registers (coalesced
// 512-byte rows) before the right rows are staged, which hides their latency.  Staging is per warp and per channel (warp w takes
// channels w, w+8, ...; a lane owns the same one or two 128-bit columns of every row, so bounds and offsets are computed once).
// The x weights are applied AFTER the channel sum (sum_c l*r0 and sum_c l*r1 are accumulated separately: 2 FMAs per channel and
// hypothesis instead of 3 operations).
constexpr int SS_TX = 128, SS_NT = 256, SS_PAD = 24, SS_WW = SS_TX + 2 * SS_PAD, SS_CK = 32, SS_Q = SS_WW / 4;

template <bool VEC4>
__global__ void __launch_bounds__(SS_NT, 4) sample_strength_kernel(const float* __restrict__ fl, const float* __restrict__ fr,
                                                                   const float* __restrict__ mu, const float* __restrict__ gate,
                                                                   float* __restrict__ strength, int B, int C, int H, int W) {
  __shared__ __align__(16) float Rc[SS_CK][SS_WW];
  __shared__ float Part[5][SS_TX];
  const int tid = threadIdx.x, tx = tid & (SS_TX - 1), half = tid >> 7, warp = tid >> 5, lane = tid & 31;
  const int x0 = blockIdx.x * SS_TX, x = x0 + tx, y = blockIdx.y, b = blockIdx.z;
  const size_t HW = (size_t)H * W;
  const float iy = warp_coord((float)y, (float)(H - 1));
  const float fy0 = floorf(iy), fy1 = fy0 + 1.0f;
  const bool vy0 = fy0 >= 0.0f && fy0 <= (float)(H - 1), vy1 = fy1 >= 0.0f && fy1 <= (float)(H - 1);
  const float wy0 = vy0 ? fy1 - iy : 0.0f, wy1 = vy1 ? iy - fy0 : 0.0f;      // uniform over the row
  const int y0 = vy0 ? (int)fy0 : 0, y1 = vy1 ? (int)fy1 : 0;
  const bool active = x < W;
  // hypothesis s: the disparity / gate of the pixel's s-th propagation neighbour; only the window offset stays live across the
  // channel loop (the rest is recomputed after it: registers decide this kernel's occupancy)
  auto hypothesis = [&](int s, float& gs, float& ix, bool& inw, int& j, float& fx0) {
    float d = 0.0f;
    gs = 0.0f;
    if (active) {
      const int ty = min(max(y + kPropDy[s], 0), H - 1), txx = min(max(x + kPropDx[s], 0), W - 1);
      d = __ldg(mu + (size_t)b * HW + (size_t)ty * W + txx);
      gs = __ldg(gate + (size_t)b * HW + (size_t)ty * W + txx);
    }
    ix = warp_coord((float)x - d, (float)(W - 1));
    fx0 = floorf(ix);
    const float jf = fx0 - (float)(x0 - SS_PAD);
    inw = jf >= 0.0f && jf <= (float)(SS_WW - 2);            // both corners inside the staged window (zeros outside the image)
    j = inw ? (int)jf : 0;                                    // out-of-window hypotheses are redone from global memory below
  };
  int j0[5];
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    float gs, ix, fx0;
    bool inw;
    hypothesis(s, gs, ix, inw, j0[s], fx0);
  }
  // staging geometry of this lane: 128-bit columns lane and lane + 32 of the 44-column window; columns outside the image (or past
  // the window) load a clamped address and are replaced by zeros, so the loop body is branch-free
  const int xa = x0 - SS_PAD + 4 * lane, xb = xa + 128;
  const bool oka = xa >= 0 && xa < W, okb = lane + 32 < SS_Q && xb < W;      // xb >= 0 always
  const int cola = oka ? xa : 0, offb = okb ? xb - cola : 0;
  const float* r0p = fr + (size_t)b * C * HW + (size_t)y0 * W + (size_t)warp * HW + cola;      // rows y0 / y1 of channel `warp`
  const float* r1p = fr + (size_t)b * C * HW + (size_t)y1 * W + (size_t)warp * HW + cola;
  const float* lp = fl + (size_t)b * C * HW + (size_t)y * W + (size_t)(half * (SS_CK / 2)) * HW + (active ? x : 0);
  const float* lb = fl + (size_t)b * C * HW + (size_t)y * W;
  const float* rb = fr + (size_t)b * C * HW;
  const bool two_rows = wy1 != 0.0f;                   // CTA-uniform
  float a0[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, a1[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int c0 = 0; c0 < C; c0 += SS_CK) {
    float l[SS_CK / 2];                                // this thread's left values of its half of the chunk
    if (c0 + SS_CK <= C) {
#pragma unroll
      for (int c = 0; c < SS_CK / 2; ++c) l[c] = __ldg(lp + (size_t)c * HW);
    } else {
#pragma unroll
      for (int c = 0; c < SS_CK / 2; ++c) l[c] = c0 + half * (SS_CK / 2) + c < C ? __ldg(lp + (size_t)c * HW) : 0.0f;
    }
    lp += (size_t)SS_CK * HW;
    if (VEC4) {        // rows are 16-byte aligned: x0 - PAD is a multiple of 4
#pragma unroll
      for (int k = 0; k < SS_CK / 8; ++k) {            // channel c0 + warp + 8k (warp-uniform)
        float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
        if (c0 + warp + 8 * k < C) {
          va = __ldg(reinterpret_cast<const float4*>(r0p));
          vb = __ldg(reinterpret_cast<const float4*>(r0p + offb));
          va.x *= wy0; va.y *= wy0; va.z *= wy0; va.w *= wy0;
          vb.x *= wy0; vb.y *= wy0; vb.z *= wy0; vb.w *= wy0;
          if (two_rows) {
            const float4 ta = __ldg(reinterpret_cast<const float4*>(r1p)), tb = __ldg(reinterpret_cast<const float4*>(r1p + offb));
            va.x = fmaf(ta.x, wy1, va.x); va.y = fmaf(ta.y, wy1, va.y); va.z = fmaf(ta.z, wy1, va.z); va.w = fmaf(ta.w, wy1, va.w);
            vb.x = fmaf(tb.x, wy1, vb.x); vb.y = fmaf(tb.y, wy1, vb.y); vb.z = fmaf(tb.z, wy1, vb.z); vb.w = fmaf(tb.w, wy1, vb.w);
          }
          if (!oka) va = make_float4(0.f, 0.f, 0.f, 0.f);
          if (!okb) vb = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        *reinterpret_cast<float4*>(&Rc[warp + 8 * k][4 * lane]) = va;
        if (lane + 32 < SS_Q) *reinterpret_cast<float4*>(&Rc[warp + 8 * k][4 * lane + 128]) = vb;
        r0p += 8 * HW;
        r1p += 8 * HW;
      }
    } else {
      for (int i = tid; i < SS_CK * SS_WW; i += SS_NT) {
        const int c = i / SS_WW, j = i - c * SS_WW;
        const int xx = x0 - SS_PAD + j;
        float v = 0.0f;
        if (c0 + c < C && xx >= 0 && xx < W) {
          const float* rp = rb + (size_t)(c0 + c) * HW + xx;
          if (wy0 != 0.0f) v = __ldg(rp + (size_t)y0 * W) * wy0;
          if (wy1 != 0.0f) v = fmaf(__ldg(rp + (size_t)y1 * W), wy1, v);
        }
        Rc[c][j] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < SS_CK / 2; ++c) {              // channels past C hold zeros on both sides
      const float* rc = Rc[half * (SS_CK / 2) + c];
#pragma unroll
      for (int s = 0; s < 5; ++s) {
        a0[s] = fmaf(l[c], rc[j0[s]], a0[s]);
        a1[s] = fmaf(l[c], rc[j0[s] + 1], a1[s]);
      }
    }
    __syncthreads();
  }
  // combine the two channel halves: the upper half hands its partial sums to the lower half, which finishes the pixel
  float acc[5];
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    float gs, ix, fx0;
    bool inw;
    int j;
    hypothesis(s, gs, ix, inw, j, fx0);
    acc[s] = inw ? fmaf(a1[s], ix - fx0, a0[s] * (fx0 + 1.0f - ix)) : 0.0f;
  }
  if (half == 1) {
#pragma unroll
    for (int s = 0; s < 5; ++s) Part[s][tx] = acc[s];
  }
  __syncthreads();
  if (half == 1) return;
#pragma unroll
  for (int s = 0; s < 5; ++s) acc[s] += Part[s][tx];
  if (active) {
    float logit[5], m = -INFINITY;
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      float gs, ix, fx0;
      bool inw;
      int j;
      hypothesis(s, gs, ix, inw, j, fx0);
      if (!inw) {                                   // rare: wild disparity, corners outside the staged window
        const Bilin q = make_bilin(ix, iy, H, W);
        float a = 0.0f;
        for (int c = 0; c < C; ++c) a += __ldg(lb + (size_t)c * HW + x) * bilin_fetch(rb + (size_t)c * HW, q);
        acc[s] = a;
      }
      logit[s] = (acc[s] / (float)C) * gs;
      m = fmaxf(m, logit[s]);
    }
    float e[5], sum = 0.0f;
#pragma unroll
    for (int s = 0; s < 5; ++s) { e[s] = expf(logit[s] - m); sum += e[s]; }
#pragma unroll
    for (int s = 0; s < 5; ++s) strength[((size_t)b * 5 + s) * HW + (size_t)y * W + x] = e[s] / sum;
  }
}

// ---------------------------------------------------------------------------------------------
// K8
// ---------------------------------------------------------------------------------------------
// Round 2 (ncu r02_three: ~6000 instructions per pixel, most of them 64-bit address arithmetic and per-bin `k < nb` branches):
// block-uniform base pointers + 32-bit element offsets, a FULL instantiation for nb == NB (the model's 32 bins), one running
// output offset, and the renormalised expectation as (sum e*d) / (sum e) -- one division instead of one per kept bin.
template <int NB, bool FULL>
__global__ void __launch_bounds__(128) topk_select_kernel(const float* __restrict__ att, const float* __restrict__ strength,
                                                          long long* __restrict__ ind_k, float* __restrict__ att_topk,
                                                          float* __restrict__ disp_topk, float* __restrict__ pred_att,
                                                          float* __restrict__ prob_out, int B, int nb_, int K, int H, int W,
                                                          float disp_offset) {
  using mask_t = typename std::conditional<(NB <= 32), unsigned, unsigned long long>::type;
  const int nb = FULL ? NB : nb_;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  const unsigned HW = (unsigned)H * (unsigned)W, pix = (unsigned)y * W + x;      // nb * HW < 2^32 (checked by the launcher)
  const float* sb = strength + (size_t)b * 5 * HW;
  const float* ab = att + (size_t)b * nb * HW;
  const unsigned HW4 = HW * 4u;                         // bytes of one bin plane (< 2^32, checked by the launcher)
  const char* ap[5];
  float st[5];
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int ty = min(max(y + kPropDy[s], 0), H - 1), tx = min(max(x + kPropDx[s], 0), W - 1);
    ap[s] = reinterpret_cast<const char*>(ab + ((unsigned)ty * W + tx));
    st[s] = __ldg(sb + ((unsigned)s * HW + pix));
  }
  float mix[NB], p[NB];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    if (FULL || k < nb) {
      // address = tap pointer + k * (HW * 4): one 32 x 32 -> 64-bit multiply-add per load
      float acc = __fmul_rn(__ldg(reinterpret_cast<const float*>(ap[0] + (unsigned long long)HW4 * (unsigned)k)), st[0]);
#pragma unroll
      for (int s = 1; s < 5; ++s)           // products then a sequential sum, as torch.sum(dim=1)
        acc = __fadd_rn(acc, __fmul_rn(__ldg(reinterpret_cast<const float*>(ap[s] + (unsigned long long)HW4 * (unsigned)k)), st[s]));
      mix[k] = acc;
      m = fmaxf(m, acc);
    } else mix[k] = -INFINITY;
  }
  float sum = 0.0f;
#pragma unroll
  for (int k = 0; k < NB; ++k) { p[k] = (FULL || k < nb) ? expf(mix[k] - m) : 0.0f; sum += p[k]; }
#pragma unroll
  for (int k = 0; k < NB; ++k) p[k] = (FULL || k < nb) ? p[k] / sum : -1.0f;
  if (prob_out) {
    float* po = prob_out + (size_t)b * nb * HW;
#pragma unroll
    for (int k = 0; k < NB; ++k)
      if (FULL || k < nb) po[(unsigned)k * HW + pix] = p[k];
  }
  // Keep the K largest probabilities, ties broken toward the lower index (total order: p descending, index ascending).
  // The model keeps 24 of 32 bins, so it is cheaper to DROP the nb-K last elements of that order one by one (minimum p,
  // among equals the highest index: `<=` lets a later index win) than to rank all bins against each other (a rank-by-count
  // variant with NB*NB independent comparisons was measured slower, 0.27 vs 0.22 ms at batch 8: round 2).
  mask_t keep = nb >= (int)(8 * sizeof(mask_t)) ? ~(mask_t)0 : (((mask_t)1 << nb) - 1);
#pragma unroll 1
  for (int r = 0; r < nb - K; ++r) {
    float mn = INFINITY;
    int mi = 0;
#pragma unroll
    for (int k = 0; k < NB; ++k)
      if (((keep >> k) & 1) && p[k] <= mn) { mn = p[k]; mi = k; }
    keep &= ~((mask_t)1 << mi);
  }
  float m2 = -INFINITY;
#pragma unroll
  for (int k = 0; k < NB; ++k)
    if ((keep >> k) & 1) m2 = fmaxf(m2, mix[k]);
  float num = 0.0f, den = 0.0f;
  unsigned o = pix;                                    // + j * HW for the j-th kept bin
  long long* ik = ind_k ? ind_k + (size_t)b * K * HW : nullptr;
  float* at = att_topk + (size_t)b * K * HW;
  float* dt = disp_topk + (size_t)b * K * HW;
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    if ((keep >> k) & 1) {
      const float d = (float)k - disp_offset;
      if (ik) ik[o] = k;
      at[o] = p[k];
      dt[o] = d;
      const float e = expf(mix[k] - m2);
      den += e;
      num = fmaf(e, d, num);
      o += HW;
    }
  }
  pred_att[(size_t)b * HW + pix] = num / den;
}

// ---------------------------------------------------------------------------------------------
// K10
// ---------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(128) regression_topk_kernel(const float* __restrict__ cost, const float* __restrict__ samples,
                                                              float* __restrict__ out, int B, int D, int K, size_t HW) {
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (pix >= HW) return;
  float c[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) c[k] = (k < D) ? __ldg(cost + ((size_t)b * D + k) * HW + pix) : -INFINITY;
  unsigned long long taken = 0ull;
  float top = 0.0f, sum = 0.0f, acc = 0.0f;
  for (int r = 0; r < K; ++r) {
    float best = -INFINITY;
    int bi = 0;
    bool found = false;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      bool free_ = k < D && !((taken >> k) & 1ull);
      if (free_ && (!found || c[k] > best)) { best = c[k]; bi = k; found = true; }   // strict > keeps the lower index on ties
    }
    taken |= 1ull << bi;
    if (r == 0) top = best;
    const float e = expf(best - top);
    sum += e;
    acc += e * __ldg(samples + ((size_t)b * D + bi) * HW + pix);
  }
  out[(size_t)b * HW + pix] = acc / sum;
}

// ---------------------------------------------------------------------------------------------
// K12 + standalone propagation / warp
// ---------------------------------------------------------------------------------------------
__global__ void regress_var_kernel(const float* __restrict__ p, const float* __restrict__ mu_in, float* __restrict__ out,
                                   int D, size_t HW, float dmin, int variance) {
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (pix >= HW) return;
  const float mu = variance ? __ldg(mu_in + (size_t)b * HW + pix) : 0.0f;
  float acc = 0.0f;
  for (int k = 0; k < D; ++k) {
    const float v = __ldg(p + ((size_t)b * D + k) * HW + pix);
    const float d = dmin + (float)k;
    acc += variance ? v * ((d - mu) * (d - mu)) : v * d;
  }
  out[(size_t)b * HW + pix] = acc;
}

// in (B,D,H,W) -> out (B,5,D,H,W); D = 1 for Propagation
__global__ void propagation_kernel(const float* __restrict__ in, float* __restrict__ out, int D, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const int y = blockIdx.y % H, d = blockIdx.y / H, b = blockIdx.z;
  const size_t HW = (size_t)H * W;
  const float* ib = in + ((size_t)b * D + d) * HW;
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int ty = min(max(y + kPropDy[s], 0), H - 1), tx = min(max(x + kPropDx[s], 0), W - 1);
    out[(((size_t)b * 5 + s) * D + d) * HW + (size_t)y * W + x] = __ldg(ib + (size_t)ty * W + tx);
  }
}

// y_warped[b,c,k,y,x] = bilinear(src[b,c], x - disp[b,k,y,x], y);  x_rep[b,c,k,y,x] = xin[b,c,y,x]
__global__ void __launch_bounds__(128) stn_kernel(const float* __restrict__ xin, const float* __restrict__ src,
                                                  const float* __restrict__ disp, float* __restrict__ y_warped,
                                                  float* __restrict__ x_rep, int C, int K, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const int y = blockIdx.y % H, k = blockIdx.y / H, b = blockIdx.z;
  const size_t HW = (size_t)H * W;
  const float d = __ldg(disp + ((size_t)b * K + k) * HW + (size_t)y * W + x);
  const Bilin q = make_bilin(warp_coord((float)x - d, (float)(W - 1)), warp_coord((float)y, (float)(H - 1)), H, W);
  for (int c = 0; c < C; ++c) {
    const size_t o = (((size_t)b * C + c) * K + k) * HW + (size_t)y * W + x;
    y_warped[o] = bilin_fetch(src + ((size_t)b * C + c) * HW, q);
    if (x_rep) x_rep[o] = __ldg(xin + ((size_t)b * C + c) * HW + (size_t)y * W + x);
  }
}

}  // namespace

#define SS_GRID_LIMIT(cond, name) SS_UNSUPPORTED(!(cond), "%s: grid dimension exceeds 65535", name)

extern "C" int ss_att_stats(const float* cost_att, const float* beta, const float* gamma, float* att_up, float* mu, float* gate,
                            int B, int D8, int H8, int W8, float dmin, void* stream) {
  SS_REQUIRE(cost_att && beta && gamma && att_up && mu && gate, "ss_att_stats: null pointer");
  SS_REQUIRE(B > 0 && D8 > 0 && H8 > 0 && W8 > 0, "ss_att_stats: non-positive dimension");
  SS_UNSUPPORTED(2 * D8 > 64, "ss_att_stats: more than 64 disparity bins (%d) unsupported", 2 * D8);
  SS_GRID_LIMIT(2 * H8 <= 65535 && B <= 65535, "ss_att_stats");
  dim3 grid(ceil_div(2 * W8, 128), 2 * H8, B);
  if (2 * D8 <= 32)
    att_stats_kernel<32><<<grid, 128, 0, (cudaStream_t)stream>>>(cost_att, beta, gamma, att_up, mu, gate, B, D8, H8, W8, dmin);
  else
    att_stats_kernel<64><<<grid, 128, 0, (cudaStream_t)stream>>>(cost_att, beta, gamma, att_up, mu, gate, B, D8, H8, W8, dmin);
  SS_CHECK_LAUNCH("ss_att_stats");
  return SS_OK;
}

extern "C" int ss_sample_strength(const float* feat_l, const float* feat_r, const float* mu, const float* gate, float* strength,
                                  int B, int C, int H, int W, void* stream) {
  SS_REQUIRE(feat_l && feat_r && mu && gate && strength, "ss_sample_strength: null pointer");
  SS_REQUIRE(B > 0 && C > 0 && H > 1 && W > 1, "ss_sample_strength: bad dimension");
  SS_GRID_LIMIT(H <= 65535 && B <= 65535, "ss_sample_strength");
  dim3 grid(ceil_div(W, SS_TX), H, B);
  if ((W & 3) == 0 && (reinterpret_cast<uintptr_t>(feat_r) & 15) == 0)
    sample_strength_kernel<true><<<grid, SS_NT, 0, (cudaStream_t)stream>>>(feat_l, feat_r, mu, gate, strength, B, C, H, W);
  else
    sample_strength_kernel<false><<<grid, SS_NT, 0, (cudaStream_t)stream>>>(feat_l, feat_r, mu, gate, strength, B, C, H, W);
  SS_CHECK_LAUNCH("ss_sample_strength");
  return SS_OK;
}

extern "C" int ss_topk_select(const float* att_up, const float* strength, long long* ind_k, float* att_topk, float* disp_topk,
                              float* pred_att, float* prob_or_null, int B, int nbins, int K, int H, int W, float disp_offset,
                              void* stream) {
  SS_REQUIRE(att_up && strength && att_topk && disp_topk && pred_att, "ss_topk_select: null pointer");
  SS_REQUIRE(B > 0 && nbins > 0 && K > 0 && K <= nbins && H > 0 && W > 0, "ss_topk_select: need 0 < K <= nbins");
  SS_UNSUPPORTED(nbins > 64, "ss_topk_select: more than 64 disparity bins (%d) unsupported", nbins);
  SS_GRID_LIMIT(H <= 65535 && B <= 65535, "ss_topk_select");
  SS_GRID_LIMIT((unsigned long long)nbins * H * W < (1ull << 32) && (unsigned long long)5 * H * W < (1ull << 32), "ss_topk_select");
  dim3 grid(ceil_div(W, 128), H, B);
#define SS_TOPK(NB, FULL)                                                                                                     \
  topk_select_kernel<NB, FULL><<<grid, 128, 0, (cudaStream_t)stream>>>(att_up, strength, ind_k, att_topk, disp_topk, pred_att, \
                                                                        prob_or_null, B, nbins, K, H, W, disp_offset)
  if (nbins == 32) SS_TOPK(32, true);
  else if (nbins < 32) SS_TOPK(32, false);
  else if (nbins == 64) SS_TOPK(64, true);
  else SS_TOPK(64, false);
#undef SS_TOPK
  SS_CHECK_LAUNCH("ss_topk_select");
  return SS_OK;
}

extern "C" int ss_regression_topk(const float* cost, const float* disp_samples, float* pred, int B, int D, int K, int H, int W,
                                  void* stream) {
  SS_REQUIRE(cost && disp_samples && pred, "ss_regression_topk: null pointer");
  SS_REQUIRE(B > 0 && D > 0 && K > 0 && K <= D && H > 0 && W > 0, "ss_regression_topk: need 0 < k <= D");
  SS_UNSUPPORTED(D > 64, "ss_regression_topk: more than 64 samples (%d) unsupported", D);
  SS_GRID_LIMIT(B <= 65535, "ss_regression_topk");
  const size_t HW = (size_t)H * W;
  dim3 grid((unsigned)ceil_div64(HW, 128), B);
  if (D <= 32) regression_topk_kernel<32><<<grid, 128, 0, (cudaStream_t)stream>>>(cost, disp_samples, pred, B, D, K, HW);
  else         regression_topk_kernel<64><<<grid, 128, 0, (cudaStream_t)stream>>>(cost, disp_samples, pred, B, D, K, HW);
  SS_CHECK_LAUNCH("ss_regression_topk");
  return SS_OK;
}

extern "C" int ss_disparity_regression(const float* prob, float* out, int B, int D, int H, int W, float dmin, void* stream) {
  SS_REQUIRE(prob && out && B > 0 && D > 0 && H > 0 && W > 0, "ss_disparity_regression: bad argument");
  SS_GRID_LIMIT(B <= 65535, "ss_disparity_regression");
  const size_t HW = (size_t)H * W;
  regress_var_kernel<<<dim3((unsigned)ceil_div64(HW, 256), B), 256, 0, (cudaStream_t)stream>>>(prob, nullptr, out, D, HW, dmin, 0);
  SS_CHECK_LAUNCH("ss_disparity_regression");
  return SS_OK;
}

extern "C" int ss_disparity_variance(const float* prob, const float* disparity, float* out, int B, int D, int H, int W, float dmin,
                                     void* stream) {
  SS_REQUIRE(prob && disparity && out && B > 0 && D > 0 && H > 0 && W > 0, "ss_disparity_variance: bad argument");
  SS_GRID_LIMIT(B <= 65535, "ss_disparity_variance");
  const size_t HW = (size_t)H * W;
  regress_var_kernel<<<dim3((unsigned)ceil_div64(HW, 256), B), 256, 0, (cudaStream_t)stream>>>(prob, disparity, out, D, HW, dmin, 1);
  SS_CHECK_LAUNCH("ss_disparity_variance");
  return SS_OK;
}

extern "C" int ss_propagation(const float* in, float* out, int B, int D, int H, int W, void* stream) {
  SS_REQUIRE(in && out && B > 0 && D > 0 && H > 0 && W > 0, "ss_propagation: bad argument");
  SS_GRID_LIMIT((int64_t)D * H <= 65535 && B <= 65535, "ss_propagation");
  propagation_kernel<<<dim3(ceil_div(W, 128), D * H, B), 128, 0, (cudaStream_t)stream>>>(in, out, D, H, W);
  SS_CHECK_LAUNCH("ss_propagation");
  return SS_OK;
}

extern "C" int ss_spatial_transformer_grid(const float* x, const float* y, const float* disp_samples, float* y_warped,
                                           float* x_rep_or_null, int B, int C, int K, int H, int W, void* stream) {
  SS_REQUIRE(y && disp_samples && y_warped && (x || !x_rep_or_null), "ss_spatial_transformer_grid: null pointer");
  SS_REQUIRE(B > 0 && C > 0 && K > 0 && H > 1 && W > 1, "ss_spatial_transformer_grid: bad dimension");
  SS_GRID_LIMIT((int64_t)K * H <= 65535 && B <= 65535, "ss_spatial_transformer_grid");
  stn_kernel<<<dim3(ceil_div(W, 128), K * H, B), 128, 0, (cudaStream_t)stream>>>(x, y, disp_samples, y_warped, x_rep_or_null,
                                                                                C, K, H, W);
  SS_CHECK_LAUNCH("ss_spatial_transformer_grid");
  return SS_OK;
}
