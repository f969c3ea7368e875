// K5: windowed 3-D multi-head self-attention of the hourglass bottleneck, one CTA per window, fully fused:
// window gather -> qkv Linear -> per-head softmax(q k^T * scale) v -> 1x1x1 output conv (+bias) -> scatter.
// Reference: attention_block.forward, models/submodule_other.py:805-837 (copy at models/submodule_.py:29-61).
// No HBM round trips between the stages: tokens, q/k/v and the head outputs live in shared memory.
#include "common.cuh"

namespace {

constexpr int kC = 128, kHeads = 16, kHd = 8;

struct AttnP {
  const float* x;       // (B,C,D,H,W)
  const float* wqkv_t;  // [C][3C]   = qkv_3d.weight^T
  const float* bqkv;    // [3C]
  const float* wo_t;    // [C][C]    = final1x1.weight^T  (wo_t[c][co])
  const float* bo;      // [C]
  float* out;           // (B,C,D,H,W)
  int B, D, H, W, bd, bh, bw, nd, nh, nw, T;
};

__global__ void __launch_bounds__(256) window_attention3d_kernel(const AttnP p) {
  extern __shared__ __align__(16) float smem[];
  const int T = p.T;
  float* Xt = smem;                       // [C][T]   tokens, later the concatenated head outputs
  float* QKV = smem + kC * T;             // [3][heads][T][hd]
  int wid = blockIdx.x;
  const int wx = wid % p.nw;  wid /= p.nw;
  const int wy = wid % p.nh;  wid /= p.nh;
  const int wz = wid % p.nd;
  const int b = wid / p.nd;
  const size_t cs = (size_t)p.D * p.H * p.W;
  const size_t base = (size_t)b * kC * cs + ((size_t)wz * p.bd * p.H + (size_t)wy * p.bh) * p.W + (size_t)wx * p.bw;
  const int bhw = p.bh * p.bw;

  // token t = (dd, hh, ww) of the window -> offset inside one channel volume
  auto tok_off = [&](int t) -> size_t {
    const int dd = t / bhw, r = t - dd * bhw, hh = r / p.bw, ww = r - hh * p.bw;
    return ((size_t)dd * p.H + hh) * p.W + ww;
  };

  for (int i = threadIdx.x; i < kC * T; i += blockDim.x) {
    const int c = i / T, t = i - c * T;
    Xt[i] = __ldg(p.x + base + (size_t)c * cs + tok_off(t));
  }
  __syncthreads();

  // ---- qkv = X W^T + b : task = (output column, block of 16 tokens) ----
  const int ntb = T / 16;
  for (int id = threadIdx.x; id < 3 * kC * ntb; id += blockDim.x) {
    const int col = id % (3 * kC), tb = id / (3 * kC);
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
    for (int c = 0; c < kC; ++c) {
      const float w = __ldg(p.wqkv_t + (size_t)c * 3 * kC + col);
      const float4* xp = reinterpret_cast<const float4*>(Xt + c * T + 16 * tb);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float4 xv = xp[q4];
        acc[4 * q4 + 0] = fmaf(xv.x, w, acc[4 * q4 + 0]);
        acc[4 * q4 + 1] = fmaf(xv.y, w, acc[4 * q4 + 1]);
        acc[4 * q4 + 2] = fmaf(xv.z, w, acc[4 * q4 + 2]);
        acc[4 * q4 + 3] = fmaf(xv.w, w, acc[4 * q4 + 3]);
      }
    }
    const float bias = __ldg(p.bqkv + col);
    const int which = col / kC, h = (col % kC) / kHd, j = col % kHd;   // qkv channel = which*C + head*hd + j
    float* dst = QKV + (((size_t)which * kHeads + h) * T + 16 * tb) * kHd + j;
#pragma unroll
    for (int i = 0; i < 16; ++i) dst[i * kHd] = acc[i] + bias;
  }
  __syncthreads();

  // ---- attention: task = (head, query token); lanes of a warp share the head -> k/v reads broadcast ----
  const float scale = 0.35355339059327379f;     // hd^-0.5 with hd = 8
  const float* Q = QKV;
  const float* Km = QKV + (size_t)kHeads * T * kHd;
  const float* V = QKV + (size_t)2 * kHeads * T * kHd;
  for (int id = threadIdx.x; id < kHeads * T; id += blockDim.x) {
    const int tq = id % T, h = id / T;
    float q[kHd];
    {
      const float4* qp = reinterpret_cast<const float4*>(Q + ((size_t)h * T + tq) * kHd);
      const float4 a = qp[0], c = qp[1];
      q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = c.x; q[5] = c.y; q[6] = c.z; q[7] = c.w;
    }
    const float4* kp = reinterpret_cast<const float4*>(Km + (size_t)h * T * kHd);
    const float4* vp = reinterpret_cast<const float4*>(V + (size_t)h * T * kHd);
    float m = -INFINITY;
    for (int tk = 0; tk < T; ++tk) {
      const float4 a = kp[2 * tk], c = kp[2 * tk + 1];
      float s = q[0] * a.x;
      s = fmaf(q[1], a.y, s); s = fmaf(q[2], a.z, s); s = fmaf(q[3], a.w, s);
      s = fmaf(q[4], c.x, s); s = fmaf(q[5], c.y, s); s = fmaf(q[6], c.z, s); s = fmaf(q[7], c.w, s);
      m = fmaxf(m, s * scale);
    }
    float l = 0.0f, o[kHd];
#pragma unroll
    for (int j = 0; j < kHd; ++j) o[j] = 0.0f;
    for (int tk = 0; tk < T; ++tk) {
      const float4 a = kp[2 * tk], c = kp[2 * tk + 1];
      float s = q[0] * a.x;
      s = fmaf(q[1], a.y, s); s = fmaf(q[2], a.z, s); s = fmaf(q[3], a.w, s);
      s = fmaf(q[4], c.x, s); s = fmaf(q[5], c.y, s); s = fmaf(q[6], c.z, s); s = fmaf(q[7], c.w, s);
      const float pexp = expf(s * scale - m);
      l += pexp;
      const float4 va = vp[2 * tk], vc = vp[2 * tk + 1];
      o[0] = fmaf(pexp, va.x, o[0]); o[1] = fmaf(pexp, va.y, o[1]); o[2] = fmaf(pexp, va.z, o[2]); o[3] = fmaf(pexp, va.w, o[3]);
      o[4] = fmaf(pexp, vc.x, o[4]); o[5] = fmaf(pexp, vc.y, o[5]); o[6] = fmaf(pexp, vc.z, o[6]); o[7] = fmaf(pexp, vc.w, o[7]);
    }
    const float inv = 1.0f / l;
#pragma unroll
    for (int j = 0; j < kHd; ++j) Xt[(h * kHd + j) * T + tq] = o[j] * inv;    // output channel = head*hd + j
  }
  __syncthreads();

  // ---- final 1x1x1 conv (+bias) and scatter back to NCDHW : task = (cout, block of 16 tokens) ----
  for (int id = threadIdx.x; id < kC * ntb; id += blockDim.x) {
    const int co = id % kC, tb = id / kC;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
    for (int c = 0; c < kC; ++c) {
      const float w = __ldg(p.wo_t + (size_t)c * kC + co);
      const float4* xp = reinterpret_cast<const float4*>(Xt + c * T + 16 * tb);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float4 xv = xp[q4];
        acc[4 * q4 + 0] = fmaf(xv.x, w, acc[4 * q4 + 0]);
        acc[4 * q4 + 1] = fmaf(xv.y, w, acc[4 * q4 + 1]);
        acc[4 * q4 + 2] = fmaf(xv.z, w, acc[4 * q4 + 2]);
        acc[4 * q4 + 3] = fmaf(xv.w, w, acc[4 * q4 + 3]);
      }
    }
    const float bias = __ldg(p.bo + co);
    float* ob = p.out + base + (size_t)co * cs;
#pragma unroll
    for (int i = 0; i < 16; ++i) ob[tok_off(16 * tb + i)] = acc[i] + bias;
  }
}

// GENERIC fallback (any window with T <= 96).  The softmax(q k^T * scale) v core alone, fp32, for the bf16x3 split route: the qkv Linear and the final 1x1x1 conv run as
// fp32-accurate tensor-core GEMMs around it (ss_conv2d_tc over the [hi | lo | hi] K-concat form), so this kernel reads the fp32
// qkv volume (B,3C,D,H,W; channel = which*C + head*hd + j) and writes the head outputs directly in that K-concat form:
// bf16 (B, 3*C/8, D, H, W, 8) -- a head (hd = 8 channels) is exactly one 16-byte channel chunk.
__global__ void __launch_bounds__(256) window_attention_core_f32_generic_kernel(const float* __restrict__ qkv, uint4* __restrict__ out, int D,
                                                                       int H, int W, int bd, int bh, int bw, int nd, int nh, int nw, int T) {
  extern __shared__ __align__(16) float smem[];        // [3][heads][T][hd]
  int wid = blockIdx.x;
  const int wx = wid % nw;  wid /= nw;
  const int wy = wid % nh;  wid /= nh;
  const int wz = wid % nd;
  const int b = wid / nd;
  const size_t cs = (size_t)D * H * W;
  const size_t wbase = ((size_t)wz * bd * H + (size_t)wy * bh) * W + (size_t)wx * bw;
  const int bhw = bh * bw;
  auto tok_off = [&](int t) -> size_t {
    const int dd = t / bhw, r = t - dd * bhw, hh = r / bw, ww = r - hh * bw;
    return ((size_t)dd * H + hh) * W + ww;
  };
  const float* src = qkv + (size_t)b * 3 * kC * cs + wbase;
  for (int i = threadIdx.x; i < 3 * kC * T; i += blockDim.x) {
    const int c = i / T, t = i - c * T;                 // c = which*C + head*hd + j
    smem[((size_t)(c >> 3) * T + t) * kHd + (c & 7)] = __ldg(src + (size_t)c * cs + tok_off(t));
  }
  __syncthreads();
  const float scale = 0.35355339059327379f;             // hd^-0.5 with hd = 8
  const float* Q = smem;
  const float* Km = smem + (size_t)kHeads * T * kHd;
  const float* V = smem + (size_t)2 * kHeads * T * kHd;
  uint4* ob = out + (size_t)b * 3 * kHeads * cs + wbase;
  for (int id = threadIdx.x; id < kHeads * T; id += blockDim.x) {
    const int tq = id % T, h = id / T;
    float q[kHd];
    {
      const float4* qp = reinterpret_cast<const float4*>(Q + ((size_t)h * T + tq) * kHd);
      const float4 a = qp[0], c = qp[1];
      q[0] = a.x * scale; q[1] = a.y * scale; q[2] = a.z * scale; q[3] = a.w * scale;
      q[4] = c.x * scale; q[5] = c.y * scale; q[6] = c.z * scale; q[7] = c.w * scale;
    }
    const float4* kp = reinterpret_cast<const float4*>(Km + (size_t)h * T * kHd);
    const float4* vp = reinterpret_cast<const float4*>(V + (size_t)h * T * kHd);
    // online softmax: one pass over the keys
    float m = -INFINITY, l = 0.0f, o[kHd];
#pragma unroll
    for (int j = 0; j < kHd; ++j) o[j] = 0.0f;
    for (int tk = 0; tk < T; ++tk) {
      const float4 a = kp[2 * tk], c = kp[2 * tk + 1];
      float sc = q[0] * a.x;
      sc = fmaf(q[1], a.y, sc); sc = fmaf(q[2], a.z, sc); sc = fmaf(q[3], a.w, sc);
      sc = fmaf(q[4], c.x, sc); sc = fmaf(q[5], c.y, sc); sc = fmaf(q[6], c.z, sc); sc = fmaf(q[7], c.w, sc);
      const float mn = fmaxf(m, sc);
      const float corr = expf(m - mn), pe = expf(sc - mn);
      m = mn;
      l = fmaf(l, corr, pe);
      const float4 va = vp[2 * tk], vc = vp[2 * tk + 1];
      o[0] = fmaf(o[0], corr, pe * va.x); o[1] = fmaf(o[1], corr, pe * va.y); o[2] = fmaf(o[2], corr, pe * va.z); o[3] = fmaf(o[3], corr, pe * va.w);
      o[4] = fmaf(o[4], corr, pe * vc.x); o[5] = fmaf(o[5], corr, pe * vc.y); o[6] = fmaf(o[6], corr, pe * vc.z); o[7] = fmaf(o[7], corr, pe * vc.w);
    }
    const float inv = 1.0f / l;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x0 = o[2 * j] * inv, x1 = o[2 * j + 1] * inv;
      __nv_bfloat162 hv = __floats2bfloat162_rn(x0, x1);
      hi[j] = *reinterpret_cast<uint32_t*>(&hv);
      __nv_bfloat162 lv = __floats2bfloat162_rn(x0 - __uint_as_float(hi[j] << 16), x1 - __uint_as_float(hi[j] & 0xffff0000u));
      lo[j] = *reinterpret_cast<uint32_t*>(&lv);
    }
    const size_t vo = tok_off(tq);
    const uint4 qh = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    ob[(size_t)h * cs + vo] = qh;
    ob[(size_t)(kHeads + h) * cs + vo] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    ob[(size_t)(2 * kHeads + h) * cs + vo] = qh;
  }
}

// Fast variant for the model's windows (bw == 4, T = bd*bh*bw in {64, 96}): 128-bit window gathers (a window row is 4 contiguous
// floats), and per (head, query) task all T scores are computed first into registers (T independent 8-term dot products), then
// max / exp / weighted sum -- no serial online-softmax chain (the generic kernel below was latency-bound: 92 us for 128 windows).
template <int T>
__global__ void __launch_bounds__(256, 2) window_attention_core_f32_kernel(const float* __restrict__ qkv, uint4* __restrict__ out,
                                                                          float* __restrict__ out_f32, int D, int H, int W, int bd, int bh,
                                                                          int nd, int nh, int nw) {
  extern __shared__ __align__(16) float smem[];        // [3][heads][T][hd]
  constexpr int BW = 4;
  int wid = blockIdx.x;
  const int wx = wid % nw;  wid /= nw;
  const int wy = wid % nh;  wid /= nh;
  const int wz = wid % nd;
  const int b = wid / nd;
  const size_t cs = (size_t)D * H * W;
  const size_t wbase = ((size_t)wz * bd * H + (size_t)wy * bh) * W + (size_t)wx * BW;
  const float* src = qkv + (size_t)b * 3 * kC * cs + wbase;
  // gather: one float4 = the 4 tokens of one (dd, hh) row of the window for one channel
  constexpr int ROWS = T / BW;
#pragma unroll 4
  for (int i = threadIdx.x; i < 3 * kC * ROWS; i += 256) {
    const int c = i / ROWS, r = i - c * ROWS;            // c = which*C + head*hd + j ; r = dd*bh + hh
    const int dd = r / bh, hh = r - dd * bh;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)c * cs + ((size_t)dd * H + hh) * W));
    float* dst = smem + ((size_t)(c >> 3) * T + r * BW) * kHd + (c & 7);
    dst[0] = v.x; dst[kHd] = v.y; dst[2 * kHd] = v.z; dst[3 * kHd] = v.w;
  }
  __syncthreads();
  const float scale = 0.35355339059327379f;             // hd^-0.5 with hd = 8
  const float* Q = smem;
  const float* Km = smem + (size_t)kHeads * T * kHd;
  const float* V = smem + (size_t)2 * kHeads * T * kHd;
  uint4* ob = out + (size_t)b * 3 * kHeads * cs + wbase;
#pragma unroll 1
  for (int id = threadIdx.x; id < kHeads * T; id += 256) {
    const int tq = id % T, h = id / T;                  // T is a multiple of 32: the lanes of a warp share the head -> broadcasts
    float q[kHd];
    {
      const float4* qp = reinterpret_cast<const float4*>(Q + ((size_t)h * T + tq) * kHd);
      const float4 a = qp[0], c = qp[1];
      q[0] = a.x * scale; q[1] = a.y * scale; q[2] = a.z * scale; q[3] = a.w * scale;
      q[4] = c.x * scale; q[5] = c.y * scale; q[6] = c.z * scale; q[7] = c.w * scale;
    }
    const float4* kp = reinterpret_cast<const float4*>(Km + (size_t)h * T * kHd);
    const float4* vp = reinterpret_cast<const float4*>(V + (size_t)h * T * kHd);
    float sc[T];
    float m = -INFINITY;
#pragma unroll
    for (int tk = 0; tk < T; ++tk) {
      const float4 a = kp[2 * tk], c = kp[2 * tk + 1];
      float x = q[0] * a.x;
      x = fmaf(q[1], a.y, x); x = fmaf(q[2], a.z, x); x = fmaf(q[3], a.w, x);
      x = fmaf(q[4], c.x, x); x = fmaf(q[5], c.y, x); x = fmaf(q[6], c.z, x); x = fmaf(q[7], c.w, x);
      sc[tk] = x;
      m = fmaxf(m, x);
    }
    float l = 0.0f, o[kHd];
#pragma unroll
    for (int j = 0; j < kHd; ++j) o[j] = 0.0f;
#pragma unroll
    for (int tk = 0; tk < T; ++tk) {
      const float pe = expf(sc[tk] - m);
      l += pe;
      const float4 va = vp[2 * tk], vc = vp[2 * tk + 1];
      o[0] = fmaf(pe, va.x, o[0]); o[1] = fmaf(pe, va.y, o[1]); o[2] = fmaf(pe, va.z, o[2]); o[3] = fmaf(pe, va.w, o[3]);
      o[4] = fmaf(pe, vc.x, o[4]); o[5] = fmaf(pe, vc.y, o[5]); o[6] = fmaf(pe, vc.z, o[6]); o[7] = fmaf(pe, vc.w, o[7]);
    }
    const float inv = 1.0f / l;
    const int dd = tq / (bh * BW), rr = tq - dd * bh * BW, hh = rr / BW, ww = rr - hh * BW;
    const size_t vo = ((size_t)dd * H + hh) * W + ww;
    if (out_f32) {                      // training / plain fp32 consumers: (B,C,D,H,W), channel = head*hd + j
      float* of = out_f32 + ((size_t)b * kC + h * kHd) * cs + wbase + vo;
#pragma unroll
      for (int j = 0; j < kHd; ++j) of[(size_t)j * cs] = o[j] * inv;
      continue;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x0 = o[2 * j] * inv, x1 = o[2 * j + 1] * inv;
      __nv_bfloat162 hv = __floats2bfloat162_rn(x0, x1);
      hi[j] = *reinterpret_cast<uint32_t*>(&hv);
      __nv_bfloat162 lv = __floats2bfloat162_rn(x0 - __uint_as_float(hi[j] << 16), x1 - __uint_as_float(hi[j] & 0xffff0000u));
      lo[j] = *reinterpret_cast<uint32_t*>(&lv);
    }
    const uint4 qh = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    ob[(size_t)h * cs + vo] = qh;
    ob[(size_t)(kHeads + h) * cs + vo] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    ob[(size_t)(2 * kHeads + h) * cs + vo] = qh;
  }
}

// Backward of the core (training, BASELINE config #5): one CTA per (window, head), thread i = query i (then key/value column i).
//   S = scale q k^T, P = softmax_j(S), O = P v;   dV_j = sum_i P_ij dO_i;  dP_ij = dO_i . v_j;  dS_ij = P_ij (dP_ij - sum_j P_ij dP_ij);
//   dQ_i = scale sum_j dS_ij k_j;  dK_j = scale sum_i dS_ij q_i.        qkv, dqkv (B,3C,D,H,W); dO (B,C,D,H,W).
__global__ void __launch_bounds__(96) window_attention_core_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ d_out,
                                                                       float* __restrict__ d_qkv, int D, int H, int W, int bd, int bh, int bw,
                                                                       int nd, int nh, int nw, int T, int H0, int W0) {
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;                    // [T][8]
  float* Ks = Qs + T * kHd;
  float* Vs = Ks + T * kHd;
  float* Gs = Vs + T * kHd;            // dO
  float* Ps = Gs + T * kHd;            // [T][T+1]
  float* Ss = Ps + T * (T + 1);        // dS [T][T+1]
  float* padf = Ss + T * (T + 1);      // [T] 1 = a zero-padded token (h >= H0 or w >= W0); the masked branch, all 0 otherwise
  const int h = blockIdx.x % kHeads;
  int wid = blockIdx.x / kHeads;
  const int wx = wid % nw;  wid /= nw;
  const int wy = wid % nh;  wid /= nh;
  const int wz = wid % nd;
  const int b = wid / nd;
  const size_t cs = (size_t)D * H * W;
  const size_t wbase = ((size_t)wz * bd * H + (size_t)wy * bh) * W + (size_t)wx * bw;
  const int bhw = bh * bw;
  const int i = threadIdx.x;
  size_t vo = 0;
  if (i < T) {
    const int dd = i / bhw, r = i - dd * bhw, hh = r / bw, ww = r - hh * bw;
    vo = wbase + ((size_t)dd * H + hh) * W + ww;
    padf[i] = (wy * bh + hh >= H0 || wx * bw + ww >= W0) ? 1.0f : 0.0f;
#pragma unroll
    for (int j = 0; j < kHd; ++j) {
      const size_t c = (size_t)h * kHd + j;
      Qs[i * kHd + j] = __ldg(qkv + ((size_t)b * 3 * kC + c) * cs + vo);
      Ks[i * kHd + j] = __ldg(qkv + ((size_t)b * 3 * kC + kC + c) * cs + vo);
      Vs[i * kHd + j] = __ldg(qkv + ((size_t)b * 3 * kC + 2 * kC + c) * cs + vo);
      Gs[i * kHd + j] = __ldg(d_out + ((size_t)b * kC + c) * cs + vo);
    }
  }
  __syncthreads();
  const float scale = 0.35355339059327379f;
  if (i < T) {
    float q[kHd], g[kHd];
#pragma unroll
    for (int j = 0; j < kHd; ++j) { q[j] = Qs[i * kHd + j]; g[j] = Gs[i * kHd + j]; }
    float m = -INFINITY;
    const float pq = padf[i];
    for (int t = 0; t < T; ++t) {
      float x = 0.f;
#pragma unroll
      for (int j = 0; j < kHd; ++j) x = fmaf(q[j], Ks[t * kHd + j], x);
      x *= scale;
      if (padf[t] != pq) x += -1000.0f;                  // attn_mask: an additive constant, so only P changes, not the gradient formulas
      Ps[i * (T + 1) + t] = x;
      m = fmaxf(m, x);
    }
    float l = 0.f;
    for (int t = 0; t < T; ++t) { const float e = expf(Ps[i * (T + 1) + t] - m); Ps[i * (T + 1) + t] = e; l += e; }
    const float inv = 1.0f / l;
    float dsum = 0.f;
    for (int t = 0; t < T; ++t) {
      const float pr = Ps[i * (T + 1) + t] * inv;
      float dp = 0.f;
#pragma unroll
      for (int j = 0; j < kHd; ++j) dp = fmaf(g[j], Vs[t * kHd + j], dp);
      Ps[i * (T + 1) + t] = pr;
      Ss[i * (T + 1) + t] = dp;
      dsum = fmaf(pr, dp, dsum);
    }
    float dq[kHd];
#pragma unroll
    for (int j = 0; j < kHd; ++j) dq[j] = 0.f;
    for (int t = 0; t < T; ++t) {
      const float ds = Ps[i * (T + 1) + t] * (Ss[i * (T + 1) + t] - dsum);
      Ss[i * (T + 1) + t] = ds;
#pragma unroll
      for (int j = 0; j < kHd; ++j) dq[j] = fmaf(ds, Ks[t * kHd + j], dq[j]);
    }
#pragma unroll
    for (int j = 0; j < kHd; ++j) d_qkv[((size_t)b * 3 * kC + (size_t)h * kHd + j) * cs + vo] = dq[j] * scale;
  }
  __syncthreads();
  if (i < T) {                         // column i: dK_i, dV_i
    float dk[kHd], dv[kHd];
#pragma unroll
    for (int j = 0; j < kHd; ++j) { dk[j] = 0.f; dv[j] = 0.f; }
    for (int t = 0; t < T; ++t) {
      const float ds = Ss[t * (T + 1) + i], pr = Ps[t * (T + 1) + i];
#pragma unroll
      for (int j = 0; j < kHd; ++j) { dk[j] = fmaf(ds, Qs[t * kHd + j], dk[j]); dv[j] = fmaf(pr, Gs[t * kHd + j], dv[j]); }
    }
#pragma unroll
    for (int j = 0; j < kHd; ++j) {
      d_qkv[((size_t)b * 3 * kC + kC + (size_t)h * kHd + j) * cs + vo] = dk[j] * scale;
      d_qkv[((size_t)b * 3 * kC + 2 * kC + (size_t)h * kHd + j) * cs + vo] = dv[j];
    }
  }
}

}  // namespace

// The MASKED core: the padded / masked branch of attention_block (submodule_other.py:809-829) when BOTH H and W were zero-padded to the
// window.  qkv is the qkv Linear of the PADDED volume (a padded token carries the bias); tokens with h >= H0 or w >= W0 are padding,
// and a score between a padded and a real token gets -1000 before the softmax, exactly as the reference's attn_mask.  fp32 output
// (B,C,D,H,W), channel = head*hd + j; the caller crops it to (H0, W0) and applies the final 1x1x1 conv.  (One padded axis needs no
// mask -- the reference's `mask[:, -0:, :]` quirk -- and runs on the unmasked kernels, see ops.window_pad.)
namespace {
__global__ void __launch_bounds__(256) window_attention_core_f32_masked_kernel(const float* __restrict__ qkv, float* __restrict__ out, int D, int H,
                                                                                int W, int bd, int bh, int bw, int nd, int nh, int nw, int T,
                                                                                int H0, int W0) {
  extern __shared__ __align__(16) float smem[];        // [3][heads][T][hd] then T pad flags
  int wid = blockIdx.x;
  const int wx = wid % nw;  wid /= nw;
  const int wy = wid % nh;  wid /= nh;
  const int wz = wid % nd;
  const int b = wid / nd;
  const size_t cs = (size_t)D * H * W;
  const size_t wbase = ((size_t)wz * bd * H + (size_t)wy * bh) * W + (size_t)wx * bw;
  const int bhw = bh * bw;
  auto tok_off = [&](int t) -> size_t {
    const int dd = t / bhw, r = t - dd * bhw, hh = r / bw, ww = r - hh * bw;
    return ((size_t)dd * H + hh) * W + ww;
  };
  float* padf = smem + (size_t)3 * kC * T;
  const float* src = qkv + (size_t)b * 3 * kC * cs + wbase;
  for (int i = threadIdx.x; i < 3 * kC * T; i += blockDim.x) {
    const int c = i / T, t = i - c * T;                 // c = which*C + head*hd + j
    smem[((size_t)(c >> 3) * T + t) * kHd + (c & 7)] = __ldg(src + (size_t)c * cs + tok_off(t));
  }
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int r = t % bhw, hh = r / bw, ww = r - hh * bw;
    padf[t] = (wy * bh + hh >= H0 || wx * bw + ww >= W0) ? 1.0f : 0.0f;
  }
  __syncthreads();
  const float scale = 0.35355339059327379f;             // hd^-0.5 with hd = 8
  const float* Q = smem;
  const float* Km = smem + (size_t)kHeads * T * kHd;
  const float* V = smem + (size_t)2 * kHeads * T * kHd;
  float* ob = out + (size_t)b * kC * cs + wbase;
  for (int id = threadIdx.x; id < kHeads * T; id += blockDim.x) {
    const int tq = id % T, h = id / T;
    const float pq = padf[tq];
    float q[kHd];
    {
      const float4* qp = reinterpret_cast<const float4*>(Q + ((size_t)h * T + tq) * kHd);
      const float4 a = qp[0], c = qp[1];
      q[0] = a.x * scale; q[1] = a.y * scale; q[2] = a.z * scale; q[3] = a.w * scale;
      q[4] = c.x * scale; q[5] = c.y * scale; q[6] = c.z * scale; q[7] = c.w * scale;
    }
    const float4* kp = reinterpret_cast<const float4*>(Km + (size_t)h * T * kHd);
    const float4* vp = reinterpret_cast<const float4*>(V + (size_t)h * T * kHd);
    float m = -INFINITY, l = 0.0f, o[kHd];
#pragma unroll
    for (int j = 0; j < kHd; ++j) o[j] = 0.0f;
    for (int tk = 0; tk < T; ++tk) {
      const float4 a = kp[2 * tk], c = kp[2 * tk + 1];
      float sc = q[0] * a.x;
      sc = fmaf(q[1], a.y, sc); sc = fmaf(q[2], a.z, sc); sc = fmaf(q[3], a.w, sc);
      sc = fmaf(q[4], c.x, sc); sc = fmaf(q[5], c.y, sc); sc = fmaf(q[6], c.z, sc); sc = fmaf(q[7], c.w, sc);
      if (padf[tk] != pq) sc += -1000.0f;                // attn + attn_mask (submodule_other.py:826-829)
      const float mn = fmaxf(m, sc);
      const float corr = expf(m - mn), pe = expf(sc - mn);
      m = mn;
      l = fmaf(l, corr, pe);
      const float4 va = vp[2 * tk], vc = vp[2 * tk + 1];
      o[0] = fmaf(o[0], corr, pe * va.x); o[1] = fmaf(o[1], corr, pe * va.y); o[2] = fmaf(o[2], corr, pe * va.z); o[3] = fmaf(o[3], corr, pe * va.w);
      o[4] = fmaf(o[4], corr, pe * vc.x); o[5] = fmaf(o[5], corr, pe * vc.y); o[6] = fmaf(o[6], corr, pe * vc.z); o[7] = fmaf(o[7], corr, pe * vc.w);
    }
    const float inv = 1.0f / l;
    float* op = ob + (size_t)h * kHd * cs + tok_off(tq);
#pragma unroll
    for (int j = 0; j < kHd; ++j) op[(size_t)j * cs] = o[j] * inv;
  }
}
}  // namespace

extern "C" int ss_window_attention_core_f32_masked(const float* qkv, float* out_f32, int B, int C, int D, int H, int W, int bd, int bh, int bw,
                                                   int num_heads, int H0, int W0, void* stream) {
  SS_REQUIRE(qkv && out_f32, "ss_window_attention_core_f32_masked: null pointer");
  SS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && bd > 0 && bh > 0 && bw > 0, "ss_window_attention_core_f32_masked: non-positive dimension");
  SS_REQUIRE(H0 > 0 && H0 <= H && W0 > 0 && W0 <= W, "ss_window_attention_core_f32_masked: the valid extent (H0, W0) must lie inside (H, W)");
  SS_UNSUPPORTED(C != kC || num_heads != kHeads, "ss_window_attention_core_f32_masked: only C=128 with 16 heads is supported");
  SS_UNSUPPORTED(D % bd || H % bh || W % bw, "ss_window_attention_core_f32_masked: the PADDED D,H,W must be multiples of the window");
  const int T = bd * bh * bw;
  SS_UNSUPPORTED(T > 128, "ss_window_attention_core_f32_masked: window of %d tokens unsupported (<= 128)", T);
  const size_t smem = ((size_t)3 * kC * T + T) * sizeof(float);
  const long long nwin = (long long)B * (D / bd) * (H / bh) * (W / bw);
  SS_UNSUPPORTED(nwin > 0x7fffffffLL, "ss_window_attention_core_f32_masked: too many windows");
  SS_CUDA(ss_allow_smem(window_attention_core_f32_masked_kernel, smem));
  window_attention_core_f32_masked_kernel<<<(unsigned)nwin, 256, smem, (cudaStream_t)stream>>>(qkv, out_f32, D, H, W, bd, bh, bw, D / bd, H / bh,
                                                                                                W / bw, T, H0, W0);
  SS_CHECK_LAUNCH("ss_window_attention_core_f32_masked");
  return SS_OK;
}

// fp32 output variant of the core (training forward) and its backward.
extern "C" int ss_window_attention_core_f32_out(const float* qkv, float* out_f32, int B, int C, int D, int H, int W, int bd, int bh, int bw,
                                                int num_heads, void* stream) {
  SS_REQUIRE(qkv && out_f32, "ss_window_attention_core_f32_out: null pointer");
  SS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && bd > 0 && bh > 0 && bw > 0, "ss_window_attention_core_f32_out: non-positive dimension");
  SS_UNSUPPORTED(C != kC || num_heads != kHeads, "ss_window_attention_core_f32_out: only C=128 with 16 heads is supported");
  SS_UNSUPPORTED(D % bd || H % bh || W % bw, "ss_window_attention_core_f32_out: D,H,W must be multiples of the window");
  const int T = bd * bh * bw;
  SS_UNSUPPORTED(!(bw == 4 && W % 4 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (T == 64 || T == 96)),
                 "ss_window_attention_core_f32_out: windows of 64 / 96 tokens with bw = 4 only");
  const size_t smem = (size_t)3 * kC * T * sizeof(float);
  const long long nwin = (long long)B * (D / bd) * (H / bh) * (W / bw);
  if (T == 64) {
    SS_CUDA(ss_allow_smem(window_attention_core_f32_kernel<64>, smem));
    window_attention_core_f32_kernel<64><<<(unsigned)nwin, 256, smem, (cudaStream_t)stream>>>(qkv, nullptr, out_f32, D, H, W, bd, bh, D / bd, H / bh, W / bw);
  } else {
    SS_CUDA(ss_allow_smem(window_attention_core_f32_kernel<96>, smem));
    window_attention_core_f32_kernel<96><<<(unsigned)nwin, 256, smem, (cudaStream_t)stream>>>(qkv, nullptr, out_f32, D, H, W, bd, bh, D / bd, H / bh, W / bw);
  }
  SS_CHECK_LAUNCH("ss_window_attention_core_f32_out");
  return SS_OK;
}

// Backward of the masked core (and, with H0 = H and W0 = W, of the unmasked one).
extern "C" int ss_window_attention_core_backward_masked(const float* qkv, const float* grad_out, float* grad_qkv, int B, int C, int D, int H,
                                                        int W, int bd, int bh, int bw, int num_heads, int H0, int W0, void* stream) {
  SS_REQUIRE(qkv && grad_out && grad_qkv, "ss_window_attention_core_backward: null pointer");
  SS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && bd > 0 && bh > 0 && bw > 0, "ss_window_attention_core_backward: non-positive dimension");
  SS_REQUIRE(H0 > 0 && H0 <= H && W0 > 0 && W0 <= W, "ss_window_attention_core_backward: the valid extent (H0, W0) must lie inside (H, W)");
  SS_UNSUPPORTED(C != kC || num_heads != kHeads, "ss_window_attention_core_backward: only C=128 with 16 heads is supported");
  SS_UNSUPPORTED(D % bd || H % bh || W % bw, "ss_window_attention_core_backward: D,H,W must be multiples of the window");
  const int T = bd * bh * bw;
  SS_UNSUPPORTED(T > 96, "ss_window_attention_core_backward: window of %d tokens unsupported (<= 96)", T);
  const size_t smem = ((size_t)4 * T * kHd + (size_t)2 * T * (T + 1) + T) * sizeof(float);
  const long long nblk = (long long)B * (D / bd) * (H / bh) * (W / bw) * kHeads;
  SS_UNSUPPORTED(nblk > 0x7fffffffLL, "ss_window_attention_core_backward: too many windows");
  SS_CUDA(ss_allow_smem(window_attention_core_bwd_kernel, smem));
  window_attention_core_bwd_kernel<<<(unsigned)nblk, 96, smem, (cudaStream_t)stream>>>(qkv, grad_out, grad_qkv, D, H, W, bd, bh, bw, D / bd, H / bh,
                                                                                      W / bw, T, H0, W0);
  SS_CHECK_LAUNCH("ss_window_attention_core_backward");
  return SS_OK;
}

extern "C" int ss_window_attention_core_backward(const float* qkv, const float* grad_out, float* grad_qkv, int B, int C, int D, int H, int W,
                                                 int bd, int bh, int bw, int num_heads, void* stream) {
  return ss_window_attention_core_backward_masked(qkv, grad_out, grad_qkv, B, C, D, H, W, bd, bh, bw, num_heads, H, W, stream);
}

extern "C" int ss_window_attention_core_f32(const float* qkv, void* out_tri, int B, int C, int D, int H, int W, int bd, int bh, int bw,
                                            int num_heads, void* stream) {
  SS_REQUIRE(qkv && out_tri, "ss_window_attention_core_f32: null pointer");
  SS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && bd > 0 && bh > 0 && bw > 0, "ss_window_attention_core_f32: non-positive dimension");
  SS_REQUIRE((reinterpret_cast<uintptr_t>(out_tri) & 15) == 0, "ss_window_attention_core_f32: output must be 16-byte aligned");
  SS_UNSUPPORTED(C != kC || num_heads != kHeads, "ss_window_attention_core_f32: only C=128 with 16 heads is supported (got C=%d, heads=%d)", C, num_heads);
  SS_UNSUPPORTED(D % bd || H % bh || W % bw, "ss_window_attention_core_f32: D,H,W (%d,%d,%d) must be multiples of the window (%d,%d,%d)", D, H, W, bd, bh, bw);
  const int T = bd * bh * bw;
  SS_UNSUPPORTED(T > 96, "ss_window_attention_core_f32: window of %d tokens unsupported (<= 96)", T);
  const size_t smem = (size_t)3 * kC * T * sizeof(float);
  const long long nwin = (long long)B * (D / bd) * (H / bh) * (W / bw);
  SS_UNSUPPORTED(nwin > 0x7fffffffLL, "ss_window_attention_core_f32: too many windows");
  const bool fast = bw == 4 && W % 4 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (T == 64 || T == 96);
  if (fast && T == 64) {
    SS_CUDA(ss_allow_smem(window_attention_core_f32_kernel<64>, smem));
    window_attention_core_f32_kernel<64><<<(unsigned)nwin, 256, smem, (cudaStream_t)stream>>>(qkv, reinterpret_cast<uint4*>(out_tri), nullptr, D, H, W,
                                                                                              bd, bh, D / bd, H / bh, W / bw);
  } else if (fast) {
    SS_CUDA(ss_allow_smem(window_attention_core_f32_kernel<96>, smem));
    window_attention_core_f32_kernel<96><<<(unsigned)nwin, 256, smem, (cudaStream_t)stream>>>(qkv, reinterpret_cast<uint4*>(out_tri), nullptr, D, H, W,
                                                                                              bd, bh, D / bd, H / bh, W / bw);
  } else {
    SS_CUDA(ss_allow_smem(window_attention_core_f32_generic_kernel, smem));
    window_attention_core_f32_generic_kernel<<<(unsigned)nwin, 256, smem, (cudaStream_t)stream>>>(qkv, reinterpret_cast<uint4*>(out_tri), D, H,
                                                                                                  W, bd, bh, bw, D / bd, H / bh, W / bw, T);
  }
  SS_CHECK_LAUNCH("ss_window_attention_core_f32");
  return SS_OK;
}

extern "C" int ss_window_attention3d(const float* x, const float* wqkv_t, const float* bqkv, const float* wo_t, const float* bo,
                                     float* out, int B, int C, int D, int H, int W, int bd, int bh, int bw, int num_heads,
                                     void* stream) {
  SS_REQUIRE(x && wqkv_t && bqkv && wo_t && bo && out, "ss_window_attention3d: null pointer");
  SS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && bd > 0 && bh > 0 && bw > 0, "ss_window_attention3d: non-positive dimension");
  SS_UNSUPPORTED(C != kC || num_heads != kHeads, "ss_window_attention3d: only C=128 with 16 heads is supported (got C=%d, heads=%d)", C, num_heads);
  // The reference pads H/W to the window and masks (submodule_other.py:809-829).  The host wrapper (ops.window_pad) pads one axis
  // around this kernel -- the reference masks nothing in that case -- and routes two through ss_window_attention_core_f32_masked;
  // the kernel itself takes divisible volumes.
  SS_UNSUPPORTED(D % bd || H % bh || W % bw, "ss_window_attention3d: D,H,W (%d,%d,%d) must be multiples of the window (%d,%d,%d)", D, H, W, bd, bh, bw);
  const int T = bd * bh * bw;
  SS_UNSUPPORTED(T % 16 || T > 96, "ss_window_attention3d: window of %d tokens unsupported (multiple of 16, <= 96)", T);
  AttnP p;
  p.x = x; p.wqkv_t = wqkv_t; p.bqkv = bqkv; p.wo_t = wo_t; p.bo = bo; p.out = out;
  p.B = B; p.D = D; p.H = H; p.W = W; p.bd = bd; p.bh = bh; p.bw = bw;
  p.nd = D / bd; p.nh = H / bh; p.nw = W / bw; p.T = T;
  const size_t smem = (size_t)(kC * T + 3 * kC * T) * sizeof(float);
  const long long nwin = (long long)B * p.nd * p.nh * p.nw;
  SS_UNSUPPORTED(nwin > 0x7fffffffLL, "ss_window_attention3d: too many windows");
  SS_CUDA(ss_allow_smem(window_attention3d_kernel, smem));
  window_attention3d_kernel<<<(unsigned)nwin, 256, smem, (cudaStream_t)stream>>>(p);
  SS_CHECK_LAUNCH("ss_window_attention3d");
  return SS_OK;
}
