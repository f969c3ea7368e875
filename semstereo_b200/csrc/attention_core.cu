// K5, tensor-core mode: the softmax(q k^T * scale) v core of attention_block (models/submodule_other.py:815-833) on the
// bf16 blocked layout.  The qkv Linear and the final 1x1x1 conv run as tensor-core 1x1 layers (ss_conv3d_tc kind 1); what is
// left per (window, head) is a T x T attention with head dim 8.  In the blocked layout a head IS a channel chunk:
// q of head h = chunk h, k = chunk 16+h, v = chunk 32+h of the (B,48,D,H,W,8) qkv tensor, and the output chunk h of (B,16,D,H,W,8).
// One CTA per window; K and V of all 16 heads are staged once in shared memory as fp32; one thread per (head, query token)
// runs a single-pass online softmax (lanes of a warp share the head, so K/V reads are broadcasts).
#include "tc_common.cuh"

namespace {

constexpr int kHeads = 16, kHd = 8;

struct CoreP {
  const uint4* qkv;   // blocked bf16 (B, 48, D, H, W, 8)
  uint4* out;         // blocked bf16 (B, 16, D, H, W, 8)
  int B, D, H, W, bd, bh, bw, nd, nh, nw, T;
};

__device__ __forceinline__ void unpack8(const uint4 q, float (&f)[8]) {
  const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(u[i] << 16);
    f[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
  }
}

__global__ void __launch_bounds__(256) window_attn_core_kernel(const CoreP p) {
  extern __shared__ __align__(16) float smem[];   // K [heads][T][8], V [heads][T][8]
  const int T = p.T;
  float* Ks = smem;
  float* Vs = smem + kHeads * T * kHd;
  int wid = blockIdx.x;
  const int wx = wid % p.nw;  wid /= p.nw;
  const int wy = wid % p.nh;  wid /= p.nh;
  const int wz = wid % p.nd;
  const int b = wid / p.nd;
  const size_t S = (size_t)p.D * p.H * p.W;
  const size_t base = ((size_t)wz * p.bd * p.H + (size_t)wy * p.bh) * p.W + (size_t)wx * p.bw;
  const int bhw = p.bh * p.bw;
  auto tok_off = [&](int t) -> size_t {
    const int dd = t / bhw, r = t - dd * bhw, hh = r / p.bw, ww = r - hh * p.bw;
    return ((size_t)dd * p.H + hh) * p.W + ww;
  };
  const uint4* qkv_b = p.qkv + (size_t)b * 3 * kHeads * S + base;

  for (int i = threadIdx.x; i < 2 * kHeads * T; i += blockDim.x) {
    const int t = i % T, hc = i / T;                 // hc in [0, 32): K heads then V heads
    float f[8];
    unpack8(__ldg(qkv_b + (size_t)(kHeads + hc) * S + tok_off(t)), f);
    float4* dst = reinterpret_cast<float4*>(smem + ((size_t)hc * T + t) * kHd);
    dst[0] = make_float4(f[0], f[1], f[2], f[3]);
    dst[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  __syncthreads();

  const float scale_log2e = 0.35355339059327379f * 1.4426950408889634f;   // hd^-0.5 * log2(e)
  for (int id = threadIdx.x; id < kHeads * T; id += blockDim.x) {
    const int tq = id % T, h = id / T;
    const size_t voff = tok_off(tq);
    float q[8];
    unpack8(__ldg(qkv_b + (size_t)h * S + voff), q);
#pragma unroll
    for (int j = 0; j < 8; ++j) q[j] *= scale_log2e;
    const float4* kp = reinterpret_cast<const float4*>(Ks + (size_t)h * T * kHd);
    const float4* vp = reinterpret_cast<const float4*>(Vs + (size_t)h * T * kHd);
    float m = -INFINITY, l = 0.0f, o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.0f;
    for (int tk = 0; tk < T; ++tk) {
      const float4 a = kp[2 * tk], c = kp[2 * tk + 1];
      float s = q[0] * a.x;
      s = fmaf(q[1], a.y, s); s = fmaf(q[2], a.z, s); s = fmaf(q[3], a.w, s);
      s = fmaf(q[4], c.x, s); s = fmaf(q[5], c.y, s); s = fmaf(q[6], c.z, s); s = fmaf(q[7], c.w, s);
      if (s > m) {                                    // new running maximum: rescale what has been accumulated
        const float corr = exp2f(m - s);
        l *= corr;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] *= corr;
        m = s;
      }
      const float pe = exp2f(s - m);
      l += pe;
      const float4 va = vp[2 * tk], vc = vp[2 * tk + 1];
      o[0] = fmaf(pe, va.x, o[0]); o[1] = fmaf(pe, va.y, o[1]); o[2] = fmaf(pe, va.z, o[2]); o[3] = fmaf(pe, va.w, o[3]);
      o[4] = fmaf(pe, vc.x, o[4]); o[5] = fmaf(pe, vc.y, o[5]); o[6] = fmaf(pe, vc.z, o[6]); o[7] = fmaf(pe, vc.w, o[7]);
    }
    const float inv = 1.0f / l;
    uint4 r;
    r.x = tc::pack_bf16x2(o[0] * inv, o[1] * inv); r.y = tc::pack_bf16x2(o[2] * inv, o[3] * inv);
    r.z = tc::pack_bf16x2(o[4] * inv, o[5] * inv); r.w = tc::pack_bf16x2(o[6] * inv, o[7] * inv);
    p.out[((size_t)b * kHeads + h) * S + base + voff] = r;     // output channel = head*8 + j  (submodule_other.py:833)
  }
}

}  // namespace

// qkv: blocked bf16 (B, 3*C/8, D, H, W, 8) with channel = which*C + head*8 + j  ->  out: blocked bf16 (B, C/8, D, H, W, 8)
extern "C" int ss_window_attention_core_blocked(const void* qkv_blocked, void* out_blocked, int B, int C, int D, int H, int W, int bd,
                                                int bh, int bw, int num_heads, void* stream) {
  SS_REQUIRE(qkv_blocked && out_blocked, "ss_window_attention_core_blocked: null pointer");
  SS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && bd > 0 && bh > 0 && bw > 0, "ss_window_attention_core_blocked: non-positive dimension");
  SS_UNSUPPORTED(C != kHeads * kHd || num_heads != kHeads,
                 "ss_window_attention_core_blocked: only C=128 with 16 heads is supported (got C=%d, heads=%d)", C, num_heads);
  SS_UNSUPPORTED(D % bd || H % bh || W % bw,
                 "ss_window_attention_core_blocked: D,H,W (%d,%d,%d) must be multiples of the window (%d,%d,%d)", D, H, W, bd, bh, bw);
  const int T = bd * bh * bw;
  SS_UNSUPPORTED(T > 128, "ss_window_attention_core_blocked: window of %d tokens unsupported (<= 128)", T);
  CoreP p;
  p.qkv = reinterpret_cast<const uint4*>(qkv_blocked);
  p.out = reinterpret_cast<uint4*>(out_blocked);
  p.B = B; p.D = D; p.H = H; p.W = W; p.bd = bd; p.bh = bh; p.bw = bw;
  p.nd = D / bd; p.nh = H / bh; p.nw = W / bw; p.T = T;
  const size_t smem = (size_t)2 * kHeads * T * kHd * sizeof(float);
  const long long nwin = (long long)B * p.nd * p.nh * p.nw;
  SS_UNSUPPORTED(nwin > 0x7fffffffLL, "ss_window_attention_core_blocked: too many windows");
  SS_CUDA(ss_allow_smem(window_attn_core_kernel, smem));
  window_attn_core_kernel<<<(unsigned)nwin, 256, smem, (cudaStream_t)stream>>>(p);
  SS_CHECK_LAUNCH("ss_window_attention_core_blocked");
  return SS_OK;
}
