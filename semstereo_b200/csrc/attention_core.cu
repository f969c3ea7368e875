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

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core variant (T = 64 or 96 tokens): per (16-query block, head) one warp computes S = Q K^T with mma.sync m16n8k8
// (K = head dim 8: the GEMM is far too thin for a tcgen05 tile -- M >= 64, and K would be zero-padded to 16), keeps the
// 16 x T score fragment in registers, does the softmax there (row max / sum across the 4 lanes of a quad) and feeds the
// probabilities, rounded to bf16, straight back as the A operand of P V (m16n8k16; V fragments via ldmatrix.trans).
// q, k, v of all 16 heads are staged once per window as bf16 [chunk][token][8] (16-byte rows: conflict-free fragment loads).
// The fp32 kernel above spent its time in FMAs (2*T*16 per query and head); here the FP32 pipe only does the exponentials.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_m16n8k8_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void mma_m16n8k16_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int T>
__global__ void __launch_bounds__(256) window_attn_core_mma_kernel(const CoreP p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];          // [48 chunks][T tokens][8 bf16]
  uint4* sm4 = reinterpret_cast<uint4*>(smem_raw);
  const uint32_t* sm32 = reinterpret_cast<const uint32_t*>(smem_raw);
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
  for (int i = threadIdx.x; i < 3 * kHeads * T; i += blockDim.x) {
    const int t = i % T, c = i / T;
    sm4[i] = __ldg(qkv_b + (size_t)c * S + tok_off(t));
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const float c_exp = 0.35355339059327379f * 1.4426950408889634f;       // hd^-0.5 * log2(e)
  uint32_t* out32 = reinterpret_cast<uint32_t*>(p.out);
  constexpr int QB = T / 16, NT = T / 8;
  for (int task = warp; task < QB * kHeads; task += 8) {
    const int qb = task % QB, h = task / QB;
    const int r0 = qb * 16 + g;
    const uint32_t a0 = sm32[((h * T) + r0) * 4 + t4], a1 = sm32[((h * T) + r0 + 8) * 4 + t4];
    float sc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.0f;
      mma_m16n8k8_bf16(sc[j], a0, a1, sm32[(((kHeads + h) * T) + j * 8 + g) * 4 + t4]);
    }
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) { m0 = fmaxf(m0, fmaxf(sc[j][0], sc[j][1])); m1 = fmaxf(m1, fmaxf(sc[j][2], sc[j][3])); }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.0f, l1 = 0.0f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      sc[j][0] = exp2f((sc[j][0] - m0) * c_exp); sc[j][1] = exp2f((sc[j][1] - m0) * c_exp);
      sc[j][2] = exp2f((sc[j][2] - m1) * c_exp); sc[j][3] = exp2f((sc[j][3] - m1) * c_exp);
      l0 += sc[j][0] + sc[j][1];
      l1 += sc[j][2] + sc[j][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    const uint32_t vbase = tc::smem_u32(smem_raw) + (uint32_t)((2 * kHeads + h) * T) * 16u;
#pragma unroll
    for (int kk = 0; kk < T / 16; ++kk) {
      uint32_t b0, b1;
      // lanes 0-7: rows (keys) kk*16 + 0..7, lanes 8-15: keys kk*16 + 8..15 (the other lanes' addresses are ignored)
      const uint32_t addr = vbase + (uint32_t)(kk * 16 + (lane & 15)) * 16u;
      asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(addr));
      mma_m16n8k16_bf16(o, tc::pack_bf16x2(sc[2 * kk][0], sc[2 * kk][1]), tc::pack_bf16x2(sc[2 * kk][2], sc[2 * kk][3]),
                        tc::pack_bf16x2(sc[2 * kk + 1][0], sc[2 * kk + 1][1]), tc::pack_bf16x2(sc[2 * kk + 1][2], sc[2 * kk + 1][3]), b0, b1);
    }
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    // output channel = head*8 + j (submodule_other.py:833): chunk h, this thread's two channels 2*t4, 2*t4+1
    out32[(((size_t)b * kHeads + h) * S + base + tok_off(r0)) * 4 + t4] = tc::pack_bf16x2(o[0] * i0, o[1] * i0);
    out32[(((size_t)b * kHeads + h) * S + base + tok_off(r0 + 8)) * 4 + t4] = tc::pack_bf16x2(o[2] * i1, o[3] * i1);
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
  const long long nwin = (long long)B * p.nd * p.nh * p.nw;
  SS_UNSUPPORTED(nwin > 0x7fffffffLL, "ss_window_attention_core_blocked: too many windows");
  if (T == 64 || T == 96) {            // the model's windows (4,4,4) and (6,4,4): tensor-core kernel
    const size_t smem_mma = (size_t)3 * kHeads * T * 16;
    if (T == 64) {
      SS_CUDA(ss_allow_smem(window_attn_core_mma_kernel<64>, smem_mma));
      window_attn_core_mma_kernel<64><<<(unsigned)nwin, 256, smem_mma, (cudaStream_t)stream>>>(p);
    } else {
      SS_CUDA(ss_allow_smem(window_attn_core_mma_kernel<96>, smem_mma));
      window_attn_core_mma_kernel<96><<<(unsigned)nwin, 256, smem_mma, (cudaStream_t)stream>>>(p);
    }
    SS_CHECK_LAUNCH("ss_window_attention_core_blocked");
    return SS_OK;
  }
  const size_t smem = (size_t)2 * kHeads * T * kHd * sizeof(float);
  SS_CUDA(ss_allow_smem(window_attn_core_kernel, smem));
  window_attn_core_kernel<<<(unsigned)nwin, 256, smem, (cudaStream_t)stream>>>(p);
  SS_CHECK_LAUNCH("ss_window_attention_core_blocked");
  return SS_OK;
}
