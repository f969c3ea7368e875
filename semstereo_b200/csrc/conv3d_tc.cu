// K4, tensor-core mode: every 3-D convolution of the hourglass stack as an implicit GEMM on the 5th-generation tensor
// cores: TMA -> shared memory -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld epilogue.
// Reference layers: convbn_3d / BasicConv(is_3d) (models/submodule_other.py:845-848, models/submodule.py:89-116) and
// nn.ConvTranspose3d(k3,s2,p1,op1) as instantiated at models/SemStereo.py:106-182, 228-236.
//
// Data layout ("blocked channels", bf16): activations are [B][C/8][D][H][W][8]: the 8 channels of a chunk are the 16 bytes
// one UMMA core-matrix row needs, and voxels that are neighbours along W are neighbours in memory.  A (TH+2)x(TW+2) halo
// tile of one depth slice therefore lands in shared memory (one TMA box, zero-filled outside the volume = the conv padding)
// as [C/8][TH+2][TW+2][16 B], which IS the canonical K-major no-swizzle UMMA layout for every in-plane tap: a tap (kh,kw)
// is just a different start address (SBO = (TW+2)*16 B between 8-voxel row groups, LBO = chunk pitch).  An input slice is
// fetched ONCE per output tile column and reused by all its taps and by up to 3 output slices (depth-sliding ring).
// GEMM tile: M = 128 voxels (16 h x 8 w of one depth slice), N = Cout tile, K = taps * Cin.
//
// Four kernels share the skeleton (persistent CTAs, 256 threads, 1 CTA/SM; warp 0 = TMA slice producer, warp 3 = weight
// producer (resident: all taps once; streamed: per-tap ring), warp 1 = MMA issuer (warp-uniform control flow, one elected
// lane issues), warp 2 = TMEM allocator, warps 4-7 = epilogue; accumulators double-buffered in TMEM, s1f: a 512-column ring):
//   s1 : Conv3d k3 s1 p1 (TAPS=27), Conv3d k1 (TAPS=1), Conv2d 3x3 on a depth-1 volume (TAPS=9).
//   s1f: Conv3d k3 s1 p1 for Cout = 32 / 64 with the three depth taps folded into the GEMM N (see the kernel's comment).
//   s2 : Conv3d k3 s2 p1 on a phase-split ("s2d") input [B][8 phases (d,h,w parity)][C/8][D/2][H/2][W/2][8]: input index
//        2o-1+k is (phase 1, o-1), (phase 0, o), (phase 1, o) for k = 0,1,2, so every tap is a dense half-resolution tile.
//   t2 : ConvTranspose3d k3 s2 p1 op1 as 8 sub-pixel output phases (1/2/4/8 taps each, no zero insertion); the hourglass
//        skip connection (redir conv output, stored phase-split) is added in the epilogue before the ReLU.
// Epilogue: y = acc*scale[co] + shift[co] (+ residual) -> ReLU -> * gate[b,co,h,w] -> bf16 blocked / phase-split, or fp32 NCDHW.
#include <type_traits>

#include "tc_common.cuh"

namespace {

constexpr int TH = 16, TW = 8, HH = TH + 2, WW = TW + 2;
constexpr uint32_t TILE_B = HH * WW * 16;       // bytes of one channel chunk of a halo tile (2880)

struct TcP {
  const __nv_bfloat16* w;         // [n_tiles][TAPS][CIN/8][N][8]
  const float* scale;             // [Cout] or null
  const float* shift;             // [Cout] or null
  const float* gate;              // sigmoid(gate logits) as fp32 blocked (B,Cout/8,OH,OW,8), or null
  const __nv_bfloat16* residual;  // t2 only: phase-split blocked [B][8][Cout/8][D][H][W][8], or null
  const __nv_bfloat16* skip_w;    // t2 only: when set, `residual` is the INPUT of the 1x1 skip conv and this is its weight
                                  // [N/8][N][8] (BN scale folded in): the skip conv runs as one more GEMM tap on the staged tile
  const float* acc_in;            // fp32 partial sums in the out_mode-3 layout [B][Cout/4][D][H][W][4], added to the accumulator BEFORE
  const float* acc_in2;           // scale/shift (bf16x3 split route: the products of the other operand halves), or null
  void* out;
  long long split_off;            // != 0: bf16 outputs are written as a hi/lo pair, lo = bf16(v - hi) at out + split_off (uint4 units)
  int out_mode;                   // 0: bf16 blocked, 1: fp32 NCDHW, 2: bf16 phase-split blocked, 3: fp32 [B][Cout/4][D][H][W][4]
  int cout_valid;                 // channels actually stored (Cout = n_tiles*N may be zero-padded), also the channel count of out
  int B, D, H, W;                 // tile space: output dims (s1, s2) / input dims (t2)
  int relu;
  int n_tiles, HT, WT, DC, n_dc, items;   // items = B*HT*WT*n_dc per n-tile
};

__device__ __forceinline__ void decode_item(const TcP& p, int s, int& b, int& h0, int& w0, int& dlo, int& dhi) {
  const int dc = s % p.n_dc;  s /= p.n_dc;
  const int wt = s % p.WT;    s /= p.WT;
  const int ht = s % p.HT;
  b = s / p.HT;
  h0 = ht * TH; w0 = wt * TW;
  dlo = dc * p.DC; dhi = min(p.D, dlo + p.DC);
}

// Applies the fused epilogue to 32 consecutive output channels (co0 ...) of one voxel and stores them.
// (od,oh,ow) / (OD,OH,OW): output voxel and output dims.  r: the 4 residual chunks of these channels, already loaded (the
// loads are issued BEFORE the wait on the accumulator so their latency hides behind the MMAs), or nullptr.
// EPI (compile time, so that the plain bf16 kernels keep their register budget): bit 0 = the bf16 output may be a hi/lo pair
// (p.split_off), bit 1 = fp32 partial sums may come in (p.acc_in*) and the fp32 [C/4][...][4] output mode 3 exists.
template <int EPI = 0>
__device__ __forceinline__ void epilogue_store32(const TcP& p, float (&v)[32], const float* sc, const float* sh, int co0, int b, int od,
                                                 int oh, int ow, int OD, int OH, int OW, const uint4* r) {
  const float lo = p.relu ? 0.0f : -INFINITY;
  if constexpr ((EPI & 2) != 0) {
    const size_t OSa = (size_t)OD * OH * OW, o0 = ((size_t)b * (p.cout_valid / 4) + co0 / 4) * OSa + ((size_t)od * OH + oh) * OW + ow;
    if (p.out_mode == 3) {            // raw partial sums out: 8 coalesced 128-bit stores (8 lanes = 128 contiguous bytes)
      float4* o = reinterpret_cast<float4*>(p.out) + o0;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (co0 + 4 * q < p.cout_valid) o[(size_t)q * OSa] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      return;
    }
    if (p.acc_in) {                   // partial sums of the other split products, added before the affine
      const float4* a = reinterpret_cast<const float4*>(p.acc_in) + o0;
      const float4* a2 = p.acc_in2 ? reinterpret_cast<const float4*>(p.acc_in2) + o0 : nullptr;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (co0 + 4 * q < p.cout_valid) {
          float4 t = __ldg(a + (size_t)q * OSa);
          if (a2) {
            const float4 u = __ldg(a2 + (size_t)q * OSa);
            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
          }
          v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
        }
    }
  }
  if (r) {
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
      const uint32_t u[4] = {r[c8].x, r[c8].y, r[c8].z, r[c8].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[8 * c8 + 2 * i] = fmaf(v[8 * c8 + 2 * i], sc[8 * c8 + 2 * i], sh[8 * c8 + 2 * i]) + __uint_as_float(u[i] << 16);
        v[8 * c8 + 2 * i + 1] = fmaf(v[8 * c8 + 2 * i + 1], sc[8 * c8 + 2 * i + 1], sh[8 * c8 + 2 * i + 1]) + __uint_as_float(u[i] & 0xffff0000u);
      }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], lo);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(fmaf(v[i], sc[i], sh[i]), lo);
  }
  const size_t OHW = (size_t)OH * OW;
  if (p.gate) {                       // pre-activated gate, fp32 blocked (B,Cout/8,OH,OW,8): two 128-bit loads per chunk
    const float4* gp = reinterpret_cast<const float4*>(p.gate) + (((size_t)b * (p.cout_valid / 8) + co0 / 8) * OHW + (size_t)oh * OW + ow) * 2;
    float4 g[8];
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
      const bool ok = co0 + 8 * c8 < p.cout_valid;
      g[2 * c8] = ok ? __ldg(gp + (size_t)c8 * OHW * 2) : make_float4(1.f, 1.f, 1.f, 1.f);
      g[2 * c8 + 1] = ok ? __ldg(gp + (size_t)c8 * OHW * 2 + 1) : make_float4(1.f, 1.f, 1.f, 1.f);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) { v[4 * q] *= g[q].x; v[4 * q + 1] *= g[q].y; v[4 * q + 2] *= g[q].z; v[4 * q + 3] *= g[q].w; }
  }
  const size_t sp = ((size_t)od * OH + oh) * OW + ow, OS = (size_t)OD * OHW;
  if (p.out_mode == 1) {
    float* o = reinterpret_cast<float*>(p.out) + ((size_t)b * p.cout_valid + co0) * OS + sp;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (co0 + i < p.cout_valid) o[(size_t)i * OS] = v[i];
  } else {
    uint4* o;
    size_t cs = OS;                   // chunk stride in uint4
    if (p.out_mode == 0) o = reinterpret_cast<uint4*>(p.out) + ((size_t)b * (p.cout_valid / 8) + co0 / 8) * OS + sp;
    else {                            // phase-split: [B][8][C/8][OD/2][OH/2][OW/2]
      const int phase = ((od & 1) << 2) | ((oh & 1) << 1) | (ow & 1);
      cs = OS >> 3;
      o = reinterpret_cast<uint4*>(p.out) + (((size_t)b * 8 + phase) * (p.cout_valid / 8) + co0 / 8) * cs +
          ((size_t)(od >> 1) * (OH >> 1) + (oh >> 1)) * (OW >> 1) + (ow >> 1);
    }
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
      if (co0 + 8 * c8 >= p.cout_valid) break;
      uint4 q;
      q.x = tc::pack_bf16x2(v[8 * c8 + 0], v[8 * c8 + 1]);
      q.y = tc::pack_bf16x2(v[8 * c8 + 2], v[8 * c8 + 3]);
      q.z = tc::pack_bf16x2(v[8 * c8 + 4], v[8 * c8 + 5]);
      q.w = tc::pack_bf16x2(v[8 * c8 + 6], v[8 * c8 + 7]);
      o[(size_t)c8 * cs] = q;
      if ((EPI & 1) != 0 && p.split_off) {      // lo half of the bf16x3 split: what bf16 rounding of the value lost
        const uint32_t u[4] = {q.x, q.y, q.z, q.w};
        uint32_t l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          l[i] = tc::pack_bf16x2(v[8 * c8 + 2 * i] - __uint_as_float(u[i] << 16), v[8 * c8 + 2 * i + 1] - __uint_as_float(u[i] & 0xffff0000u));
        o[(size_t)c8 * cs + p.split_off] = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
  }
}

// Common prologue: barriers, TMEM, folded-BN staging.
#define TC_KERNEL_PROLOGUE(NSLOTS, NWSLOTS, RESIDENT) TC_KERNEL_PROLOGUE_N(NSLOTS, NWSLOTS, RESIDENT, 2)
#define TC_KERNEL_PROLOGUE_N(NSLOTS, NWSLOTS, RESIDENT, NACC) TC_KERNEL_PROLOGUE_NF(NSLOTS, NWSLOTS, RESIDENT, NACC, 1)
#define TC_KERNEL_PROLOGUE_NF(NSLOTS, NWSLOTS, RESIDENT, NACC, FULLCNT)                                                    \
  extern __shared__ __align__(1024) uint8_t smem[];                                                                        \
  __shared__ __align__(8) uint64_t a_full[NSLOTS], a_empty[NSLOTS], w_full[NWSLOTS], w_empty[NWSLOTS], acc_full[NACC], acc_empty[NACC]; \
  __shared__ uint32_t tmem_base_s;                                                                                         \
  __shared__ float s_scale[N], s_shift[N];                                                                                 \
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;                                                              \
  const int nt = blockIdx.x % p.n_tiles;                                                                                   \
  const int cta_s = blockIdx.x / p.n_tiles, cta_stride = gridDim.x / p.n_tiles;                                            \
  for (int i = threadIdx.x; i < N; i += blockDim.x) {                                                                      \
    s_scale[i] = (p.scale && nt * N + i < p.cout_valid) ? __ldg(p.scale + nt * N + i) : 1.0f;                              \
    s_shift[i] = (p.shift && nt * N + i < p.cout_valid) ? __ldg(p.shift + nt * N + i) : 0.0f;                              \
  }                                                                                                                        \
  if (threadIdx.x == 0) {                                                                                                  \
    tc::prefetch_tmap(&tmA);                                                                                               \
    for (int i = 0; i < NSLOTS; ++i) { tc::mbar_init(&a_full[i], 1); tc::mbar_init(&a_empty[i], 1); }                      \
    for (int i = 0; i < ((RESIDENT) ? 1 : NWSLOTS); ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }  \
    for (int i = 0; i < NACC; ++i) { tc::mbar_init(&acc_full[i], FULLCNT); tc::mbar_init(&acc_empty[i], 128); }            \
    tc::fence_barrier_init();                                                                                              \
  }                                                                                                                        \
  if (warp == 2) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);                                                                  \
  tc::fence_before_sync();                                                                                                 \
  __syncthreads();                                                                                                         \
  tc::fence_after_sync();                                                                                                  \
  const uint32_t tmem_base = tmem_base_s;

#define TC_KERNEL_EPILOGUE()                                   \
  tc::fence_before_sync();                                     \
  __syncthreads();                                             \
  if (warp == 2) tc::tmem_dealloc(tmem_base, TMEM_COLS);

constexpr uint32_t tmem_cols_for(int n2) { return n2 <= 32 ? 32 : n2 <= 64 ? 64 : n2 <= 128 ? 128 : n2 <= 256 ? 256 : 512; }

// =====================================================================================================================
// s1: Conv3d k3 s1 p1 (TAPS = 27) / Conv3d k1 (TAPS = 1).  NWS == TAPS -> weights resident, else streamed tap ring.
// =====================================================================================================================
template <int CIN, int N, int NS, int NWS, int TAPS, int EPI = 0>
__global__ void __launch_bounds__(256, 1) conv3d_tc_s1_kernel(const __grid_constant__ CUtensorMap tmA, const TcP p) {
  constexpr bool kResident = (NWS == TAPS);
  constexpr bool k3 = (TAPS == 27);            // taps along the depth axis
  constexpr bool kHW3 = (TAPS >= 9);           // 3x3 in-plane taps (TAPS == 9: a 2-D 3x3 conv on a depth-1 volume)
  constexpr uint32_t SLICE = (CIN / 8) * TILE_B;
  constexpr uint32_t TAPB = CIN * N * 2;
  constexpr int KS = CIN / 16;
  constexpr uint32_t LBO_A = TILE_B, SBO_A = WW * 16, LBO_B = N * 16, SBO_B = 128;
  constexpr uint32_t TMEM_COLS = tmem_cols_for(2 * N);
  constexpr uint32_t IDESC = tc::make_idesc_bf16(128, N);
  TC_KERNEL_PROLOGUE(NS, NWS, kResident)
  uint8_t* Abase = smem;
  uint8_t* Wbase = smem + NS * SLICE;
  const int halo = k3 ? 1 : 0;

  if (warp == 0 && lane == 0) {
    // ===== input-slice producer =====
    uint32_t g = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int din0 = max(dlo - halo, 0), din1 = min(dhi - 1 + halo, p.D - 1);
      for (int d_in = din0; d_in <= din1; ++d_in, ++g) {
        const uint32_t slot = g % NS;
        tc::mbar_wait(&a_empty[slot], ((g / NS) & 1) ^ 1);
        tc::mbar_expect_tx(&a_full[slot], SLICE);
        tc::tma_load_4d(Abase + slot * SLICE, &tmA, &a_full[slot], (w0 - 1) * 8, h0 - 1, d_in, b * (CIN / 8));
      }
    }
  } else if (warp == 3 && lane == 0) {
    // ===== weight producer =====
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) + (size_t)nt * TAPS * TAPB;
    if (kResident) {
      if (cta_s < p.items) {
        tc::mbar_expect_tx(&w_full[0], TAPS * TAPB);
        for (int tap = 0; tap < TAPS; ++tap) tc::bulk_load(Wbase + tap * TAPB, wsrc + (size_t)tap * TAPB, TAPB, &w_full[0]);
      }
    } else {
      uint32_t wc = 0;
      for (int s = cta_s; s < p.items; s += cta_stride) {
        int b, h0, w0, dlo, dhi;
        decode_item(p, s, b, h0, w0, dlo, dhi);
        for (int d_out = dlo; d_out < dhi; ++d_out)
          for (int kd = (k3 ? 0 : 1); kd < (k3 ? 3 : 2); ++kd) {
            const int d_in = d_out + kd - 1;
            if (d_in < 0 || d_in >= p.D) continue;
            for (int t9 = (kHW3 ? 0 : 4); t9 < (kHW3 ? 9 : 5); ++t9, ++wc) {
              const uint32_t slot = wc % NWS;
              tc::mbar_wait(&w_empty[slot], ((wc / NWS) & 1) ^ 1);
              tc::mbar_expect_tx(&w_full[slot], TAPB);
              tc::bulk_load(Wbase + slot * TAPB, wsrc + (size_t)(k3 ? kd * 9 + t9 : (kHW3 ? t9 : 0)) * TAPB, TAPB, &w_full[slot]);
            }
          }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const bool leader = tc::elect_one();
    const uint32_t a_lo0 = tc::desc_lo(tc::smem_u32(Abase), LBO_A), a_hi = tc::desc_hi(SBO_A);
    const uint32_t b_lo0 = tc::desc_lo(tc::smem_u32(Wbase), LBO_B), b_hi = tc::desc_hi(SBO_B);
    if (kResident && cta_s < p.items) tc::mbar_wait(&w_full[0], 0);
    uint32_t g_base = 0, acc_it = 0, wc = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int din0 = max(dlo - halo, 0), din1 = min(dhi - 1 + halo, p.D - 1);
      for (int d_out = dlo; d_out < dhi; ++d_out, ++acc_it) {
        const uint32_t as = acc_it & 1;
        tc::mbar_wait(&acc_empty[as], ((acc_it >> 1) & 1) ^ 1);
        tc::fence_after_sync();
        const uint32_t tmem_d = tmem_base + as * N;
        uint32_t accumulate = 0;
#pragma unroll 1
        for (int kd = (k3 ? 0 : 1); kd < (k3 ? 3 : 2); ++kd) {
          const int d_in = d_out + kd - 1;
          if (d_in < 0 || d_in >= p.D) continue;
          const uint32_t gs = g_base + (uint32_t)(d_in - din0), slot = gs % NS;
          tc::mbar_wait(&a_full[slot], (gs / NS) & 1);
          tc::fence_after_sync();
          const uint32_t a_lo = a_lo0 + slot * (SLICE >> 4);
#pragma unroll
          for (int t9 = (kHW3 ? 0 : 4); t9 < (kHW3 ? 9 : 5); ++t9) {
            const int kh = t9 / 3, kw = t9 - 3 * kh;
            uint32_t b_lo, wslot = 0;
            if (kResident) b_lo = b_lo0 + (uint32_t)(k3 ? kd * 9 + t9 : (kHW3 ? t9 : 0)) * (TAPB >> 4);
            else {
              wslot = wc % NWS;
              tc::mbar_wait(&w_full[wslot], (wc / NWS) & 1);
              tc::fence_after_sync();
              b_lo = b_lo0 + wslot * (TAPB >> 4);
            }
            if (leader) {
#pragma unroll
              for (int ks = 0; ks < KS; ++ks) {
                tc::mma_bf16_lohi(tmem_d, a_lo + (uint32_t)((kh * WW + kw) * 16 + ks * 2 * LBO_A) / 16, a_hi,
                                  b_lo + (uint32_t)(ks * 2 * LBO_B) / 16, b_hi, IDESC, accumulate);
                accumulate = 1;
              }
              if (!kResident) tc::mma_commit(&w_empty[wslot]);
            }
            accumulate = 1;
            if (!kResident) ++wc;
          }
        }
        if (leader) {
          tc::mma_commit(&acc_full[as]);
          if (k3) {   // input slices no later output slice of this item needs
            if (d_out - 1 >= din0) tc::mma_commit(&a_empty[(g_base + (uint32_t)(d_out - 1 - din0)) % NS]);
            if (d_out == dhi - 1) {
              tc::mma_commit(&a_empty[(g_base + (uint32_t)(d_out - din0)) % NS]);
              if (d_out + 1 <= din1) tc::mma_commit(&a_empty[(g_base + (uint32_t)(d_out + 1 - din0)) % NS]);
            }
          } else {
            tc::mma_commit(&a_empty[(g_base + (uint32_t)(d_out - din0)) % NS]);
          }
        }
        __syncwarp();
      }
      g_base += (uint32_t)(din1 - din0 + 1);
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int e = warp - 4, m = e * 32 + lane, hh = m >> 3, ww = m & 7;
    uint32_t acc_it = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int h = h0 + hh, w = w0 + ww;
      const bool valid = h < p.H && w < p.W;
      for (int d_out = dlo; d_out < dhi; ++d_out, ++acc_it) {
        const uint32_t as = acc_it & 1;
        tc::mbar_wait(&acc_full[as], (acc_it >> 1) & 1);
        tc::fence_after_sync();
#pragma unroll 1
        for (int j = 0; j < N / 32; ++j) {
          float v[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(e * 32) << 16) + as * N + j * 32, v);
          if (j == N / 32 - 1) {            // accumulator fully read: hand the TMEM buffer back before the stores
            tc::fence_before_sync();
            tc::mbar_arrive(&acc_empty[as]);
          }
          if (!valid || nt * N + j * 32 >= p.cout_valid) continue;
          epilogue_store32<EPI>(p, v, s_scale + j * 32, s_shift + j * 32, nt * N + j * 32, b, d_out, h, w, p.D, p.H, p.W, nullptr);
        }
      }
    }
  }
  TC_KERNEL_EPILOGUE()
}

// =====================================================================================================================
// s1f: Conv3d k3 s1 p1 for the narrow layers (Cout = 32 / 64) with the three DEPTH taps folded into the GEMM's N.
// An N = 32 MMA reads 4 KB of A and 1 KB of B from shared memory for 128x32x16 MACs, so the shared-memory read port (not the
// tensor pipe) bounds it.  Here ONE MMA of N = 3*Cout multiplies an in-plane tap of an input slice by the weights of all three
// depth taps: column block j of the product belongs to output depth d_in - 1 + j (kd = 2 - j).  The accumulators of
// consecutive output depths are consecutive column blocks of a ring over all 512 TMEM columns, so the three partial
// products land directly in the three accumulators they belong to: 3x fewer MMAs and 3x fewer A bytes per MAC; every input
// slice is used exactly once.  One instruction carries one accumulate flag, so blocks are ALWAYS accumulated into and the
// epilogue warps write zeros back (tcgen05.st) into a block after draining it.  Where the three blocks wrap around the ring
// (or the depth range is cut by the volume / chunk boundary) the MMA is split into two narrower ones / narrowed.
// Weights: [9 in-plane taps][CIN/8][3*N (j*N + co)][8].
// =====================================================================================================================
// SP (bf16x3 split route, in-kernel): the input is a split tensor (hi batches [0,B) | lo batches [B,2B)), a staged slice holds the
// hi chunks followed by the lo chunks (two TMA boxes), the weights of a tap are [hi: CIN/8 chunks][lo: CIN/8 chunks], and every
// K step issues three MMAs into the same accumulator: x_hi*w_hi + x_lo*w_hi + x_hi*w_lo (fp32-accurate product, fp32 accumulation).
// EPG: epilogue warp groups (4 warps each, 128 + 128*EPG threads): with EPG = 2 the groups drain alternate accumulator blocks, so the
// TMEM load -> affine -> store latency of one output slice overlaps the next one's.
// NISS = 2 (resident weights): two issuer warps take alternate slices IN TURN (a turn barrier keeps the MMAs in slice order, so the
// accumulation order and the results are those of one issuer).  A register read by a queued tcgen05.mma cannot be rewritten before
// that MMA is dispatched, so a lone issuer reaches its commits and its next record only once its ~7 queued MMAs have drained, and the
// pipe idles behind it (ncu: tensor data pipe 72 % busy); with two warps (two uniform-register files) one loads and converts its next
// record while the other's MMAs execute.  A commit covers only the committing thread's MMAs, so an accumulator block is full after
// ONE arrival from each issuer (acc_full count 2): from the owner of slice d, and from the owner of the block's last other
// contributor (slice d+1, or d-1 where the volume ends at d).
template <int CIN, int N, int NS, int NWS, bool SP = false, int EPG = 1, int NISS = (NWS == 9 ? 2 : 1)>
__global__ void __launch_bounds__(128 + 128 * EPG, 1) conv3d_tc_s1f_kernel(const __grid_constant__ CUtensorMap tmA, const TcP p) {
  constexpr bool kResident = (NWS == 9);
  constexpr uint32_t HALF_A = (CIN / 8) * TILE_B, HALF_B = CIN * 3 * N * 2;      // one operand half (hi or lo) of a slice / tap
  constexpr uint32_t SLICE = (SP ? 2 : 1) * HALF_A;
  constexpr uint32_t TAPB = (SP ? 2 : 1) * HALF_B;
  constexpr int KS = CIN / 16;
  constexpr uint32_t LBO_A = TILE_B, SBO_A = WW * 16, LBO_B = 3 * N * 16, SBO_B = 128;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t NB = 512 / N;                     // accumulator blocks in the TMEM ring
  constexpr uint32_t PLN = 4;                          // slice records in flight between the planner and the MMA issuer
  static_assert(NISS == 1 || NWS == 9, "two issuers need resident weights (the streamed tap ring has one consumer)");
  __shared__ __align__(16) uint32_t plan[NISS][PLN][12];
  __shared__ __align__(8) uint64_t plan_full[NISS][PLN], plan_empty[NISS][PLN], turn[2];
  TC_KERNEL_PROLOGUE_NF(NS, NWS, kResident, NB, NISS)
  uint8_t* Abase = smem;
  uint8_t* Wbase = smem + NS * SLICE;
  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < NISS * PLN; ++i) { tc::mbar_init(&plan_full[0][i], 1); tc::mbar_init(&plan_empty[0][i], 1); }
    tc::mbar_init(&turn[0], 1); tc::mbar_init(&turn[1], 1);
    tc::fence_barrier_init();
  }
  if (warp >= 4 && warp < 8) {                         // all accumulator blocks start out zero
#pragma unroll 1
    for (uint32_t c = 0; c < 512; c += 32) tc::tmem_zero32(tmem_base + ((uint32_t)((warp - 4) * 32) << 16) + c);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();

  if (kResident && warp == 3) {                        // resident weights: loaded once; warp 3 then becomes the second MMA issuer
    if (lane == 0 && cta_s < p.items) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w);
      tc::mbar_expect_tx(&w_full[0], 9 * TAPB);
      for (int t9 = 0; t9 < 9; ++t9) tc::bulk_load(Wbase + t9 * TAPB, wsrc + (size_t)t9 * TAPB, TAPB, &w_full[0]);
    }
    __syncwarp();
  }
  if (warp == 0 && lane == 0) {
    // ===== input-slice producer =====
    uint32_t g = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int din0 = max(dlo - 1, 0), din1 = min(dhi, p.D - 1);
      for (int d_in = din0; d_in <= din1; ++d_in, ++g) {
        const uint32_t slot = g % NS;
        tc::mbar_wait(&a_empty[slot], ((g / NS) & 1) ^ 1);
        tc::mbar_expect_tx(&a_full[slot], SLICE);
        tc::tma_load_4d(Abase + slot * SLICE, &tmA, &a_full[slot], (w0 - 1) * 8, h0 - 1, d_in, b * (CIN / 8));
        if (SP) tc::tma_load_4d(Abase + slot * SLICE + HALF_A, &tmA, &a_full[slot], (w0 - 1) * 8, h0 - 1, d_in, (b + p.B) * (CIN / 8));
      }
    }
  } else if (!kResident && warp == 3 && lane == 0) {
    // ===== weight producer (streamed tap ring) =====
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w);
    {
      uint32_t wc = 0;
      for (int s = cta_s; s < p.items; s += cta_stride) {
        int b, h0, w0, dlo, dhi;
        decode_item(p, s, b, h0, w0, dlo, dhi);
        const int din0 = max(dlo - 1, 0), din1 = min(dhi, p.D - 1);
        for (int d_in = din0; d_in <= din1; ++d_in)
          for (int t9 = 0; t9 < 9; ++t9, ++wc) {
            const uint32_t slot = wc % NWS;
            tc::mbar_wait(&w_empty[slot], ((wc / NWS) & 1) ^ 1);
            tc::mbar_expect_tx(&w_full[slot], TAPB);
            tc::bulk_load(Wbase + slot * TAPB, wsrc + (size_t)t9 * TAPB, TAPB, &w_full[slot]);
          }
      }
    }
  } else if (warp == 2) {
    // ===== planner: everything the MMA issuer needs to know about a slice, and every wait it would have to do =====
    // (ncu r02_s1f / r02_k9: the issuing thread never waited on a barrier, yet the tensor pipe idled a third of the time.  A
    // register read by a queued tcgen05.mma cannot be rewritten before that MMA is dispatched, so the issuer's own per-slice
    // bookkeeping -- ring indices, descriptor selection, barrier addresses, ~190 instructions -- only started once the queue of
    // ~7 pending MMAs had drained, and the pipe then sat idle behind it.  This warp does that work one or more slices ahead and
    // hands the issuer a 48-byte record; the issuer's path between two batches of MMAs shrinks to one wait and three loads.)
    const uint32_t a_lo0 = tc::desc_lo(tc::smem_u32(Abase), LBO_A);
    const uint32_t bar_af = tc::smem_u32(&a_full[0]), bar_ae = tc::smem_u32(&a_empty[0]), bar_cf = tc::smem_u32(&acc_full[0]),
                   bar_ce = tc::smem_u32(&acc_empty[0]), bar_pf = tc::smem_u32(&plan_full[0][0]), bar_pe = tc::smem_u32(&plan_empty[0][0]);
    uint32_t g = 0;              // slice counter: slice g belongs to issuer g % NISS, as record g / NISS of its ring
    uint32_t acc_base = 0;       // unwrapped ring index of the accumulator of (this item, dlo)
    uint32_t acquired = 0;       // accumulator blocks handed to the MMAs so far (unwrapped)
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int din0 = max(dlo - 1, 0), din1 = min(dhi, p.D - 1);
#pragma unroll 1
      for (int d_in = din0; d_in <= din1; ++d_in, ++g) {
        const int j0 = max(0, dlo - d_in + 1), j1 = min(3, dhi - d_in + 1);       // column blocks [j0, j1) exist
        const uint32_t u0 = acc_base + (uint32_t)(d_in - 1 + j0 - dlo), nb = (uint32_t)(j1 - j0);
        while (acquired < u0 + nb) {
          tc::mbar_wait_a(bar_ce + (acquired % NB) * 8, ((acquired / NB) & 1) ^ 1);
          ++acquired;
        }
        const uint32_t slot = g % NS;
        tc::mbar_wait_a(bar_af + slot * 8, (g / NS) & 1);
        const uint32_t blk = u0 % NB, n1 = min(nb, NB - blk), n2 = nb - n1;
        const uint32_t own = g % NISS, q = g / NISS, pk = own * PLN + q % PLN;
        tc::mbar_wait_a(bar_pe + pk * 8, ((q / PLN) & 1) ^ 1);
        if (lane == 0) {
          uint32_t* r = plan[0][pk];
          auto full_of = [&](int d) { return bar_cf + ((acc_base + (uint32_t)(d - dlo)) % NB) * 8; };     // acc_full of output depth d
          r[0] = tmem_base + blk * N;
          r[1] = n1 == 1 ? tc::make_idesc_bf16(128, N) : n1 == 2 ? tc::make_idesc_bf16(128, 2 * N) : tc::make_idesc_bf16(128, 3 * N);
          r[2] = a_lo0 + slot * (SLICE >> 4);
          r[3] = (uint32_t)j0 * (N / 8) * (SBO_B >> 4);
          r[4] = n2;
          r[5] = n2 == 1 ? tc::make_idesc_bf16(128, N) : tc::make_idesc_bf16(128, 2 * N);
          r[6] = (uint32_t)(j0 + n1) * (N / 8) * (SBO_B >> 4);
          r[7] = bar_ae + slot * 8;
          if (NISS == 1) {       // block d-1 is complete after slice d; the last block also after the last slice of a volume's end
            r[8] = d_in - 1 >= dlo ? full_of(d_in - 1) : 0u;
            r[9] = (d_in == din1 && din1 == dhi - 1) ? full_of(d_in) : 0u;
            r[11] = 0u;
          } else {               // one arrival per issuer and block (see the kernel comment); a one-slice item arrives twice
            r[8] = (d_in >= dlo && d_in < dhi) ? full_of(d_in) : 0u;
            r[9] = d_in - 1 >= dlo ? full_of(d_in - 1) : 0u;
            r[11] = (din1 == dhi - 1 && ((d_in == dhi - 2 && d_in >= din0) || (d_in == dhi - 1 && dhi - 2 < din0))) ? full_of(dhi - 1) : 0u;
          }
          r[10] = 1u;
          tc::mbar_arrive_a(bar_pf + pk * 8);          // release: the record, and the a_full / acc_empty phases observed above
        }
        __syncwarp();
      }
      acc_base += (uint32_t)(dhi - dlo);
    }
    for (uint32_t e = 0; e < (uint32_t)NISS; ++e, ++g) {      // end markers, one per issuer
      const uint32_t own = g % NISS, q = g / NISS, pk = own * PLN + q % PLN;
      tc::mbar_wait_a(bar_pe + pk * 8, ((q / PLN) & 1) ^ 1);
      if (lane == 0) {
        plan[0][pk][10] = 0u;
        tc::mbar_arrive_a(bar_pf + pk * 8);
      }
      __syncwarp();
    }
  } else if (warp == 1 || (NISS == 2 && warp == 3)) {
    // ===== MMA issuer(s) =====
    const uint32_t me = warp == 1 ? 0u : 1u;
    const bool leader = tc::elect_one();
    const uint32_t a_hi = tc::desc_hi(SBO_A);
    const uint32_t b_lo0 = tc::desc_lo(tc::smem_u32(Wbase), LBO_B), b_hi = tc::desc_hi(SBO_B);
    if (kResident && cta_s < p.items) tc::mbar_wait(&w_full[0], 0);
    const uint32_t bar_wf = tc::smem_u32(&w_full[0]), bar_we = tc::smem_u32(&w_empty[0]), bar_pf = tc::smem_u32(&plan_full[me][0]),
                   bar_pe = tc::smem_u32(&plan_empty[me][0]), bar_turn = tc::smem_u32(&turn[0]);
    uint32_t wc = 0;
#pragma unroll 1
    for (uint32_t k = 0;; ++k) {
      const uint32_t pk = k % PLN;
      tc::mbar_wait_a(bar_pf + pk * 8, (k / PLN) & 1);
      const uint4 r0 = *reinterpret_cast<const uint4*>(&plan[me][pk][0]), r1 = *reinterpret_cast<const uint4*>(&plan[me][pk][4]),
                  r2 = *reinterpret_cast<const uint4*>(&plan[me][pk][8]);
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_a(bar_pe + pk * 8);
      if (r2.z == 0u) break;
      if (NISS == 2) tc::mbar_wait_a(bar_turn + me * 8, (k & 1) ^ (me ^ 1));      // issuer 0 starts; then strictly alternating
      tc::fence_after_sync();
      const uint32_t d1 = r0.x, id1 = r0.y, a_lo = r0.z, brow1 = r0.w, n2 = r1.x, id2 = r1.y, brow2 = r1.z, d2 = tmem_base;
      auto issue = [&](auto wrap_tag) {                // the ring-wrap case, which doubles every MMA, is a separate copy of the loop
        constexpr bool WRAP = decltype(wrap_tag)::value;
#pragma unroll
        for (int t9 = 0; t9 < 9; ++t9) {
          const int kh = t9 / 3, kw = t9 - 3 * kh;
          uint32_t b_lo, wslot = 0;
          if (kResident) b_lo = b_lo0 + (uint32_t)t9 * (TAPB >> 4);
          else {
            wslot = wc % NWS;
            tc::mbar_wait_a(bar_wf + wslot * 8, (wc / NWS) & 1);
            tc::fence_after_sync();
            b_lo = b_lo0 + wslot * (TAPB >> 4);
          }
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              const uint32_t a = a_lo + (uint32_t)((kh * WW + kw) * 16 + ks * 2 * LBO_A) / 16;
              const uint32_t bb = b_lo + (uint32_t)(ks * 2 * LBO_B) / 16;
              tc::mma_bf16_lohi(d1, a, a_hi, bb + brow1, b_hi, id1, 1u);
              if (WRAP) tc::mma_bf16_lohi(d2, a, a_hi, bb + brow2, b_hi, id2, 1u);
              if (SP) {
                tc::mma_bf16_lohi(d1, a + (HALF_A >> 4), a_hi, bb + brow1, b_hi, id1, 1u);
                if (WRAP) tc::mma_bf16_lohi(d2, a + (HALF_A >> 4), a_hi, bb + brow2, b_hi, id2, 1u);
                tc::mma_bf16_lohi(d1, a, a_hi, bb + (HALF_B >> 4) + brow1, b_hi, id1, 1u);
                if (WRAP) tc::mma_bf16_lohi(d2, a, a_hi, bb + (HALF_B >> 4) + brow2, b_hi, id2, 1u);
              }
            }
            if (!kResident) tc::mma_commit_a(bar_we + wslot * 8);
          }
          if (!kResident) ++wc;
        }
      };
      if (n2) issue(std::true_type{});
      else issue(std::false_type{});
      if (leader) {
        if (NISS == 2) tc::mbar_arrive_a(bar_turn + (me ^ 1) * 8);     // this slice's MMAs are in the queue: the other issuer's turn
        tc::mma_commit_a(r1.w);
        if (r2.x) tc::mma_commit_a(r2.x);
        if (r2.y) tc::mma_commit_a(r2.y);
        if (NISS == 2 && r2.w) tc::mma_commit_a(r2.w);
      }
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 4 + 4 * EPG) {
    // ===== epilogue (warp & 3 = the TMEM lane quarter this warp may read) =====
    const int e = warp & 3, eg = (warp - 4) >> 2, m = e * 32 + lane, hh = m >> 3, ww = m & 7;
    uint32_t u = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int h = h0 + hh, w = w0 + ww;
      const bool valid = h < p.H && w < p.W;
      for (int d_out = dlo; d_out < dhi; ++d_out, ++u) {
        if (EPG > 1 && (int)(u % EPG) != eg) continue;
        const uint32_t blk = u % NB;
        tc::mbar_wait(&acc_full[blk], (u / NB) & 1);
        tc::fence_after_sync();
#pragma unroll 1
        for (int j = 0; j < N / 32; ++j) {
          float v[32];
          const uint32_t ta = tmem_base + ((uint32_t)(e * 32) << 16) + blk * N + j * 32;
          tc::tmem_ld32(ta, v);
          tc::tmem_zero32(ta);
          if (j == N / 32 - 1) {            // block drained and zeroed: hand it back before the stores
            tc::fence_before_sync();
            tc::mbar_arrive(&acc_empty[blk]);
          }
          if (!valid) continue;
          epilogue_store32<SP ? 1 : 0>(p, v, s_scale + j * 32, s_shift + j * 32, j * 32, b, d_out, h, w, p.D, p.H, p.W, nullptr);
        }
      }
    }
  }
  TC_KERNEL_EPILOGUE()
}

// =====================================================================================================================
// k9: concat_stem (Conv3d 64 -> 32 k3 s1 + BN + ReLU, gate) with the sparse concat volume generated INSIDE the kernel
// (SemStereo.py:241-244, 316-319): the 403 MB/pair volume  V[c,k] = [cf_l[c] | cf_r[c](x - d_k)] * a_k  never exists in HBM.
// Same GEMM as s1f<64,32> (depth taps folded into N, TMEM accumulator ring), but the A operand of input slice k is WRITTEN by
// four producer warps instead of fetched by TMA: per work item one TMA round stages the bf16 cf_l halo tile and the cf_r rows
// with a window wide enough for every disparity bin (41 px); the disp_topk / att_topk values of a thread's halo pixels are
// prefetched from global memory one slice ahead; for each slice k the producers scale the left chunk and the
// disparity-shifted right chunk of every halo pixel by a_k and store them in the [chunk][18x10 px][16 B] K-major layout the
// MMAs read (generic-proxy stores -> fence.proxy.async -> mbarrier).  The weights stay resident (110 KB): shared memory is
// the bottleneck resource of this layer, so nothing is streamed through it that does not have to be.
// Out-of-image pixels have a_k = 0 and out-of-image right taps are 0 through the TMA zero fill: exactly the zero padding of
// the convolution and of grid_sample.  Disparity samples are the integer bins dmin .. dmin+31 (disparity_sample_topk, :305);
// the reference's bilinear weights differ from this integer shift by the 1e-6 of its fp32 grid round trip, below bf16 resolution.
// =====================================================================================================================
constexpr int K9_NS = 3, K9_RW = 41, K9_BINS = 32;
constexpr uint32_t K9_SLICE = 8 * TILE_B, K9_TAPB = 64 * 96 * 2;
constexpr uint32_t K9_OFF_W = K9_NS * K9_SLICE, K9_OFF_R = K9_OFF_W + 9 * K9_TAPB, K9_SMEM = K9_OFF_R + 4 * HH * K9_RW * 16;
static_assert(K9_OFF_W % 128 == 0 && K9_OFF_R % 128 == 0, "TMA destinations");
static_assert(K9_SMEM <= 227 * 1024 - 2048, "shared memory budget");

struct K9P {
  TcP t;              // the conv part (w = s1f-packed weights [9][8][96][8]); t.D = number of samples K
  const uint4* cf_l;  // bf16 blocked (B,4,H,W,8): read directly by the producer threads (their pixels' 64 bytes, once per item)
  const float* disp;  // (B,K,H,W) integer-valued samples
  const float* att;   // (B,K,H,W)
  int dmin;           // lowest disparity bin (-(maxdisp/4) signed, 0 unsigned); samples are integers in [dmin, dmin + 31]
};

__device__ __forceinline__ uint4 scale8(const uint4 q, float a) {
  const uint32_t u[4] = {q.x, q.y, q.z, q.w};
  uint32_t r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = tc::pack_bf16x2(__uint_as_float(u[i] << 16) * a, __uint_as_float(u[i] & 0xffff0000u) * a);
  return make_uint4(r[0], r[1], r[2], r[3]);
}

__global__ void __launch_bounds__(384, 1) concat_stem_k9_kernel(const __grid_constant__ CUtensorMap tmR, const K9P kp) {
  constexpr int N = 32, KS = 4;
  constexpr uint32_t LBO_A = TILE_B, SBO_A = WW * 16, LBO_B = 3 * N * 16, SBO_B = 128;
  constexpr uint32_t NB = 512 / N;
  const TcP& p = kp.t;
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr uint32_t PLN = 4;                          // planner -> issuer records in flight, per issuer (see s1f)
  __shared__ __align__(16) uint32_t plan[2][PLN][12];
  __shared__ __align__(8) uint64_t a_full[K9_NS], a_empty[K9_NS], w_full, acc_full[NB], acc_empty[NB], st_full, st_empty;
  __shared__ __align__(8) uint64_t plan_full[2][PLN], plan_empty[2][PLN], turn[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_scale[N], s_shift[N];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta_s = blockIdx.x, cta_stride = gridDim.x;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    s_scale[i] = p.scale ? __ldg(p.scale + i) : 1.0f;
    s_shift[i] = p.shift ? __ldg(p.shift + i) : 0.0f;
  }
  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmR);
    for (int i = 0; i < K9_NS; ++i) { tc::mbar_init(&a_full[i], 128); tc::mbar_init(&a_empty[i], 1); }
    tc::mbar_init(&w_full, 1);
    for (uint32_t i = 0; i < NB; ++i) { tc::mbar_init(&acc_full[i], 2); tc::mbar_init(&acc_empty[i], 128); }      // one arrival per issuer
    tc::mbar_init(&st_full, 1); tc::mbar_init(&st_empty, 128);
    for (uint32_t i = 0; i < 2 * PLN; ++i) { tc::mbar_init(&plan_full[0][i], 1); tc::mbar_init(&plan_empty[0][i], 1); }
    tc::mbar_init(&turn[0], 1); tc::mbar_init(&turn[1], 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(&tmem_base_s, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  if (warp >= 4 && warp < 8) {
#pragma unroll 1
    for (uint32_t c = 0; c < 512; c += 32) tc::tmem_zero32(tmem_base + ((uint32_t)((warp - 4) * 32) << 16) + c);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  uint8_t* Abase = smem;
  uint8_t* Wbase = smem + K9_OFF_W;
  const int dhi_bin = kp.dmin + K9_BINS - 1;                 // largest disparity a sample can take

  if (warp == 3) {                                       // resident weights, one bulk copy per tap; warp 3 then is the second MMA issuer
    if (lane == 0 && cta_s < p.items) {
      tc::mbar_expect_tx(&w_full, 9 * K9_TAPB);
      for (int t9 = 0; t9 < 9; ++t9)
        tc::bulk_load(Wbase + t9 * K9_TAPB, reinterpret_cast<const uint8_t*>(p.w) + (size_t)t9 * K9_TAPB, K9_TAPB, &w_full);
    }
    __syncwarp();
  }
  if (warp == 0 && lane == 0) {
    // ===== stager: per item, the cf_r window =====
    uint32_t it = 0;
    for (int s = cta_s; s < p.items; s += cta_stride, ++it) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      tc::mbar_wait(&st_empty, (it & 1) ^ 1);
      tc::mbar_expect_tx(&st_full, K9_SMEM - K9_OFF_R);
      tc::tma_load_4d(smem + K9_OFF_R, &tmR, &st_full, 0, w0 - 1 - dhi_bin, h0 - 1, b * 4);
    }
  } else if (warp >= 8) {
    // ===== A producers: the slice k of the sparse concat volume for the 180 halo pixels =====
    const int pt = threadIdx.x - 256;
    const uint4* Rs = reinterpret_cast<const uint4*>(smem + K9_OFF_R);     // [4][18][41]
    const size_t HW = (size_t)p.H * p.W;
    const int px1 = pt + 128;                                                // this thread's halo pixels: pt and (if < 180) pt + 128
    const int r0 = pt / WW, c0 = pt - r0 * WW, r1 = px1 / WW, c1 = px1 - r1 * WW;
    uint32_t g = 0, it = 0;
    for (int s = cta_s; s < p.items; s += cta_stride, ++it) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int din0 = max(dlo - 1, 0), din1 = min(dhi, p.D - 1);
      // sample values of this thread's pixels (a = 0 outside the image: the convolution's zero padding), one slice ahead
      const int y0 = h0 - 1 + r0, x0 = w0 - 1 + c0, y1 = h0 - 1 + r1, x1 = w0 - 1 + c1;
      const bool in0 = y0 >= 0 && y0 < p.H && x0 >= 0 && x0 < p.W;
      const bool in1 = px1 < HH * WW && y1 >= 0 && y1 < p.H && x1 >= 0 && x1 < p.W;
      const size_t o0 = (size_t)b * p.D * HW + (size_t)(in0 ? y0 : 0) * p.W + (in0 ? x0 : 0);
      const size_t o1 = (size_t)b * p.D * HW + (size_t)(in1 ? y1 : 0) * p.W + (in1 ? x1 : 0);
      float na0 = in0 ? __ldg(kp.att + o0 + (size_t)din0 * HW) : 0.0f, nd0 = in0 ? __ldg(kp.disp + o0 + (size_t)din0 * HW) : 0.0f;
      float na1 = in1 ? __ldg(kp.att + o1 + (size_t)din0 * HW) : 0.0f, nd1 = in1 ? __ldg(kp.disp + o1 + (size_t)din0 * HW) : 0.0f;
      // the left features of this thread's pixels stay in registers for the whole item (they are the same for every sample k): the
      // 11.5 KB shared-memory tile they used to be staged in is what pays for the third A slot
      uint4 Lr[2][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const size_t cj = ((size_t)b * 4 + j) * HW;
        Lr[0][j] = in0 ? __ldg(kp.cf_l + cj + (size_t)y0 * p.W + x0) : make_uint4(0u, 0u, 0u, 0u);
        Lr[1][j] = in1 ? __ldg(kp.cf_l + cj + (size_t)y1 * p.W + x1) : make_uint4(0u, 0u, 0u, 0u);
      }
      tc::mbar_wait(&st_full, it & 1);
#pragma unroll 1
      for (int d_in = din0; d_in <= din1; ++d_in, ++g) {
        const float a0 = na0, d0 = nd0, a1 = na1, d1 = nd1;
        if (d_in < din1) {
          const size_t ko = (size_t)(d_in + 1) * HW;
          na0 = in0 ? __ldg(kp.att + o0 + ko) : 0.0f; nd0 = in0 ? __ldg(kp.disp + o0 + ko) : 0.0f;
          na1 = in1 ? __ldg(kp.att + o1 + ko) : 0.0f; nd1 = in1 ? __ldg(kp.disp + o1 + ko) : 0.0f;
        }
        const uint32_t slot = g % K9_NS;
        tc::mbar_wait(&a_empty[slot], ((g / K9_NS) & 1) ^ 1);
        uint4* At = reinterpret_cast<uint4*>(Abase + slot * K9_SLICE);       // [8][180]
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          const int px = rep ? px1 : pt, r = rep ? r1 : r0, c = rep ? c1 : c0;
          if (px < HH * WW) {
            const float a = rep ? a1 : a0;
            int cs = c + dhi_bin - __float2int_rn(rep ? d1 : d0);
            const bool ok = cs >= 0 && cs < K9_RW;                           // always true for samples in [dmin, dmin + 31]
            cs = ok ? cs : 0;
            const float ar = ok ? a : 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              At[j * (HH * WW) + px] = scale8(Lr[rep][j], a);
              At[(4 + j) * (HH * WW) + px] = scale8(Rs[(j * HH + r) * K9_RW + cs], ar);
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic stores -> visible to the tensor core's reads
        tc::mbar_arrive(&a_full[slot]);
      }
      tc::mbar_arrive(&st_empty);                                            // the staged item is no longer needed by this thread
    }
  } else if (warp == 2) {
    // ===== planner (as s1f): waits and per-slice bookkeeping, one 48-byte record per slice for the issuer that owns it =====
    const uint32_t a_lo0 = tc::desc_lo(tc::smem_u32(Abase), LBO_A);
    const uint32_t bar_af = tc::smem_u32(&a_full[0]), bar_ae = tc::smem_u32(&a_empty[0]), bar_cf = tc::smem_u32(&acc_full[0]),
                   bar_ce = tc::smem_u32(&acc_empty[0]), bar_pf = tc::smem_u32(&plan_full[0][0]), bar_pe = tc::smem_u32(&plan_empty[0][0]);
    uint32_t g = 0, acc_base = 0, acquired = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int din0 = max(dlo - 1, 0), din1 = min(dhi, p.D - 1);
#pragma unroll 1
      for (int d_in = din0; d_in <= din1; ++d_in, ++g) {
        const int j0 = max(0, dlo - d_in + 1), j1 = min(3, dhi - d_in + 1);
        const uint32_t u0 = acc_base + (uint32_t)(d_in - 1 + j0 - dlo), nb = (uint32_t)(j1 - j0);
        while (acquired < u0 + nb) {
          tc::mbar_wait_a(bar_ce + (acquired % NB) * 8, ((acquired / NB) & 1) ^ 1);
          ++acquired;
        }
        const uint32_t slot = g % K9_NS;
        tc::mbar_wait_a(bar_af + slot * 8, (g / K9_NS) & 1);
        const uint32_t blk = u0 % NB, n1 = min(nb, NB - blk), n2 = nb - n1;
        const uint32_t own = g & 1u, q = g >> 1, pk = own * PLN + q % PLN;
        tc::mbar_wait_a(bar_pe + pk * 8, ((q / PLN) & 1) ^ 1);
        if (lane == 0) {
          uint32_t* r = plan[0][pk];
          auto full_of = [&](int d) { return bar_cf + ((acc_base + (uint32_t)(d - dlo)) % NB) * 8; };
          r[0] = tmem_base + blk * N;
          r[1] = n1 == 1 ? tc::make_idesc_bf16(128, N) : n1 == 2 ? tc::make_idesc_bf16(128, 2 * N) : tc::make_idesc_bf16(128, 3 * N);
          r[2] = a_lo0 + slot * (K9_SLICE >> 4);
          r[3] = (uint32_t)j0 * (N / 8) * (SBO_B >> 4);
          r[4] = n2;
          r[5] = n2 == 1 ? tc::make_idesc_bf16(128, N) : tc::make_idesc_bf16(128, 2 * N);
          r[6] = (uint32_t)(j0 + n1) * (N / 8) * (SBO_B >> 4);
          r[7] = bar_ae + slot * 8;
          r[8] = (d_in >= dlo && d_in < dhi) ? full_of(d_in) : 0u;
          r[9] = d_in - 1 >= dlo ? full_of(d_in - 1) : 0u;
          r[10] = 1u;
          r[11] = (din1 == dhi - 1 && ((d_in == dhi - 2 && d_in >= din0) || (d_in == dhi - 1 && dhi - 2 < din0))) ? full_of(dhi - 1) : 0u;
          tc::mbar_arrive_a(bar_pf + pk * 8);
        }
        __syncwarp();
      }
      acc_base += (uint32_t)(dhi - dlo);
    }
    for (uint32_t e = 0; e < 2u; ++e, ++g) {             // end markers, one per issuer
      const uint32_t own = g & 1u, q = g >> 1, pk = own * PLN + q % PLN;
      tc::mbar_wait_a(bar_pe + pk * 8, ((q / PLN) & 1) ^ 1);
      if (lane == 0) {
        plan[0][pk][10] = 0u;
        tc::mbar_arrive_a(bar_pf + pk * 8);
      }
      __syncwarp();
    }
  } else if (warp == 1 || warp == 3) {
    // ===== MMA issuers (as s1f: alternate slices, in turn; weights resident) =====
    const uint32_t me = warp == 1 ? 0u : 1u;
    const bool leader = tc::elect_one();
    const uint32_t a_hi = tc::desc_hi(SBO_A);
    const uint32_t b_lo0 = tc::desc_lo(tc::smem_u32(Wbase), LBO_B), b_hi = tc::desc_hi(SBO_B);
    if (cta_s < p.items) tc::mbar_wait(&w_full, 0);
    const uint32_t bar_pf = tc::smem_u32(&plan_full[me][0]), bar_pe = tc::smem_u32(&plan_empty[me][0]), bar_turn = tc::smem_u32(&turn[0]);
#pragma unroll 1
    for (uint32_t k = 0;; ++k) {
      const uint32_t pk = k % PLN;
      tc::mbar_wait_a(bar_pf + pk * 8, (k / PLN) & 1);
      const uint4 r0 = *reinterpret_cast<const uint4*>(&plan[me][pk][0]), r1 = *reinterpret_cast<const uint4*>(&plan[me][pk][4]),
                  r2 = *reinterpret_cast<const uint4*>(&plan[me][pk][8]);
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_a(bar_pe + pk * 8);
      if (r2.z == 0u) break;
      tc::mbar_wait_a(bar_turn + me * 8, (k & 1) ^ (me ^ 1));
      tc::fence_after_sync();
      const uint32_t d1 = r0.x, id1 = r0.y, a_lo = r0.z, brow1 = r0.w, n2 = r1.x, id2 = r1.y, brow2 = r1.z, d2 = tmem_base;
      auto issue = [&](auto wrap_tag) {              // the ring-wrap case is a separate copy of the loop (see s1f)
        constexpr bool WRAP = decltype(wrap_tag)::value;
#pragma unroll
        for (int t9 = 0; t9 < 9; ++t9) {
          const int kh = t9 / 3, kw = t9 - 3 * kh;
          const uint32_t b_lo = b_lo0 + (uint32_t)t9 * (K9_TAPB >> 4);
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            const uint32_t a = a_lo + (uint32_t)((kh * WW + kw) * 16 + ks * 2 * LBO_A) / 16;
            const uint32_t bb = b_lo + (uint32_t)(ks * 2 * LBO_B) / 16;
            tc::mma_bf16_lohi(d1, a, a_hi, bb + brow1, b_hi, id1, 1u);
            if (WRAP) tc::mma_bf16_lohi(d2, a, a_hi, bb + brow2, b_hi, id2, 1u);
          }
        }
      };
      if (leader) {
        if (n2) issue(std::true_type{});
        else issue(std::false_type{});
        tc::mbar_arrive_a(bar_turn + (me ^ 1) * 8);    // this slice's MMAs are in the queue: the other issuer's turn
        tc::mma_commit_a(r1.w);
        if (r2.x) tc::mma_commit_a(r2.x);
        if (r2.y) tc::mma_commit_a(r2.y);
        if (r2.w) tc::mma_commit_a(r2.w);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===== epilogue (as s1f) =====
    const int e = warp - 4, m = e * 32 + lane, hh = m >> 3, ww = m & 7;
    uint32_t u = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int h = h0 + hh, w = w0 + ww;
      const bool valid = h < p.H && w < p.W;
      for (int d_out = dlo; d_out < dhi; ++d_out, ++u) {
        const uint32_t blk = u % NB;
        tc::mbar_wait(&acc_full[blk], (u / NB) & 1);
        tc::fence_after_sync();
        float v[32];
        const uint32_t ta = tmem_base + ((uint32_t)(e * 32) << 16) + blk * N;
        tc::tmem_ld32(ta, v);
        tc::tmem_zero32(ta);
        tc::fence_before_sync();
        tc::mbar_arrive(&acc_empty[blk]);
        if (!valid) continue;
        epilogue_store32(p, v, s_scale, s_shift, 0, b, d_out, h, w, p.D, p.H, p.W, nullptr);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================================
// s2: Conv3d k3 s2 p1 on the phase-split input.  A staged slice = (d-phase pd, half-res depth d') = 4 (h,w)-phase halo tiles.
// Per output depth d the slice uses are, in order: (1,d-1) [kd=0], (0,d) [kd=1], (1,d) [kd=2, kept for d+1's kd=0].
// =====================================================================================================================
template <int CIN, int N, int NS, int NWS, bool SP = false, int EPI = 0>      // SP: in-kernel bf16x3 split, see the s1f kernel
__global__ void __launch_bounds__(256, 1) conv3d_tc_s2_kernel(const __grid_constant__ CUtensorMap tmA, const TcP p) {
  constexpr bool kResident = (NWS == 27);
  constexpr int C8 = CIN / 8;
  constexpr uint32_t HALF_A = 4 * C8 * TILE_B, HALF_B = CIN * N * 2;
  constexpr uint32_t SLICE = (SP ? 2 : 1) * HALF_A;
  constexpr uint32_t TAPB = (SP ? 2 : 1) * HALF_B;
  constexpr int KS = CIN / 16;
  constexpr uint32_t LBO_A = TILE_B, SBO_A = WW * 16, LBO_B = N * 16, SBO_B = 128;
  constexpr uint32_t TMEM_COLS = tmem_cols_for(2 * N);
  constexpr uint32_t IDESC = tc::make_idesc_bf16(128, N);
  TC_KERNEL_PROLOGUE(NS, NWS, kResident)
  uint8_t* Abase = smem;
  uint8_t* Wbase = smem + NS * SLICE;

  if (warp == 0 && lane == 0) {
    uint32_t g = 0;
    auto load = [&](int b, int h0, int w0, int pd, int d) {
      const uint32_t slot = g % NS;
      tc::mbar_wait(&a_empty[slot], ((g / NS) & 1) ^ 1);
      tc::mbar_expect_tx(&a_full[slot], SLICE);
      tc::tma_load_4d(Abase + slot * SLICE, &tmA, &a_full[slot], (w0 - 1) * 8, h0 - 1, d, (b * 8 + pd * 4) * C8);
      if (SP) tc::tma_load_4d(Abase + slot * SLICE + HALF_A, &tmA, &a_full[slot], (w0 - 1) * 8, h0 - 1, d, ((b + p.B) * 8 + pd * 4) * C8);
      ++g;
    };
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      if (dlo > 0) load(b, h0, w0, 1, dlo - 1);
      for (int d = dlo; d < dhi; ++d) { load(b, h0, w0, 0, d); load(b, h0, w0, 1, d); }
    }
  } else if (warp == 3 && lane == 0) {
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) + (size_t)nt * 27 * TAPB;
    if (kResident) {
      if (cta_s < p.items) {
        tc::mbar_expect_tx(&w_full[0], 27 * TAPB);
        for (int tap = 0; tap < 27; ++tap) tc::bulk_load(Wbase + tap * TAPB, wsrc + (size_t)tap * TAPB, TAPB, &w_full[0]);
      }
    } else {
      uint32_t wc = 0;
      for (int s = cta_s; s < p.items; s += cta_stride) {
        int b, h0, w0, dlo, dhi;
        decode_item(p, s, b, h0, w0, dlo, dhi);
        for (int d_out = dlo; d_out < dhi; ++d_out)
          for (int kd = (d_out == 0 ? 1 : 0); kd < 3; ++kd)
            for (int t9 = 0; t9 < 9; ++t9, ++wc) {
              const uint32_t slot = wc % NWS;
              tc::mbar_wait(&w_empty[slot], ((wc / NWS) & 1) ^ 1);
              tc::mbar_expect_tx(&w_full[slot], TAPB);
              tc::bulk_load(Wbase + slot * TAPB, wsrc + (size_t)(kd * 9 + t9) * TAPB, TAPB, &w_full[slot]);
            }
      }
    }
  } else if (warp == 1) {
    const bool leader = tc::elect_one();
    const uint32_t a_lo0 = tc::desc_lo(tc::smem_u32(Abase), LBO_A), a_hi = tc::desc_hi(SBO_A);
    const uint32_t b_lo0 = tc::desc_lo(tc::smem_u32(Wbase), LBO_B), b_hi = tc::desc_hi(SBO_B);
    if (kResident && cta_s < p.items) tc::mbar_wait(&w_full[0], 0);
    uint32_t g_base = 0, acc_it = 0, wc = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const uint32_t hp = dlo > 0 ? 1u : 0u;
      for (int d_out = dlo; d_out < dhi; ++d_out, ++acc_it) {
        const uint32_t j = (uint32_t)(d_out - dlo);
        const uint32_t as = acc_it & 1;
        tc::mbar_wait(&acc_empty[as], ((acc_it >> 1) & 1) ^ 1);
        tc::fence_after_sync();
        const uint32_t tmem_d = tmem_base + as * N;
        uint32_t accumulate = 0;
#pragma unroll 1
        for (int kd = 0; kd < 3; ++kd) {
          if (kd == 0 && d_out == 0) continue;
          const uint32_t pos = kd == 0 ? (j == 0 ? 0u : hp + 2 * (j - 1) + 1) : hp + 2 * j + (uint32_t)(kd - 1);
          const uint32_t gs = g_base + pos, slot = gs % NS;
          tc::mbar_wait(&a_full[slot], (gs / NS) & 1);
          tc::fence_after_sync();
          const uint32_t a_lo = a_lo0 + slot * (SLICE >> 4);
#pragma unroll
          for (int t9 = 0; t9 < 9; ++t9) {
            const int kh = t9 / 3, kw = t9 - 3 * kh;
            const int ph = (kh == 1) ? 0 : 1, oh = (kh == 0) ? 0 : 1;     // input row 2h-1+kh = (phase, halo row offset)
            const int pw = (kw == 1) ? 0 : 1, ow = (kw == 0) ? 0 : 1;
            uint32_t b_lo, wslot = 0;
            if (kResident) b_lo = b_lo0 + (uint32_t)(kd * 9 + t9) * (TAPB >> 4);
            else {
              wslot = wc % NWS;
              tc::mbar_wait(&w_full[wslot], (wc / NWS) & 1);
              tc::fence_after_sync();
              b_lo = b_lo0 + wslot * (TAPB >> 4);
            }
            if (leader) {
#pragma unroll
              for (int ks = 0; ks < KS; ++ks) {
                const uint32_t aa = a_lo + (uint32_t)(((ph * 2 + pw) * C8 + 2 * ks) * LBO_A + (oh * WW + ow) * 16) / 16;
                const uint32_t bb = b_lo + (uint32_t)(ks * 2 * LBO_B) / 16;
                tc::mma_bf16_lohi(tmem_d, aa, a_hi, bb, b_hi, IDESC, accumulate);
                accumulate = 1;
                if (SP) {
                  tc::mma_bf16_lohi(tmem_d, aa + (HALF_A >> 4), a_hi, bb, b_hi, IDESC, 1u);
                  tc::mma_bf16_lohi(tmem_d, aa, a_hi, bb + (HALF_B >> 4), b_hi, IDESC, 1u);
                }
              }
              if (!kResident) tc::mma_commit(&w_empty[wslot]);
            }
            accumulate = 1;
            if (!kResident) ++wc;
          }
          // last use of this slice?  (1,d-1) and (0,d) always; (1,d) only at the end of the item
          if (leader && (kd < 2 || d_out == dhi - 1)) tc::mma_commit(&a_empty[slot]);
        }
        if (leader) tc::mma_commit(&acc_full[as]);
        __syncwarp();
      }
      g_base += hp + 2 * (uint32_t)(dhi - dlo);
    }
  } else if (warp >= 4) {
    const int e = warp - 4, m = e * 32 + lane, hh = m >> 3, ww = m & 7;
    uint32_t acc_it = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int h = h0 + hh, w = w0 + ww;
      const bool valid = h < p.H && w < p.W;
      for (int d_out = dlo; d_out < dhi; ++d_out, ++acc_it) {
        const uint32_t as = acc_it & 1;
        tc::mbar_wait(&acc_full[as], (acc_it >> 1) & 1);
        tc::fence_after_sync();
#pragma unroll 1
        for (int j = 0; j < N / 32; ++j) {
          float v[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(e * 32) << 16) + as * N + j * 32, v);
          if (j == N / 32 - 1) {
            tc::fence_before_sync();
            tc::mbar_arrive(&acc_empty[as]);
          }
          if (!valid || nt * N + j * 32 >= p.cout_valid) continue;
          epilogue_store32<(SP ? 1 : 0) | EPI>(p, v, s_scale + j * 32, s_shift + j * 32, nt * N + j * 32, b, d_out, h, w, p.D, p.H, p.W, nullptr);
        }
      }
    }
  }
  TC_KERNEL_EPILOGUE()
}

// =====================================================================================================================
// t2: ConvTranspose3d k3 s2 p1 op1.  Tile space = INPUT voxels; output voxel 2i+p per dim:  p=0: tap k=1 from input i;
// p=1: tap k=0 from input i+1 and tap k=2 from input i.  Per input depth i the 8 output phases are 8 accumulators in a row.
// =====================================================================================================================
// RS = depth of the ring that prefetches the skip-connection tiles (one TH x TW x N tile per output phase) by TMA: the layer
// is memory-heavy (it reads a full-resolution residual and writes a full-resolution output per 27/8 taps of math), and 128
// epilogue threads issuing just-in-time loads cannot keep enough bytes in flight; the ring keeps RS tiles ahead.
template <int CIN, int N, int NS, int NWS, int RS, bool SP = false, int EPI = 0>      // SP: in-kernel bf16x3 split, see the s1f kernel
__global__ void __launch_bounds__(256, 1) conv3d_tc_t2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmRes, const TcP p) {
  constexpr bool kResident = (NWS == 27);
  constexpr uint32_t HALF_R = (N / 8) * TH * TW * 16, HALF_A = (CIN / 8) * TILE_B, HALF_B = CIN * N * 2, HALF_S = N * N * 2;
  constexpr uint32_t RTILE = (SP ? 2 : 1) * HALF_R;           // bytes of one residual tile
  constexpr uint32_t SLICE = (SP ? 2 : 1) * HALF_A;
  constexpr uint32_t TAPB = (SP ? 2 : 1) * HALF_B;
  constexpr int KS = CIN / 16;
  constexpr uint32_t LBO_A = TILE_B, SBO_A = WW * 16, LBO_B = N * 16, SBO_B = 128;
  constexpr uint32_t TMEM_COLS = tmem_cols_for(2 * N);
  constexpr uint32_t IDESC = tc::make_idesc_bf16(128, N);
  __shared__ __align__(8) uint64_t res_full[RS], res_empty[RS], skip_full;
  constexpr uint32_t SKIPB = (SP ? 2 : 1) * HALF_S;           // bytes of the fused 1x1 skip weight
  if (threadIdx.x == 0) {      // made visible to the async proxy by the fence in the prologue (same thread)
    for (int i = 0; i < RS; ++i) { tc::mbar_init(&res_full[i], 1); tc::mbar_init(&res_empty[i], p.skip_w ? 1 : 128); }
    tc::mbar_init(&skip_full, 1);
  }
  TC_KERNEL_PROLOGUE(NS, NWS, kResident)
  uint8_t* Abase = smem;
  uint8_t* Wbase = smem + NS * SLICE;
  uint8_t* Rbase = Wbase + NWS * TAPB;
  uint8_t* Sbase = Rbase + RS * RTILE;
  const bool has_res = p.residual != nullptr;
  const bool fuse_skip = p.skip_w != nullptr;

  // tap (shift, k) lists of one dimension for output parity q: q=0 -> {(0,1)}, q=1 -> {(1,0),(0,2)}
  auto ntaps = [](int q) { return q ? 2 : 1; };
  auto tap_shift = [](int q, int t) { return q ? (t == 0 ? 1 : 0) : 0; };
  auto tap_k = [](int q, int t) { return q ? (t == 0 ? 0 : 2) : 1; };

  if (warp == 0 && lane == 0) {
    uint32_t g = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int din1 = min(dhi, p.D - 1);
      for (int d_in = dlo; d_in <= din1; ++d_in, ++g) {
        const uint32_t slot = g % NS;
        tc::mbar_wait(&a_empty[slot], ((g / NS) & 1) ^ 1);
        tc::mbar_expect_tx(&a_full[slot], SLICE);
        tc::tma_load_4d(Abase + slot * SLICE, &tmA, &a_full[slot], w0 * 8, h0, d_in, b * (CIN / 8));
        if (SP) tc::tma_load_4d(Abase + slot * SLICE + HALF_A, &tmA, &a_full[slot], w0 * 8, h0, d_in, (b + p.B) * (CIN / 8));
      }
    }
  } else if (warp == 3 && lane == 0) {
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) + (size_t)nt * 27 * TAPB;
    if (kResident) {
      if (cta_s < p.items) {
        tc::mbar_expect_tx(&w_full[0], 27 * TAPB);
        for (int tap = 0; tap < 27; ++tap) tc::bulk_load(Wbase + tap * TAPB, wsrc + (size_t)tap * TAPB, TAPB, &w_full[0]);
      }
    } else {
      uint32_t wc = 0;
      for (int s = cta_s; s < p.items; s += cta_stride) {
        int b, h0, w0, dlo, dhi;
        decode_item(p, s, b, h0, w0, dlo, dhi);
        for (int i = dlo; i < dhi; ++i)
          for (int ph8 = 0; ph8 < 8; ++ph8) {
            const int pd = ph8 >> 2, ph = (ph8 >> 1) & 1, pw = ph8 & 1;
            for (int td = 0; td < ntaps(pd); ++td) {
              if (i + tap_shift(pd, td) >= p.D) continue;
              for (int th = 0; th < ntaps(ph); ++th)
                for (int tw = 0; tw < ntaps(pw); ++tw, ++wc) {
                  const int tap = tap_k(pd, td) * 9 + tap_k(ph, th) * 3 + tap_k(pw, tw);
                  const uint32_t slot = wc % NWS;
                  tc::mbar_wait(&w_empty[slot], ((wc / NWS) & 1) ^ 1);
                  tc::mbar_expect_tx(&w_full[slot], TAPB);
                  tc::bulk_load(Wbase + slot * TAPB, wsrc + (size_t)tap * TAPB, TAPB, &w_full[slot]);
                }
            }
          }
      }
    }
  } else if (warp == 2 && lane == 0 && has_res) {
    // ===== skip-connection producer: one dense [N/8][TH][TW][16 B] tile of the phase-split residual per output phase =====
    if (fuse_skip && cta_s < p.items) {
      tc::mbar_expect_tx(&skip_full, SKIPB);
      tc::bulk_load(Sbase, p.skip_w, SKIPB, &skip_full);
    }
    uint32_t r = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      for (int i = dlo; i < dhi; ++i)
        for (int ph8 = 0; ph8 < 8; ++ph8, ++r) {
          const uint32_t slot = r % RS;
          tc::mbar_wait(&res_empty[slot], ((r / RS) & 1) ^ 1);
          tc::mbar_expect_tx(&res_full[slot], RTILE);
          tc::tma_load_4d(Rbase + slot * RTILE, &tmRes, &res_full[slot], w0 * 8, h0, i, (b * 8 + ph8) * (p.cout_valid / 8) + nt * (N / 8));
          if (SP) tc::tma_load_4d(Rbase + slot * RTILE + HALF_R, &tmRes, &res_full[slot], w0 * 8, h0, i,
                                  ((b + p.B) * 8 + ph8) * (p.cout_valid / 8) + nt * (N / 8));
        }
    }
  } else if (warp == 1) {
    const bool leader = tc::elect_one();
    const uint32_t a_lo0 = tc::desc_lo(tc::smem_u32(Abase), LBO_A), a_hi = tc::desc_hi(SBO_A);
    const uint32_t b_lo0 = tc::desc_lo(tc::smem_u32(Wbase), LBO_B), b_hi = tc::desc_hi(SBO_B);
    if (kResident && cta_s < p.items) tc::mbar_wait(&w_full[0], 0);
    if (fuse_skip && cta_s < p.items) tc::mbar_wait(&skip_full, 0);
    const uint32_t r_lo0 = tc::desc_lo(tc::smem_u32(Rbase), TH * TW * 16), r_hi = tc::desc_hi(128);
    const uint32_t s_lo0 = tc::desc_lo(tc::smem_u32(Sbase), LBO_B);
    uint32_t g_base = 0, acc_it = 0, wc = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int din1 = min(dhi, p.D - 1);
      for (int i = dlo; i < dhi; ++i) {
        const uint32_t gs0 = g_base + (uint32_t)(i - dlo), slot0 = gs0 % NS, slot1 = (gs0 + 1) % NS;
        tc::mbar_wait(&a_full[slot0], (gs0 / NS) & 1);
        if (i + 1 <= din1) tc::mbar_wait(&a_full[slot1], ((gs0 + 1) / NS) & 1);
        tc::fence_after_sync();
        // fully unrolled over the 8 output phases and their taps: every tap index / operand offset is an immediate, so the
        // issuing warp spends ~3 instructions per MMA (a rolled loop made this layer issue-bound at 4x the MMA time)
#pragma unroll
        for (int ph8 = 0; ph8 < 8; ++ph8, ++acc_it) {
          const int pd = ph8 >> 2, ph = (ph8 >> 1) & 1, pw = ph8 & 1;
          const uint32_t as = acc_it & 1;
          tc::mbar_wait(&acc_empty[as], ((acc_it >> 1) & 1) ^ 1);
          tc::fence_after_sync();
          const uint32_t tmem_d = tmem_base + as * N;
          uint32_t accumulate = 0;
#pragma unroll
          for (int td = 0; td < ntaps(pd); ++td) {
            const int sd = tap_shift(pd, td);
            if (i + sd >= p.D) continue;
            const uint32_t a_lo = a_lo0 + (sd ? slot1 : slot0) * (SLICE >> 4);
#pragma unroll
            for (int th = 0; th < ntaps(ph); ++th)
#pragma unroll
              for (int tw = 0; tw < ntaps(pw); ++tw) {
                const int tap = tap_k(pd, td) * 9 + tap_k(ph, th) * 3 + tap_k(pw, tw);
                const uint32_t a_off = (uint32_t)(tap_shift(ph, th) * WW + tap_shift(pw, tw));     // 16-byte units
                uint32_t b_lo, wslot = 0;
                if (kResident) b_lo = b_lo0 + (uint32_t)tap * (TAPB >> 4);
                else {
                  wslot = wc % NWS;
                  tc::mbar_wait(&w_full[wslot], (wc / NWS) & 1);
                  tc::fence_after_sync();
                  b_lo = b_lo0 + wslot * (TAPB >> 4);
                }
                if (leader) {
#pragma unroll
                  for (int ks = 0; ks < KS; ++ks) {
                    const uint32_t aa = a_lo + a_off + (uint32_t)(ks * 2 * LBO_A) / 16, bb = b_lo + (uint32_t)(ks * 2 * LBO_B) / 16;
                    tc::mma_bf16_lohi(tmem_d, aa, a_hi, bb, b_hi, IDESC, accumulate);
                    accumulate = 1;
                    if (SP) {
                      tc::mma_bf16_lohi(tmem_d, aa + (HALF_A >> 4), a_hi, bb, b_hi, IDESC, 1u);
                      tc::mma_bf16_lohi(tmem_d, aa, a_hi, bb + (HALF_B >> 4), b_hi, IDESC, 1u);
                    }
                  }
                  if (!kResident) tc::mma_commit(&w_empty[wslot]);
                }
                accumulate = 1;
                if (!kResident) ++wc;
              }
          }
          if (fuse_skip) {          // the 1x1 skip conv: the TMA-staged tile of its input is one more K-major A operand (K = N)
            const uint32_t rslot = acc_it % RS;
            tc::mbar_wait(&res_full[rslot], (acc_it / RS) & 1);
            tc::fence_after_sync();
            if (leader) {
#pragma unroll
              for (int ks = 0; ks < N / 16; ++ks) {
                const uint32_t ra = r_lo0 + rslot * (RTILE >> 4) + (uint32_t)(ks * 2 * TH * TW), sb = s_lo0 + (uint32_t)(ks * 2 * LBO_B) / 16;
                tc::mma_bf16_lohi(tmem_d, ra, r_hi, sb, b_hi, IDESC, 1);
                if (SP) {
                  tc::mma_bf16_lohi(tmem_d, ra + (HALF_R >> 4), r_hi, sb, b_hi, IDESC, 1);
                  tc::mma_bf16_lohi(tmem_d, ra, r_hi, sb + (HALF_S >> 4), b_hi, IDESC, 1);
                }
              }
              tc::mma_commit(&res_empty[rslot]);
            }
          }
          if (leader) tc::mma_commit(&acc_full[as]);
          __syncwarp();
        }
        if (leader) {
          tc::mma_commit(&a_empty[slot0]);                                   // depth i is done with slice i
          if (i == dhi - 1 && i + 1 <= din1) tc::mma_commit(&a_empty[slot1]); // item ends: slice i+1 was loaded only for us
        }
        __syncwarp();
      }
      g_base += (uint32_t)(din1 - dlo + 1);
    }
  } else if (warp >= 4) {
    const int e = warp - 4, m = e * 32 + lane, hh = m >> 3, ww = m & 7;
    uint32_t acc_it = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int h = h0 + hh, w = w0 + ww;
      const bool valid = h < p.H && w < p.W;
      for (int i = dlo; i < dhi; ++i)
        for (int ph8 = 0; ph8 < 8; ++ph8, ++acc_it) {
          const int pd = ph8 >> 2, ph = (ph8 >> 1) & 1, pw = ph8 & 1;
          const uint32_t as = acc_it & 1;
          uint4 rpre[N / 32][4];                       // skip-connection chunks of this thread's voxel, from the TMA ring
          const bool use_res = has_res && !fuse_skip;
          if (use_res) {
            const uint32_t slot = acc_it % RS;
            tc::mbar_wait(&res_full[slot], (acc_it / RS) & 1);
            const uint4* rt = reinterpret_cast<const uint4*>(Rbase + slot * RTILE) + m;      // [chunk][TH*TW] uint4
#pragma unroll
            for (int q = 0; q < N / 8; ++q) rpre[q / 4][q % 4] = rt[q * (TH * TW)];
            tc::mbar_arrive(&res_empty[slot]);
          }
          tc::mbar_wait(&acc_full[as], (acc_it >> 1) & 1);
          tc::fence_after_sync();
#pragma unroll
          for (int j = 0; j < N / 32; ++j) {
            float v[32];
            tc::tmem_ld32(tmem_base + ((uint32_t)(e * 32) << 16) + as * N + j * 32, v);
            if (j == N / 32 - 1) {
              tc::fence_before_sync();
              tc::mbar_arrive(&acc_empty[as]);
            }
            const int co0 = nt * N + j * 32;
            if (!valid || co0 >= p.cout_valid) continue;
            epilogue_store32<(SP ? 1 : 0) | EPI>(p, v, s_scale + j * 32, s_shift + j * 32, co0, b, 2 * i + pd, 2 * h + ph, 2 * w + pw, 2 * p.D, 2 * p.H,
                             2 * p.W, use_res ? rpre[j] : nullptr);
          }
        }
    }
  }
  TC_KERNEL_EPILOGUE()
}

// ---- layout converters ----------------------------------------------------------------------------------------------
// fp32 NCDHW -> bf16 blocked [B][C/8][D][H][W][8], or (s2d) phase-split [B][8][C/8][D/2][H/2][W/2][8]
// split_off != 0: hi/lo pair of the bf16x3 split route, lo = bf16(x - hi) written split_off uint4 further ([2][B]... stacking).
// tri != 0 (not with s2d): channel-stacked K-concat form [hi | lo | hi] with 3*C/8 chunks per sample, for 1x1 convs that run the
// three split products as ONE GEMM with K = 3*C against the weights [w_hi | w_hi | w_lo].
__global__ void __launch_bounds__(256) to_blocked_kernel(const float* __restrict__ in, uint4* __restrict__ out, int C, int D, int H,
                                                         int W, int s2d, size_t split_off, int tri) {
  const size_t S = (size_t)D * H * W;
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // voxel within (D,H,W)
  if (v >= S) return;
  const int chunk = blockIdx.y, b = blockIdx.z;
  const float* ip = in + ((size_t)b * C + chunk * 8) * S + v;
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = __ldg(ip + (size_t)i * S);
  uint4 q;
  q.x = tc::pack_bf16x2(f[0], f[1]); q.y = tc::pack_bf16x2(f[2], f[3]);
  q.z = tc::pack_bf16x2(f[4], f[5]); q.w = tc::pack_bf16x2(f[6], f[7]);
  size_t o;
  if (tri) {
    o = ((size_t)b * 3 * (C / 8) + chunk) * S + v;
    out[o + 2 * (size_t)(C / 8) * S] = q;
    split_off = (size_t)(C / 8) * S;
  } else if (!s2d) o = ((size_t)b * (C / 8) + chunk) * S + v;
  else {
    const int x = (int)(v % W), y = (int)((v / W) % H), d = (int)(v / ((size_t)W * H));
    const int phase = ((d & 1) << 2) | ((y & 1) << 1) | (x & 1);
    o = (((size_t)b * 8 + phase) * (C / 8) + chunk) * (S / 8) + ((size_t)(d >> 1) * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1);
  }
  out[o] = q;
  if (split_off) {
    const uint32_t u[4] = {q.x, q.y, q.z, q.w};
    uint4 l;
    l.x = tc::pack_bf16x2(f[0] - __uint_as_float(u[0] << 16), f[1] - __uint_as_float(u[0] & 0xffff0000u));
    l.y = tc::pack_bf16x2(f[2] - __uint_as_float(u[1] << 16), f[3] - __uint_as_float(u[1] & 0xffff0000u));
    l.z = tc::pack_bf16x2(f[4] - __uint_as_float(u[2] << 16), f[5] - __uint_as_float(u[2] & 0xffff0000u));
    l.w = tc::pack_bf16x2(f[6] - __uint_as_float(u[3] << 16), f[7] - __uint_as_float(u[3] & 0xffff0000u));
    out[o + split_off] = l;
  }
}

__global__ void __launch_bounds__(256) from_blocked_kernel(const uint4* __restrict__ in, float* __restrict__ out, int C, size_t S) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= S) return;
  const int chunk = blockIdx.y, b = blockIdx.z;
  const uint4 q = in[((size_t)b * (C / 8) + chunk) * S + v];
  const uint32_t u[4] = {q.x, q.y, q.z, q.w};
  float* op = out + ((size_t)b * C + chunk * 8) * S + v;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    op[(size_t)(2 * i) * S] = __uint_as_float(u[i] << 16);
    op[(size_t)(2 * i + 1) * S] = __uint_as_float(u[i] & 0xffff0000u);
  }
}

// bf16 blocked -> bf16 phase-split blocked (pure permutation of 16-byte voxel chunks)
__global__ void __launch_bounds__(256) blocked_to_s2d_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int C8, int D, int H,
                                                             int W) {
  const size_t S = (size_t)D * H * W;
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= S) return;
  const int chunk = blockIdx.y, b = blockIdx.z;
  const int x = (int)(v % W), y = (int)((v / W) % H), d = (int)(v / ((size_t)W * H));
  const int phase = ((d & 1) << 2) | ((y & 1) << 1) | (x & 1);
  out[(((size_t)b * 8 + phase) * C8 + chunk) * (S / 8) + ((size_t)(d >> 1) * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)] =
      in[((size_t)b * C8 + chunk) * S + v];
}

// bf16 -> fp32, 8 elements per thread (host boundary: features shipped over PCIe as bf16 are widened on the device)
__global__ void __launch_bounds__(256) widen_bf16_kernel(const uint4* __restrict__ in, float4* __restrict__ out, size_t n8) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 q = __ldcs(in + i);
  out[2 * i] = make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u), __uint_as_float(q.y << 16), __uint_as_float(q.y & 0xffff0000u));
  out[2 * i + 1] = make_float4(__uint_as_float(q.z << 16), __uint_as_float(q.z & 0xffff0000u), __uint_as_float(q.w << 16), __uint_as_float(q.w & 0xffff0000u));
}

// dims: (W*8, H, D, outer) ; box (80, 18, 1, box_outer)
int make_act_tmap(CUtensorMap* tm, const void* base, int W, int H, int D, long long outer, int box_outer) {
  ss_encode_tiled_fn enc = ss_get_encode_tiled();
  if (!enc) return SS_ERR_CUDA;
  cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)outer};
  cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
  cuuint32_t box[4] = {(cuuint32_t)WW * 8, (cuuint32_t)HH, 1u, (cuuint32_t)box_outer};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ss_set_error("cuTensorMapEncodeTiled failed with CUresult %d (W=%d H=%d D=%d outer=%lld box_outer=%d)", (int)r, W, H, D, outer, box_outer);
    return SS_ERR_CUDA;
  }
  return SS_OK;
}

// Picks the depth chunk (balances waves against the halo slices every chunk re-reads) and sizes the persistent grid.
void plan_tc(TcP& p, double halo_cost, int& grid) {
  grid = ss_num_sms();
  grid -= grid % p.n_tiles;
  if (grid < p.n_tiles) grid = p.n_tiles;
  const int spatial = p.B * p.HT * p.WT;
  int best = p.D;
  double best_cost = 1e30;
  for (int dc = 1; dc <= p.D; ++dc) {
    if (p.D % dc) continue;
    const long long items = (long long)spatial * (p.D / dc) * p.n_tiles;
    const double cost = (double)ceil_div64(items, grid) * (dc + halo_cost);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = dc; }
  }
  p.DC = best;
  p.n_dc = p.D / best;
  p.items = spatial * p.n_dc;
  const long long total = (long long)p.items * p.n_tiles;
  if (total < grid) grid = (int)total;
}

template <typename K>
int launch_tc(K kernel, size_t smem, const CUtensorMap& tm, TcP p, double halo_cost, cudaStream_t st, const char* name, int threads = 256) {
  SS_CUDA(ss_allow_smem(kernel, smem));
  int grid;
  plan_tc(p, halo_cost, grid);
  kernel<<<grid, threads, smem, st>>>(tm, p);
  SS_CHECK_LAUNCH(name);
  return SS_OK;
}

template <int CIN, int N, int NS, int NWS, int TAPS, int EPI = 0>
int launch_s1(const CUtensorMap& tm, const TcP& p, cudaStream_t st) {
  constexpr size_t smem = (size_t)NS * (CIN / 8) * TILE_B + (size_t)NWS * CIN * N * 2;
  static_assert(smem <= 227 * 1024 - 2048, "shared memory budget");
  return launch_tc(conv3d_tc_s1_kernel<CIN, N, NS, NWS, TAPS, EPI>, smem, tm, p, TAPS == 27 ? 0.35 : 0.0, st, "ss_conv3d_tc(s1)");
}
template <int CIN, int N, int NS, int NWS, bool SP = false, int EPG = 1>
int launch_s1f(const CUtensorMap& tm, const TcP& p, cudaStream_t st) {
  constexpr size_t smem = (SP ? 2 : 1) * ((size_t)NS * (CIN / 8) * TILE_B + (size_t)NWS * CIN * 3 * N * 2);
  static_assert(smem <= 227 * 1024 - 2048, "shared memory budget");
  return launch_tc(conv3d_tc_s1f_kernel<CIN, N, NS, NWS, SP, EPG>, smem, tm, p, 1.4, st, "ss_conv3d_tc(s1f)", 128 + 128 * EPG);
}
template <int CIN, int N, int NS, int NWS, bool SP = false, int EPI = 0>
int launch_s2(const CUtensorMap& tm, const TcP& p, cudaStream_t st) {
  constexpr size_t smem = (SP ? 2 : 1) * ((size_t)NS * 4 * (CIN / 8) * TILE_B + (size_t)NWS * CIN * N * 2);
  static_assert(smem <= 227 * 1024 - 2048, "shared memory budget");
  return launch_tc(conv3d_tc_s2_kernel<CIN, N, NS, NWS, SP, EPI>, smem, tm, p, 0.4, st, "ss_conv3d_tc(s2)");
}
template <int CIN, int N, int NS, int NWS, int RS, bool SP = false, int EPI = 0>
int launch_t2(const CUtensorMap& tm, const TcP& p, cudaStream_t st) {
  constexpr size_t smem = (SP ? 2 : 1) * ((size_t)NS * (CIN / 8) * TILE_B + (size_t)NWS * CIN * N * 2 + (size_t)RS * (N / 8) * TH * TW * 16 + (size_t)N * N * 2);
  static_assert(smem <= 227 * 1024 - 2048, "shared memory budget");
  CUtensorMap tmr = tm;                      // placeholder when there is no residual (never dereferenced then)
  if (p.residual) {                          // phase-split residual: dims (W*8, H, D, B*8*Cout/8), dense TW x TH tiles
    ss_encode_tiled_fn enc = ss_get_encode_tiled();
    if (!enc) return SS_ERR_CUDA;
    cuuint64_t dims[4] = {(cuuint64_t)p.W * 8, (cuuint64_t)p.H, (cuuint64_t)p.D, (cuuint64_t)(SP ? 2 : 1) * p.B * 8 * (p.cout_valid / 8)};
    cuuint64_t strides[3] = {(cuuint64_t)p.W * 16, (cuuint64_t)p.H * p.W * 16, (cuuint64_t)p.D * p.H * p.W * 16};
    cuuint32_t box[4] = {(cuuint32_t)TW * 8, (cuuint32_t)TH, 1u, (cuuint32_t)(N / 8)};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(p.residual), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      ss_set_error("cuTensorMapEncodeTiled(residual) failed with CUresult %d", (int)r);
      return SS_ERR_CUDA;
    }
  }
  auto kernel = conv3d_tc_t2_kernel<CIN, N, NS, NWS, RS, SP, EPI>;
  SS_CUDA(ss_allow_smem(kernel, smem));
  TcP q = p;
  int grid;
  plan_tc(q, 0.3, grid);
  kernel<<<grid, 256, smem, st>>>(tm, tmr, q);
  SS_CHECK_LAUNCH("ss_conv3d_tc(t2)");
  return SS_OK;
}

}  // namespace

ss_encode_tiled_fn ss_get_encode_tiled() {
  static ss_encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<ss_encode_tiled_fn>(sym);
  }
  if (!fn) ss_set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
  return fn;
}

// kind: 0 = Conv3d k3 s1, 1 = Conv3d k1, 2 = Conv3d k3 s2 (phase-split input), 3 = ConvTranspose3d k3 s2 p1 op1,
//       4 = Conv2d k3 s1 p1 on a depth-1 volume (9 in-plane taps),
//       5 = Conv3d k3 s1 with the depth taps folded into N (Cout == N in {32, 64}; weights [9][Cin/8][3N][8]).
// Returns the Cout tile N the kernel uses (weights are packed [ceil(Cout/N)][taps][Cin/8][N][8]); 0 = unsupported.
extern "C" int ss_conv3d_tc_ntile(int kind, int Cin, int Cout) {
  switch (kind) {
    case 0:
      if ((Cin == 32 || Cin == 64) && Cout <= 64 && (Cout <= 32 || Cout == 64)) return 32;
      if (Cin == 128 && Cout == 128) return 128;
      return 0;
    case 1:
      if ((Cin == 32 && Cout == 32) || (Cin == 64 && Cout == 64)) return Cout;
      if (Cin == 128 && (Cout == 128 || Cout == 384)) return 128;      // attention_block: qkv Linear and final 1x1x1 conv
      if ((Cin == 128 && Cout == 64) || (Cin == 64 && Cout == 32)) return Cout;   // channelAtt im_att at 1/4 res (SemStereo.py:93-95)
      return 0;
    case 2:
      if (Cin == 32 && Cout == 64) return 64;
      if (Cin == 64 && Cout == 128) return 128;
      return 0;
    case 3:
      if (Cin == 128 && Cout == 64) return 64;
      if (Cin == 64 && Cout == 32) return 32;
      return 0;
    case 4:                                        // concat_feature (SemStereo.py:221-223): 128 -> 64 -> 32 at 1/4 resolution
      if (Cin == 128 && Cout == 64) return 64;
      if (Cin == 64 && Cout == 32) return 32;
      return 0;
    case 5:
      if ((Cin == 32 || Cin == 64) && (Cout == 32 || Cout == 64)) return Cout;
      return 0;
  }
  return 0;
}

// Layers with an in-kernel bf16x3 split configuration (shared memory holds hi and lo of a slice ring and of the weights).
extern "C" int ss_conv3d_tc_split_supported(int kind, int Cin, int Cout) {
  if (kind == 5) return (Cin == 32 && Cout == 32) || (Cin == 64 && Cout == 64);
  if (kind == 2) return Cin == 32 && Cout == 64;
  if (kind == 3) return Cin == 64 && Cout == 32;
  return 0;
}

// D,H,W are the INPUT dims of the layer (kind 2: of the full-resolution input, all even; the tensor itself is phase-split).
extern "C" int ss_conv3d_tc(int kind, const void* in_blocked, const void* weight_packed, const float* scale_or_null,
                            const float* shift_or_null, const float* gate_blocked_or_null, const void* residual_s2d_or_null,
                            const void* skip_weight_or_null, void* out, int out_mode, int B, int Cin, int Cout, int D, int H, int W, int relu, void* stream) {
  return ss_conv3d_tc_ex(kind, in_blocked, weight_packed, scale_or_null, shift_or_null, gate_blocked_or_null, residual_s2d_or_null,
                         skip_weight_or_null, nullptr, nullptr, out, out_mode, 0, 0, B, Cin, Cout, D, H, W, relu, stream);
}

// The same layers with the two hooks of the bf16x3 split route (see the header): fp32 partial sums added to the accumulator
// before the affine, and the bf16 output written as a hi/lo pair stacked on the batch axis ([2][B]...).
extern "C" int ss_conv3d_tc_ex(int kind, const void* in_blocked, const void* weight_packed, const float* scale_or_null,
                               const float* shift_or_null, const float* gate_blocked_or_null, const void* residual_s2d_or_null,
                               const void* skip_weight_or_null, const float* acc_in_or_null, const float* acc_in2_or_null, void* out,
                               int out_mode, int out_split, int in_split, int B, int Cin, int Cout, int D, int H, int W, int relu,
                               void* stream) {
  SS_REQUIRE(in_blocked && weight_packed && out, "ss_conv3d_tc: null pointer");
  SS_UNSUPPORTED(in_split && !ss_conv3d_tc_split_supported(kind, Cin, Cout),
                 "ss_conv3d_tc: no in-kernel split configuration for kind %d (Cin=%d, Cout=%d); use the two-launch route", kind, Cin, Cout);
  SS_REQUIRE(!in_split || !residual_s2d_or_null || skip_weight_or_null, "ss_conv3d_tc: the split route adds the skip only through the fused skip conv");
  SS_REQUIRE(acc_in_or_null || !acc_in2_or_null, "ss_conv3d_tc: acc_in2 needs acc_in");
  SS_REQUIRE(!out_split || out_mode == 0 || out_mode == 2, "ss_conv3d_tc: a split (hi/lo) output exists only for the bf16 layouts");
  SS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && Cout > 0, "ss_conv3d_tc: non-positive dimension");
  const int N = ss_conv3d_tc_ntile(kind, Cin, Cout);
  SS_UNSUPPORTED(N == 0, "ss_conv3d_tc: kind %d with (Cin=%d, Cout=%d) has no tensor-core configuration", kind, Cin, Cout);
  SS_REQUIRE((reinterpret_cast<uintptr_t>(in_blocked) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(weight_packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(residual_s2d_or_null) & 15) == 0,
             "ss_conv3d_tc: pointers must be 16-byte aligned");
  SS_REQUIRE(out_mode >= 0 && out_mode <= 3, "ss_conv3d_tc: out_mode must be 0, 1, 2 or 3");
  SS_REQUIRE(out_mode == 1 || Cout % 8 == 0, "ss_conv3d_tc: a blocked output needs Cout %% 8 == 0");
  // the fp32 partial-sum hooks (two-launch split route) are compiled only into the kernels of the layers that need them
  const bool ex = acc_in_or_null || out_mode == 3 || (out_split && !in_split);
  SS_UNSUPPORTED(ex && !(!in_split && ((kind == 0 && Cin == 128 && Cout == 128) || (kind == 2 && Cin == 64 && Cout == 128) ||
                                      (kind == 3 && Cin == 128 && Cout == 64))),
                 "ss_conv3d_tc: partial-sum hooks / out_mode 3 exist only for the 128-channel layers (kind %d, Cin=%d, Cout=%d)", kind, Cin, Cout);
  SS_REQUIRE((reinterpret_cast<uintptr_t>(acc_in_or_null) & 15) == 0 && (reinterpret_cast<uintptr_t>(acc_in2_or_null) & 15) == 0,
             "ss_conv3d_tc: partial sums must be 16-byte aligned");
  SS_REQUIRE(!gate_blocked_or_null || Cout % 8 == 0, "ss_conv3d_tc: the blocked gate needs Cout %% 8 == 0");
  SS_REQUIRE(out_mode != 2 || (kind != 3 && ((kind == 2 ? D / 2 : D) % 2 == 0) && ((kind == 2 ? H / 2 : H) % 2 == 0) &&
                               ((kind == 2 ? W / 2 : W) % 2 == 0)),
             "ss_conv3d_tc: phase-split output needs even output dims and is not available for the transposed layer");
  SS_REQUIRE(kind == 3 || !residual_s2d_or_null, "ss_conv3d_tc: the residual input exists only for the transposed layer");
  SS_REQUIRE(!skip_weight_or_null || (residual_s2d_or_null && Cout == N && (reinterpret_cast<uintptr_t>(skip_weight_or_null) & 15) == 0),
             "ss_conv3d_tc: a fused skip conv needs its phase-split input as `residual` and Cout == the layer's tile");
  SS_REQUIRE(kind != 2 || (D % 2 == 0 && H % 2 == 0 && W % 2 == 0), "ss_conv3d_tc: stride-2 layer needs even input dims");
  SS_REQUIRE(kind != 4 || D == 1, "ss_conv3d_tc: the 2-D layer (kind 4) takes a depth-1 volume");
  SS_REQUIRE(kind >= 0 && kind <= 5, "ss_conv3d_tc: unknown kind %d", kind);
  TcP p;
  p.w = reinterpret_cast<const __nv_bfloat16*>(weight_packed);
  p.scale = scale_or_null; p.shift = shift_or_null; p.gate = gate_blocked_or_null;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual_s2d_or_null);
  p.skip_w = reinterpret_cast<const __nv_bfloat16*>(skip_weight_or_null);
  p.out = out; p.out_mode = out_mode; p.cout_valid = Cout;
  p.acc_in = acc_in_or_null; p.acc_in2 = acc_in2_or_null;
  {   // lo half = the same tensor one batch-stack further: B * (Cout/8) * output voxels 16-byte chunks (both bf16 layouts)
    const long long osz = (long long)(kind == 3 ? 8 : 1) * (kind == 2 ? D / 2 : D) * (kind == 2 ? H / 2 : H) * (kind == 2 ? W / 2 : W);
    p.split_off = out_split ? (long long)B * (Cout / 8) * osz : 0;
  }
  p.B = B; p.relu = relu;
  p.n_tiles = ceil_div(Cout, N);
  p.DC = 1; p.n_dc = 1; p.items = 0;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tm;
  int rc;
  if (kind == 2) {
    p.D = D / 2; p.H = H / 2; p.W = W / 2;
    rc = make_act_tmap(&tm, in_blocked, p.W, p.H, p.D, (long long)(in_split ? 2 : 1) * B * 8 * (Cin / 8), 4 * (Cin / 8));
  } else {
    p.D = D; p.H = H; p.W = W;
    rc = make_act_tmap(&tm, in_blocked, W, H, D, (long long)(in_split ? 2 : 1) * B * (Cin / 8), Cin / 8);
  }
  if (rc != SS_OK) return rc;
  p.HT = ceil_div(p.H, TH); p.WT = ceil_div(p.W, TW);
  if (in_split) {
    if (kind == 5) return Cin == 32 ? launch_s1f<32, 32, 4, 9, true, 2>(tm, p, st) : launch_s1f<64, 64, 2, 2, true, 2>(tm, p, st);
    if (kind == 2) return launch_s2<32, 64, 2, 4, true>(tm, p, st);
    return launch_t2<64, 32, 3, 4, 2, true>(tm, p, st);
  }
  switch (kind) {
    case 0:
      if (Cin == 32) return launch_s1<32, 32, 8, 27, 27>(tm, p, st);
      if (Cin == 64) return launch_s1<64, 32, 5, 27, 27>(tm, p, st);
      return ex ? launch_s1<128, 128, 3, 2, 27, 3>(tm, p, st) : launch_s1<128, 128, 3, 2, 27>(tm, p, st);
    case 1:
      if (Cin == 32) return launch_s1<32, 32, 4, 1, 1>(tm, p, st);
      if (Cin == 64) return N == 64 ? launch_s1<64, 64, 4, 1, 1>(tm, p, st) : launch_s1<64, 32, 4, 1, 1>(tm, p, st);
      return N == 128 ? launch_s1<128, 128, 4, 1, 1>(tm, p, st) : launch_s1<128, 64, 4, 1, 1>(tm, p, st);
    case 2:
      if (Cin == 32) return launch_s2<32, 64, 2, 27>(tm, p, st);
      return ex ? launch_s2<64, 128, 2, 2, false, 3>(tm, p, st) : launch_s2<64, 128, 2, 2>(tm, p, st);
    case 4:
      if (Cin == 128) return launch_s1<128, 64, 2, 8, 9>(tm, p, st);      // 8-deep tap ring: a 2-deep one exposes the 16 KB reload latency
      return launch_s1<64, 32, 4, 9, 9>(tm, p, st);
    case 5:
      if (Cin == 32 && Cout == 32) return launch_s1f<32, 32, 6, 9, false, 2>(tm, p, st);
      if (Cin == 64 && Cout == 32) return launch_s1f<64, 32, 4, 9, false, 2>(tm, p, st);
      if (Cin == 32 && Cout == 64) return launch_s1f<32, 64, 6, 9, false, 2>(tm, p, st);
      return launch_s1f<64, 64, 2, 7, false, 2>(tm, p, st);
    default:
      if (Cin == 128) return ex ? launch_t2<128, 64, 2, 5, 2, false, 3>(tm, p, st) : launch_t2<128, 64, 2, 5, 2>(tm, p, st);
      return launch_t2<64, 32, 3, 27, 4>(tm, p, st);
  }
}

// concat_stem with the sparse concat volume generated inside the kernel (k9 kernel above).
extern "C" int ss_concat_stem_fused(const void* cf_l_blocked, const void* cf_r_blocked, const float* disp_topk, const float* att_topk,
                                    const void* weight_packed, const float* scale_or_null, const float* shift_or_null,
                                    const float* gate_blocked_or_null, void* out, int out_mode, int B, int K, int H, int W, int dmin,
                                    int relu, void* stream) {
  SS_REQUIRE(cf_l_blocked && cf_r_blocked && disp_topk && att_topk && weight_packed && out, "ss_concat_stem_fused: null pointer");
  SS_REQUIRE(B > 0 && K > 0 && H > 0 && W > 0, "ss_concat_stem_fused: non-positive dimension");
  SS_REQUIRE(out_mode >= 0 && out_mode <= 2, "ss_concat_stem_fused: out_mode must be 0, 1 or 2");
  SS_REQUIRE(out_mode != 2 || (K % 2 == 0 && H % 2 == 0 && W % 2 == 0), "ss_concat_stem_fused: phase-split output needs even dims");
  SS_REQUIRE(((reinterpret_cast<uintptr_t>(cf_l_blocked) | reinterpret_cast<uintptr_t>(cf_r_blocked) | reinterpret_cast<uintptr_t>(disp_topk) |
               reinterpret_cast<uintptr_t>(att_topk) | reinterpret_cast<uintptr_t>(weight_packed) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
             "ss_concat_stem_fused: pointers must be 16-byte aligned");
  ss_encode_tiled_fn enc = ss_get_encode_tiled();
  if (!enc) return SS_ERR_CUDA;
  K9P kp;
  TcP& p = kp.t;
  p.w = reinterpret_cast<const __nv_bfloat16*>(weight_packed);
  p.scale = scale_or_null; p.shift = shift_or_null; p.gate = gate_blocked_or_null; p.residual = nullptr; p.skip_w = nullptr;
  p.out = out; p.out_mode = out_mode; p.cout_valid = 32;
  p.acc_in = nullptr; p.acc_in2 = nullptr; p.split_off = 0;
  p.B = B; p.D = K; p.H = H; p.W = W; p.relu = relu;
  p.n_tiles = 1; p.HT = ceil_div(H, TH); p.WT = ceil_div(W, TW);
  p.DC = 1; p.n_dc = 1; p.items = 0;
  kp.dmin = dmin; kp.disp = disp_topk; kp.att = att_topk; kp.cf_l = reinterpret_cast<const uint4*>(cf_l_blocked);
  CUtensorMap tmR;
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  auto encode = [&](CUtensorMap* tm, CUtensorMapDataType dt, const void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                    const cuuint32_t* box) -> int {
    CUresult r = enc(tm, dt, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      ss_set_error("ss_concat_stem_fused: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
      return SS_ERR_CUDA;
    }
    return SS_OK;
  };
  {
    // (8 channels, x, y, chunk): a box dimension may not exceed 256 elements, so the 41-pixel window cannot be 328 bf16 in one dim
    const cuuint64_t dims[4] = {8u, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * 4};
    const cuuint64_t strides[3] = {16u, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
    const cuuint32_t boxr[4] = {8u, (cuuint32_t)K9_RW, (cuuint32_t)HH, 4u};
    int rc = encode(&tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, cf_r_blocked, dims, strides, boxr);
    if (rc != SS_OK) return rc;
  }
  SS_CUDA(ss_allow_smem(concat_stem_k9_kernel, K9_SMEM));
  int grid;
  plan_tc(p, 1.4, grid);
  concat_stem_k9_kernel<<<grid, 384, K9_SMEM, (cudaStream_t)stream>>>(tmR, kp);
  SS_CHECK_LAUNCH("ss_concat_stem_fused");
  return SS_OK;
}

extern "C" int ss_to_blocked_bf16(const float* in_ncdhw, void* out_blocked, int B, int C, int D, int H, int W, int s2d, void* stream) {
  return ss_to_blocked_bf16_ex(in_ncdhw, out_blocked, B, C, D, H, W, s2d, 0, stream);
}

extern "C" int ss_to_blocked_bf16_ex(const float* in_ncdhw, void* out_blocked, int B, int C, int D, int H, int W, int s2d, int split,
                                     void* stream) {
  SS_REQUIRE(in_ncdhw && out_blocked && B > 0 && C > 0 && D > 0 && H > 0 && W > 0, "ss_to_blocked_bf16: bad argument");
  SS_REQUIRE(C % 8 == 0, "ss_to_blocked_bf16: C=%d must be a multiple of 8", C);
  SS_REQUIRE(!s2d || (D % 2 == 0 && H % 2 == 0 && W % 2 == 0), "ss_to_blocked_bf16: phase-split layout needs even dims");
  SS_REQUIRE(split >= 0 && split <= 2 && !(split == 2 && s2d), "ss_to_blocked_bf16: split must be 0, 1 or 2 (2 not with s2d)");
  SS_UNSUPPORTED(C / 8 > 65535 || B > 65535, "ss_to_blocked_bf16: grid dimension exceeds 65535");
  const size_t S = (size_t)D * H * W;
  to_blocked_kernel<<<dim3((unsigned)ceil_div64(S, 256), C / 8, B), 256, 0, (cudaStream_t)stream>>>(
      in_ncdhw, reinterpret_cast<uint4*>(out_blocked), C, D, H, W, s2d, split == 1 ? (size_t)B * (C / 8) * S : (size_t)0, split == 2);
  SS_CHECK_LAUNCH("ss_to_blocked_bf16");
  return SS_OK;
}

extern "C" int ss_widen_bf16(const void* in_bf16, float* out_f32, long long n, void* stream) {
  SS_REQUIRE(in_bf16 && out_f32 && n > 0, "ss_widen_bf16: bad argument");
  SS_REQUIRE(n % 8 == 0 && ((reinterpret_cast<uintptr_t>(in_bf16) | reinterpret_cast<uintptr_t>(out_f32)) & 15) == 0,
             "ss_widen_bf16: element count must be a multiple of 8 and the pointers 16-byte aligned");
  const size_t n8 = (size_t)n / 8;
  SS_UNSUPPORTED(ceil_div64(n8, 256) > 0x7fffffffLL, "ss_widen_bf16: too many elements");
  widen_bf16_kernel<<<(unsigned)ceil_div64(n8, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(in_bf16),
                                                                                    reinterpret_cast<float4*>(out_f32), n8);
  SS_CHECK_LAUNCH("ss_widen_bf16");
  return SS_OK;
}

extern "C" int ss_from_blocked_bf16(const void* in_blocked, float* out_ncdhw, int B, int C, int D, int H, int W, void* stream) {
  SS_REQUIRE(in_blocked && out_ncdhw && B > 0 && C > 0 && D > 0 && H > 0 && W > 0, "ss_from_blocked_bf16: bad argument");
  SS_REQUIRE(C % 8 == 0, "ss_from_blocked_bf16: C=%d must be a multiple of 8", C);
  SS_UNSUPPORTED(C / 8 > 65535 || B > 65535, "ss_from_blocked_bf16: grid dimension exceeds 65535");
  const size_t S = (size_t)D * H * W;
  from_blocked_kernel<<<dim3((unsigned)ceil_div64(S, 256), C / 8, B), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(in_blocked), out_ncdhw, C, S);
  SS_CHECK_LAUNCH("ss_from_blocked_bf16");
  return SS_OK;
}

extern "C" int ss_blocked_to_s2d(const void* in_blocked, void* out_s2d, int B, int C, int D, int H, int W, void* stream) {
  SS_REQUIRE(in_blocked && out_s2d && B > 0 && C > 0 && D > 0 && H > 0 && W > 0, "ss_blocked_to_s2d: bad argument");
  SS_REQUIRE(C % 8 == 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "ss_blocked_to_s2d: C %% 8 == 0 and even dims required");
  SS_UNSUPPORTED(C / 8 > 65535 || B > 65535, "ss_blocked_to_s2d: grid dimension exceeds 65535");
  const size_t S = (size_t)D * H * W;
  blocked_to_s2d_kernel<<<dim3((unsigned)ceil_div64(S, 256), C / 8, B), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(in_blocked), reinterpret_cast<uint4*>(out_s2d), C / 8, D, H, W);
  SS_CHECK_LAUNCH("ss_blocked_to_s2d");
  return SS_OK;
}
