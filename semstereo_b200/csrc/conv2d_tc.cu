// 2-D decoder convolutions around the disparity path as tcgen05 implicit GEMMs:
//   conv   : Conv2d 3x3 s1 p1 / Conv2d 1x1 over a (virtual) channel concat of up to two inputs  -- Conv2x.conv2, segmenthead.conv1/2,
//            chal_* (models/submodule.py:31-52, 119-161; models/SemStereo.py:207-216)
//   deconv : ConvTranspose2d k4 s2 p1 as 4 sub-pixel output phases                               -- Conv2x.conv1 (deconv=True), spx2
// Layout: bf16 blocked [B][C/8][H][W][8] (the depth-1 case of the 3-D layout in conv3d_tc.cu), same 18x10 halo tile per 8-channel
// chunk (TMA box, zero fill outside the image = the conv padding), so every in-plane tap is a descriptor start address.
// GEMM: M = 128 pixels (16 h x 8 w), N = Cout tile, K = Cin * taps walked in 64-channel blocks; the K loop crosses from input 0
// to input 1 at a block boundary, which is how torch.cat((x, rem), 1) (submodule.py:155) never materialises.
// deconv: out[2m+ph, 2n+pw] = sum over shifts (sh,sw) of x[m+sh, n+sw] * w[k(ph,sh), k(pw,sw)] with k(0,0)=1, k(0,-1)=3, k(1,0)=2,
// k(1,+1)=0.  Shift (0,0) feeds all four phases, edge shifts two, corner shifts one: the four phase accumulators sit side by
// side in TMEM and one MMA of N = 4*NP / 2*NP / NP covers every phase a shifted tile feeds (11 MMAs instead of 16 per K step).
// Skeleton as in conv3d_tc.cu: persistent CTAs, warp 0 TMA tile producer, warp 3 weight-slab producer (bulk copies through a
// ring), warp 1 MMA issuer (warp-uniform control, one elected lane), warp 2 TMEM allocator, warps 4-7 epilogue
// (y = acc*scale + shift -> ReLU -> bf16 blocked or fp32 NCHW); accumulators double-buffered in TMEM.
#include "tc_common.cuh"

namespace {

constexpr int TH = 16, TW = 8, HH = TH + 2, WW = TW + 2;
constexpr uint32_t TILE_B = HH * WW * 16;       // one 8-channel chunk of a halo tile
constexpr int CB = 64;                          // channels per K block
constexpr uint32_t SLICE = (CB / 8) * TILE_B;   // 23040 B

struct C2P {
  const __nv_bfloat16* w;
  const float* scale;   // [Cout] or null (1)
  const float* shift;   // [Cout] or null (0)
  const __nv_bfloat16* residual;   // 1x1 mode only: bf16 blocked tensor shaped like the output, added after the activation, or null
  void* out;
  int out_mode;         // 0: bf16 blocked, 1: fp32 NCHW
  int act;              // 1x1 mode only (else `relu`): 0 none, 1 ReLU, 2 SiLU
  int cout;             // channels stored
  int B, H, W;          // input dims (conv: also output dims; deconv writes 2H x 2W)
  int relu;
  int ncb0, ncb;        // 64-channel blocks of input 0 / of both inputs
  int n_tiles, HT, WT, items;   // items per n-tile = B*HT*WT
};

__device__ __forceinline__ void decode(const C2P& p, int s, int& nt, int& b, int& h0, int& w0) {
  const int wt = s % p.WT;  s /= p.WT;
  const int ht = s % p.HT;  s /= p.HT;
  b = s % p.B;
  nt = s / p.B;
  h0 = ht * TH; w0 = wt * TW;
}

__device__ __forceinline__ void affine_relu32(float (&v)[32], const float* sc, const float* sh, bool relu) {
  const float lo = relu ? 0.0f : -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = fmaxf(fmaf(v[i], sc[i], sh[i]), lo);
}

__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 q;
  q.x = tc::pack_bf16x2(v[0], v[1]); q.y = tc::pack_bf16x2(v[2], v[3]);
  q.z = tc::pack_bf16x2(v[4], v[5]); q.w = tc::pack_bf16x2(v[6], v[7]);
  return q;
}

#define C2_PROLOGUE(NSLOTS, NWSLOTS, NSTAGE, TMEM_COLS)                                                                    \
  extern __shared__ __align__(1024) uint8_t smem[];                                                                        \
  __shared__ __align__(8) uint64_t a_full[NSLOTS], a_empty[NSLOTS], w_full[NWSLOTS], w_empty[NWSLOTS], acc_full[4], acc_empty[4]; \
  __shared__ uint32_t tmem_base_s;                                                                                         \
  __shared__ float s_scale[NSTAGE], s_shift[NSTAGE];                                                                       \
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;                                                              \
  if (threadIdx.x == 0) {                                                                                                  \
    tc::prefetch_tmap(&tm0);                                                                                               \
    tc::prefetch_tmap(&tm1);                                                                                               \
    for (int i = 0; i < NSLOTS; ++i) { tc::mbar_init(&a_full[i], 1); tc::mbar_init(&a_empty[i], 1); }                      \
    for (int i = 0; i < NWSLOTS; ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }                     \
    for (int i = 0; i < 4; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 128); }                     \
    tc::fence_barrier_init();                                                                                              \
  }                                                                                                                        \
  if (warp == 2) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);                                                                  \
  tc::fence_before_sync();                                                                                                 \
  __syncthreads();                                                                                                         \
  tc::fence_after_sync();                                                                                                  \
  const uint32_t tmem_base = tmem_base_s;

// The tile producer is the same for both kernels: one 64-channel halo tile per K block, from input 0 or input 1.
#define C2_TILE_PRODUCER(NSLOTS)                                                                                           \
  uint32_t g = 0;                                                                                                          \
  for (int s = blockIdx.x; s < total; s += gridDim.x) {                                                                    \
    int nt, b, h0, w0;                                                                                                     \
    decode(p, s, nt, b, h0, w0);                                                                                           \
    for (int cb = 0; cb < p.ncb; ++cb, ++g) {                                                                              \
      const uint32_t slot = g % NSLOTS;                                                                                    \
      tc::mbar_wait(&a_empty[slot], ((g / NSLOTS) & 1) ^ 1);                                                               \
      tc::mbar_expect_tx(&a_full[slot], SLICE);                                                                            \
      if (cb < p.ncb0) tc::tma_load_4d(Abase + slot * SLICE, &tm0, &a_full[slot], (w0 - 1) * 8, h0 - 1, 0, (b * p.ncb0 + cb) * (CB / 8)); \
      else tc::tma_load_4d(Abase + slot * SLICE, &tm1, &a_full[slot], (w0 - 1) * 8, h0 - 1, 0,                             \
                           (b * (p.ncb - p.ncb0) + cb - p.ncb0) * (CB / 8));                                               \
    }                                                                                                                      \
  }

// =====================================================================================================================
// conv: TAPS = 9 (3x3, pad 1) or 1 (1x1).  Weights: [n_tiles][ncb][TAPS][CB/8][N][8].
// =====================================================================================================================
// EXT (1x1 mode; compile time so that the 3x3 decoder kernels keep their registers): SiLU activation and a residual input in
// the epilogue -- the inverted-residual / transformer blocks of the MobileViTv2 backbone.
// EPG: epilogue warp groups.  A 1x1 conv with few input channels does 4-16 MMAs per 128-pixel tile but a 128-channel affine + SiLU
// + pack + store epilogue: with one group of 4 warps the epilogue, not the tensor core or HBM, bounds the layer (the backbone's 512^2
// expand convs ran at 1/3 of their HBM time).  With EPG = 2 each group owns one of the two accumulator stages.
template <int N, int TAPS, int NS, int NWS, bool EXT = false, int EPG = 1>
__global__ void __launch_bounds__(128 + 128 * EPG, 1) conv2d_tc_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                                                           const C2P p) {
  constexpr uint32_t TAPB = CB * N * 2;
  constexpr int KS = CB / 16;
  constexpr uint32_t LBO_A = TILE_B, SBO_A = WW * 16, LBO_B = N * 16, SBO_B = 128;
  constexpr uint32_t ACC = N < 32 ? 32 : N;            // accumulator stride in TMEM columns
  constexpr uint32_t IDESC = tc::make_idesc_bf16(128, N);
  constexpr int NST = (N < 32 ? 32 : N);
  constexpr uint32_t NACC = EPG > 1 ? EPG : 2;         // accumulator stages in TMEM: one per epilogue group
  constexpr uint32_t TCOLS = NACC * ACC <= 32 ? 32 : NACC * ACC <= 64 ? 64 : NACC * ACC <= 128 ? 128 : NACC * ACC <= 256 ? 256 : 512;
  static_assert(NACC * ACC <= 512 && NACC <= 4, "TMEM columns");
  C2_PROLOGUE(NS, NWS, EPG * NST, TCOLS)
  uint8_t* Abase = smem;
  uint8_t* Wbase = smem + NS * SLICE;
  const int total = p.items * p.n_tiles;

  if (warp == 0 && lane == 0) {
    C2_TILE_PRODUCER(NS)
  } else if (warp == 3 && lane == 0) {
    // 1x1 convs whose K blocks all fit in the slab ring keep their weights RESIDENT while the Cout tile does not change (the backbone's
    // 512^2 layers do 4 MMAs per tile: re-streaming 16 KB per tile through the ring cost 3x the layer's HBM time, round 2)
    const bool resident = TAPS == 1 && p.ncb <= NWS;
    uint32_t wc = 0, reloads = 0;
    int nt_loaded = -1;
    for (int s = blockIdx.x; s < total; s += gridDim.x) {
      int nt, b, h0, w0;
      decode(p, s, nt, b, h0, w0);
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) + (size_t)nt * p.ncb * TAPS * TAPB;
      if (resident) {
        if (nt == nt_loaded) continue;
        for (int i = 0; i < p.ncb; ++i) {
          tc::mbar_wait(&w_empty[i], (reloads & 1) ^ 1);
          tc::mbar_expect_tx(&w_full[i], TAPB);
          tc::bulk_load(Wbase + i * TAPB, wsrc + (size_t)i * TAPB, TAPB, &w_full[i]);
        }
        nt_loaded = nt;
        ++reloads;
        continue;
      }
      for (int i = 0; i < p.ncb * TAPS; ++i, ++wc) {
        const uint32_t slot = wc % NWS;
        tc::mbar_wait(&w_empty[slot], ((wc / NWS) & 1) ^ 1);
        tc::mbar_expect_tx(&w_full[slot], TAPB);
        tc::bulk_load(Wbase + slot * TAPB, wsrc + (size_t)i * TAPB, TAPB, &w_full[slot]);
      }
    }
  } else if (warp == 1) {
    const bool leader = tc::elect_one();
    const bool resident = TAPS == 1 && p.ncb <= NWS;
    const uint32_t a_lo0 = tc::desc_lo(tc::smem_u32(Abase), LBO_A), a_hi = tc::desc_hi(SBO_A);
    const uint32_t b_lo0 = tc::desc_lo(tc::smem_u32(Wbase), LBO_B), b_hi = tc::desc_hi(SBO_B);
    uint32_t g = 0, wc = 0, acc_it = 0, reloads = 0;
    int nt_loaded = -1;
    for (int s = blockIdx.x; s < total; s += gridDim.x, ++acc_it) {
      int nt = 0, nt_next = -1;
      if (resident) {                        // nt is the slowest index of s: it changes at most n_tiles - 1 times per CTA
        nt = s / p.items;
        nt_next = s + (int)gridDim.x < total ? (s + (int)gridDim.x) / p.items : -1;
        if (nt != nt_loaded) {
          for (int i = 0; i < p.ncb; ++i) tc::mbar_wait(&w_full[i], reloads & 1);
          tc::fence_after_sync();
          nt_loaded = nt;
          ++reloads;
        }
      }
      const uint32_t as = acc_it % NACC;
      tc::mbar_wait(&acc_empty[as], ((acc_it / NACC) & 1) ^ 1);
      tc::fence_after_sync();
      const uint32_t tmem_d = tmem_base + as * ACC;
      uint32_t accumulate = 0;
#pragma unroll 1
      for (int cb = 0; cb < p.ncb; ++cb, ++g) {
        const uint32_t slot = g % NS;
        tc::mbar_wait(&a_full[slot], (g / NS) & 1);
        tc::fence_after_sync();
        const uint32_t a_lo = a_lo0 + slot * (SLICE >> 4);
#pragma unroll
        for (int t = 0; t < TAPS; ++t, ++wc) {
          const int kh = TAPS == 9 ? t / 3 : 1, kw = TAPS == 9 ? t % 3 : 1;
          const uint32_t wslot = resident ? (uint32_t)cb : wc % NWS;
          if (!resident) {
            tc::mbar_wait(&w_full[wslot], (wc / NWS) & 1);
            tc::fence_after_sync();
          }
          const uint32_t b_lo = b_lo0 + wslot * (TAPB >> 4);
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              tc::mma_bf16_lohi(tmem_d, a_lo + (uint32_t)((kh * WW + kw) * 16 + ks * 2 * LBO_A) / 16, a_hi,
                                b_lo + (uint32_t)(ks * 2 * LBO_B) / 16, b_hi, IDESC, accumulate);
              accumulate = 1;
            }
            if (!resident || nt_next != nt) tc::mma_commit(&w_empty[wslot]);      // resident: released after the tile's last item
          }
          accumulate = 1;
        }
        if (leader) tc::mma_commit(&a_empty[slot]);
        __syncwarp();
      }
      if (leader) tc::mma_commit(&acc_full[as]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int e = warp & 3, eg = (warp - 4) >> 2, m = e * 32 + lane, hh = m >> 3, ww = m & 7;
    float* const s_scale_g = s_scale + eg * NST;       // each epilogue group stages its own copy (named barrier 1 + group)
    float* const s_shift_g = s_shift + eg * NST;
    uint32_t acc_it = 0;
    int nt_staged = -1;
    for (int s = blockIdx.x; s < total; s += gridDim.x, ++acc_it) {
      if (EPG > 1 && (int)(acc_it % NACC) != eg) continue;   // group g drains accumulator stage g
      int nt, b, h0, w0;
      decode(p, s, nt, b, h0, w0);
      if (nt != nt_staged) {                 // folded-BN constants of this Cout tile
        asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
        for (int i = (threadIdx.x & 127); i < N; i += 128) {
          const int co = nt * N + i;
          s_scale_g[i] = (p.scale && co < p.cout) ? __ldg(p.scale + co) : 1.0f;
          s_shift_g[i] = (p.shift && co < p.cout) ? __ldg(p.shift + co) : 0.0f;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
        nt_staged = nt;
      }
      const int h = h0 + hh, w = w0 + ww;
      const bool valid = h < p.H && w < p.W;
      const uint32_t as = acc_it % NACC;
      tc::mbar_wait(&acc_full[as], (acc_it / NACC) & 1);
      tc::fence_after_sync();
      constexpr int NJ = (N + 31) / 32;
#pragma unroll 1
      for (int j = 0; j < NJ; ++j) {
        float v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(e * 32) << 16) + as * ACC + j * 32, v);
        if (j == NJ - 1) {
          tc::fence_before_sync();
          tc::mbar_arrive(&acc_empty[as]);
        }
        const int co0 = nt * N + j * 32;
        if (!valid || co0 >= p.cout) continue;
        const size_t HW = (size_t)p.H * p.W, sp = (size_t)h * p.W + w;
        if (!EXT) affine_relu32(v, s_scale_g + (N < 32 ? 0 : j * 32), s_shift_g + (N < 32 ? 0 : j * 32), p.relu);
        else {
          affine_relu32(v, s_scale_g + (N < 32 ? 0 : j * 32), s_shift_g + (N < 32 ? 0 : j * 32), p.act == 1);
          if (p.act == 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __fdividef(v[i], 1.0f + __expf(-v[i]));
          }
          if (p.residual) {
            const uint4* r = reinterpret_cast<const uint4*>(p.residual) + ((size_t)b * (p.cout / 8) + co0 / 8) * HW + sp;
#pragma unroll
            for (int c8 = 0; c8 < (N < 32 ? N / 8 : 4); ++c8)
              if (co0 + 8 * c8 < p.cout) {
                const uint4 q = __ldg(r + (size_t)c8 * HW);
                const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  v[8 * c8 + 2 * i] += __uint_as_float(u[i] << 16);
                  v[8 * c8 + 2 * i + 1] += __uint_as_float(u[i] & 0xffff0000u);
                }
              }
          }
        }
        if (p.out_mode == 1) {
          float* o = reinterpret_cast<float*>(p.out) + ((size_t)b * p.cout + co0) * HW + sp;
#pragma unroll
          for (int i = 0; i < (N < 32 ? N : 32); ++i)
            if (co0 + i < p.cout) o[(size_t)i * HW] = v[i];
        } else {
          uint4* o = reinterpret_cast<uint4*>(p.out) + ((size_t)b * (p.cout / 8) + co0 / 8) * HW + sp;
#pragma unroll
          for (int c8 = 0; c8 < (N < 32 ? N / 8 : 4); ++c8)
            if (co0 + 8 * c8 < p.cout) o[(size_t)c8 * HW] = pack8(v + 8 * c8);
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, TCOLS);
}

// =====================================================================================================================
// deconv: ConvTranspose2d k4 s2 p1.  NP = Cout tile per output phase.  Weight slabs per (n-tile, K block), in this order:
//   shift ( 0, 0): phases 0,1,2,3   (rows 4*NP)      phase = 2*ph + pw
//   shift (-1, 0): phases 0,1       shift (+1, 0): phases 2,3
//   shift ( 0,-1): phases 0,2       shift ( 0,+1): phases 1,3
//   shift (-1,-1): 0    (-1,+1): 1    (+1,-1): 2    (+1,+1): 3
// each slab [CB/8][rows][8]; 16*NP rows per K block in total.
// =====================================================================================================================
__host__ __device__ constexpr int slab_rows(int s) { return s == 0 ? 4 : s <= 4 ? 2 : 1; }          // in units of NP
__host__ __device__ constexpr int slab_row_off(int s) {                                               // cumulative, units of NP
  return s == 0 ? 0 : s == 1 ? 4 : s == 2 ? 6 : s == 3 ? 8 : s == 4 ? 10 : 12 + (s - 5);
}
__host__ __device__ constexpr int slab_sh(int s) { return s == 1 || s == 5 || s == 6 ? -1 : (s == 2 || s == 7 || s == 8 ? 1 : 0); }
__host__ __device__ constexpr int slab_sw(int s) { return s == 3 || s == 5 || s == 7 ? -1 : (s == 4 || s == 6 || s == 8 ? 1 : 0); }

template <int NP, int NS, int NWS>
__global__ void __launch_bounds__(256, 1) deconv2d_tc_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                                                             const C2P p) {
  constexpr uint32_t SLOTB = CB * 4 * NP * 2;            // ring slot = the largest slab
  constexpr int KS = CB / 16;
  constexpr uint32_t LBO_A = TILE_B, SBO_A = WW * 16, SBO_B = 128;
  constexpr uint32_t ACC = 4 * NP < 32 ? 32 : 4 * NP;
  C2_PROLOGUE(NS, NWS, (NP < 32 ? 32 : NP), 2 * ACC)
  uint8_t* Abase = smem;
  uint8_t* Wbase = smem + NS * SLICE;
  const int total = p.items * p.n_tiles;

  if (warp == 0 && lane == 0) {
    C2_TILE_PRODUCER(NS)
  } else if (warp == 3 && lane == 0) {
    uint32_t wc = 0;
    for (int s = blockIdx.x; s < total; s += gridDim.x) {
      int nt, b, h0, w0;
      decode(p, s, nt, b, h0, w0);
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) + (size_t)nt * p.ncb * (16 * NP * CB * 2);
      for (int cb = 0; cb < p.ncb; ++cb)
#pragma unroll
        for (int sl = 0; sl < 9; ++sl, ++wc) {
          const uint32_t slot = wc % NWS, bytes = (uint32_t)slab_rows(sl) * NP * CB * 2;
          tc::mbar_wait(&w_empty[slot], ((wc / NWS) & 1) ^ 1);
          tc::mbar_expect_tx(&w_full[slot], bytes);
          tc::bulk_load(Wbase + slot * SLOTB, wsrc + ((size_t)cb * 16 + slab_row_off(sl)) * NP * CB * 2, bytes, &w_full[slot]);
        }
    }
  } else if (warp == 1) {
    const bool leader = tc::elect_one();
    const uint32_t a_lo0 = tc::desc_lo(tc::smem_u32(Abase), LBO_A), a_hi = tc::desc_hi(SBO_A);
    const uint32_t b_hi = tc::desc_hi(SBO_B), w_addr = tc::smem_u32(Wbase);
    uint32_t g = 0, wc = 0, acc_it = 0;
    for (int s = blockIdx.x; s < total; s += gridDim.x, ++acc_it) {
      const uint32_t as = acc_it & 1;
      tc::mbar_wait(&acc_empty[as], ((acc_it >> 1) & 1) ^ 1);
      tc::fence_after_sync();
      const uint32_t tmem_d = tmem_base + as * ACC;
      uint32_t accumulate = 0;
#pragma unroll 1
      for (int cb = 0; cb < p.ncb; ++cb, ++g) {
        const uint32_t slot = g % NS;
        tc::mbar_wait(&a_full[slot], (g / NS) & 1);
        tc::fence_after_sync();
        const uint32_t a_lo = a_lo0 + slot * (SLICE >> 4);
#pragma unroll
        for (int sl = 0; sl < 9; ++sl, ++wc) {
          const int sh = slab_sh(sl), sw = slab_sw(sl), rows = slab_rows(sl) * NP;
          const uint32_t wslot = wc % NWS;
          tc::mbar_wait(&w_full[wslot], (wc / NWS) & 1);
          tc::fence_after_sync();
          const uint32_t b_lo = tc::desc_lo(w_addr + wslot * SLOTB, (uint32_t)rows * 16);
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              const uint32_t a = a_lo + (uint32_t)(((1 + sh) * WW + 1 + sw) * 16 + ks * 2 * LBO_A) / 16;
              const uint32_t bb = b_lo + (uint32_t)(ks * 2 * rows * 16) / 16;
              if (sl == 0) {
                tc::mma_bf16_lohi(tmem_d, a, a_hi, bb, b_hi, tc::make_idesc_bf16(128, 4 * NP), accumulate);
                accumulate = 1;
              } else if (sl == 1 || sl == 2) {
                tc::mma_bf16_lohi(tmem_d + (sl == 2 ? 2 * NP : 0), a, a_hi, bb, b_hi, tc::make_idesc_bf16(128, 2 * NP), 1u);
              } else if (sl == 3 || sl == 4) {
                const uint32_t c0 = sl == 4 ? NP : 0;
                tc::mma_bf16_lohi(tmem_d + c0, a, a_hi, bb, b_hi, tc::make_idesc_bf16(128, NP), 1u);
                tc::mma_bf16_lohi(tmem_d + c0 + 2 * NP, a, a_hi, bb + (NP / 8) * (SBO_B >> 4), b_hi, tc::make_idesc_bf16(128, NP), 1u);
              } else {
                tc::mma_bf16_lohi(tmem_d + (sl - 5) * NP, a, a_hi, bb, b_hi, tc::make_idesc_bf16(128, NP), 1u);
              }
            }
            tc::mma_commit(&w_empty[wslot]);
          }
          accumulate = 1;
        }
        if (leader) tc::mma_commit(&a_empty[slot]);
        __syncwarp();
      }
      if (leader) tc::mma_commit(&acc_full[as]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int e = warp - 4, m = e * 32 + lane, hh = m >> 3, ww = m & 7;
    uint32_t acc_it = 0;
    int nt_staged = -1;
    const int OH = 2 * p.H, OW = 2 * p.W;
    const size_t OHW = (size_t)OH * OW;
    for (int s = blockIdx.x; s < total; s += gridDim.x, ++acc_it) {
      int nt, b, h0, w0;
      decode(p, s, nt, b, h0, w0);
      if (nt != nt_staged) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int i = threadIdx.x - 128; i < NP; i += 128) {
          const int co = nt * NP + i;
          s_scale[i] = (p.scale && co < p.cout) ? __ldg(p.scale + co) : 1.0f;
          s_shift[i] = (p.shift && co < p.cout) ? __ldg(p.shift + co) : 0.0f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        nt_staged = nt;
      }
      const int h = h0 + hh, w = w0 + ww;
      const bool valid = h < p.H && w < p.W;
      const uint32_t as = acc_it & 1;
      tc::mbar_wait(&acc_full[as], (acc_it >> 1) & 1);
      tc::fence_after_sync();
      const uint32_t tbase = tmem_base + ((uint32_t)(e * 32) << 16) + as * ACC;
      const float lo = p.relu ? 0.0f : -INFINITY;
      if (NP >= 32) {
        // columns [phase][NP]: take the two horizontally adjacent phases (ph,0),(ph,1) together so stores are 32 B runs
#pragma unroll 1
        for (int ph = 0; ph < 2; ++ph)
#pragma unroll 1
          for (int j = 0; j < NP / 32; ++j) {
            float v0[32], v1[32];
            tc::tmem_ld32(tbase + (2 * ph) * NP + j * 32, v0);
            tc::tmem_ld32(tbase + (2 * ph + 1) * NP + j * 32, v1);
            if (ph == 1 && j == NP / 32 - 1) {
              tc::fence_before_sync();
              tc::mbar_arrive(&acc_empty[as]);
            }
            const int co0 = nt * NP + j * 32;
            if (!valid || co0 >= p.cout) continue;
            affine_relu32(v0, s_scale + j * 32, s_shift + j * 32, p.relu);
            affine_relu32(v1, s_scale + j * 32, s_shift + j * 32, p.relu);
            const size_t sp = (size_t)(2 * h + ph) * OW + 2 * w;
            if (p.out_mode == 1) {
              float* o = reinterpret_cast<float*>(p.out) + ((size_t)b * p.cout + co0) * OHW + sp;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (co0 + i < p.cout) *reinterpret_cast<float2*>(o + (size_t)i * OHW) = make_float2(v0[i], v1[i]);
            } else {
              uint4* o = reinterpret_cast<uint4*>(p.out) + ((size_t)b * (p.cout / 8) + co0 / 8) * OHW + sp;
#pragma unroll
              for (int c8 = 0; c8 < 4; ++c8)
                if (co0 + 8 * c8 < p.cout) {
                  o[(size_t)c8 * OHW] = pack8(v0 + 8 * c8);
                  o[(size_t)c8 * OHW + 1] = pack8(v1 + 8 * c8);
                }
            }
          }
      } else {
        // NP == 16: one 32-column load holds phases (ph,0) and (ph,1)
#pragma unroll 1
        for (int ph = 0; ph < 2; ++ph) {
          float v[32];
          tc::tmem_ld32(tbase + ph * 32, v);
          if (ph == 1) {
            tc::fence_before_sync();
            tc::mbar_arrive(&acc_empty[as]);
          }
          if (!valid) continue;
          const int co0 = nt * NP;
          const size_t sp = (size_t)(2 * h + ph) * OW + 2 * w;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            v[i] = fmaxf(fmaf(v[i], s_scale[i], s_shift[i]), lo);
            v[16 + i] = fmaxf(fmaf(v[16 + i], s_scale[i], s_shift[i]), lo);
          }
          if (p.out_mode == 1) {
            float* o = reinterpret_cast<float*>(p.out) + ((size_t)b * p.cout + co0) * OHW + sp;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (co0 + i < p.cout) *reinterpret_cast<float2*>(o + (size_t)i * OHW) = make_float2(v[i], v[16 + i]);
          } else {
            uint4* o = reinterpret_cast<uint4*>(p.out) + ((size_t)b * (p.cout / 8) + co0 / 8) * OHW + sp;
#pragma unroll
            for (int c8 = 0; c8 < 2; ++c8)
              if (co0 + 8 * c8 < p.cout) {
                o[(size_t)c8 * OHW] = pack8(v + 8 * c8);
                o[(size_t)c8 * OHW + 1] = pack8(v + 16 + 8 * c8);
              }
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, 2 * ACC);
}

// bilinear x2, align_corners=False (segmenthead's F.interpolate, models/submodule.py:46-51): fp32 NCHW planes.
__global__ void __launch_bounds__(256) bilinear_up2_kernel(const float* __restrict__ in, float* __restrict__ out, int h, int w) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y;
  const size_t plane = blockIdx.z;
  if (X >= 2 * w) return;
  const float sy = fmaxf(0.0f, (Y + 0.5f) * 0.5f - 0.5f), sx = fmaxf(0.0f, (X + 0.5f) * 0.5f - 0.5f);
  const int y0 = (int)sy, x0 = (int)sx, y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  const float ly = sy - y0, lx = sx - x0;
  const float* src = in + plane * h * w;
  const float top = (1.0f - lx) * __ldg(src + (size_t)y0 * w + x0) + lx * __ldg(src + (size_t)y0 * w + x1);
  const float bot = (1.0f - lx) * __ldg(src + (size_t)y1 * w + x0) + lx * __ldg(src + (size_t)y1 * w + x1);
  out[plane * 4 * h * w + (size_t)Y * 2 * w + X] = (1.0f - ly) * top + ly * bot;
}

int make_tmap2d(CUtensorMap* tm, const void* base, int W, int H, long long outer) {
  ss_encode_tiled_fn enc = ss_get_encode_tiled();
  if (!enc) return SS_ERR_CUDA;
  cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, 1u, (cuuint64_t)outer};
  cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)H * W * 16};
  cuuint32_t box[4] = {(cuuint32_t)WW * 8, (cuuint32_t)HH, 1u, (cuuint32_t)(CB / 8)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ss_set_error("cuTensorMapEncodeTiled failed with CUresult %d (W=%d H=%d outer=%lld)", (int)r, W, H, outer);
    return SS_ERR_CUDA;
  }
  return SS_OK;
}

template <typename K>
int launch2d(K kernel, size_t smem, const CUtensorMap& t0, const CUtensorMap& t1, const C2P& p, cudaStream_t st, const char* name, int threads = 256) {
  SS_CUDA(ss_allow_smem(kernel, smem));
  const long long total = (long long)p.items * p.n_tiles;
  const int grid = (int)(total < ss_num_sms() ? total : ss_num_sms());
  kernel<<<grid, threads, smem, st>>>(t0, t1, p);
  SS_CHECK_LAUNCH(name);
  return SS_OK;
}

template <int N, int TAPS>
int launch_conv(const CUtensorMap& t0, const CUtensorMap& t1, const C2P& p, cudaStream_t st) {
  constexpr int NS = 4, NWS = N >= 128 ? 4 : 6;
  constexpr size_t smem = (size_t)NS * SLICE + (size_t)NWS * CB * N * 2;
  static_assert(smem <= 227 * 1024 - 2048, "shared memory budget");
  constexpr int EPG = TAPS == 1 ? (N >= 64 ? 3 : 2) : 1;
  if (TAPS == 1 && (p.act == 2 || p.residual))
    return launch2d(conv2d_tc_kernel<N, TAPS, NS, NWS, TAPS == 1, EPG>, smem, t0, t1, p, st, "ss_conv2d_tc(conv, ext)", 128 + 128 * EPG);
  return launch2d(conv2d_tc_kernel<N, TAPS, NS, NWS, false, EPG>, smem, t0, t1, p, st, "ss_conv2d_tc(conv)", 128 + 128 * EPG);
}
template <int NP>
int launch_deconv(const CUtensorMap& t0, const CUtensorMap& t1, const C2P& p, cudaStream_t st) {
  constexpr int NS = 4, NWS = 3;
  constexpr size_t smem = (size_t)NS * SLICE + (size_t)NWS * CB * 4 * NP * 2;
  static_assert(smem <= 227 * 1024 - 2048, "shared memory budget");
  return launch2d(deconv2d_tc_kernel<NP, NS, NWS>, smem, t0, t1, p, st, "ss_conv2d_tc(deconv)");
}

}  // namespace

// mode: 0 = Conv2d 3x3 s1 p1, 1 = Conv2d 1x1, 2 = ConvTranspose2d k4 s2 p1.  Returns the Cout tile (per output phase for mode 2).
extern "C" int ss_conv2d_tc_ntile(int mode, int Cin, int Cout) {
  if (Cin <= 0 || Cin % CB || Cout <= 0) return 0;
  if (mode == 2) return Cout <= 16 ? 16 : 64;
  if (mode != 0 && mode != 1) return 0;
  return Cout <= 16 ? 16 : Cout <= 32 ? 32 : Cout <= 64 ? 64 : 128;
}

extern "C" int ss_conv2d_tc(int mode, const void* in0_blocked, int C0, const void* in1_blocked_or_null, int C1, const void* weight_packed,
                            const float* scale_or_null, const float* shift_or_null, void* out, int out_mode, int B, int Cout, int H, int W,
                            int relu, void* stream) {
  return ss_conv2d_tc_ex(mode, in0_blocked, C0, in1_blocked_or_null, C1, weight_packed, scale_or_null, shift_or_null, nullptr, out, out_mode,
                         B, Cout, H, W, relu ? 1 : 0, stream);
}

// act: 0 none, 1 ReLU, 2 SiLU (mode 1 only); residual_blocked: bf16 blocked (B,Cout/8,H,W,8) added after the activation (mode 1 only).
extern "C" int ss_conv2d_tc_ex(int mode, const void* in0_blocked, int C0, const void* in1_blocked_or_null, int C1, const void* weight_packed,
                               const float* scale_or_null, const float* shift_or_null, const void* residual_blocked_or_null, void* out,
                               int out_mode, int B, int Cout, int H, int W, int act, void* stream) {
  const int relu = act == 1;
  SS_REQUIRE(act >= 0 && act <= 2, "ss_conv2d_tc: act must be 0 (none), 1 (ReLU) or 2 (SiLU)");
  SS_UNSUPPORTED(mode != 1 && (act == 2 || residual_blocked_or_null), "ss_conv2d_tc: SiLU / residual epilogue exists for the 1x1 mode only");
  SS_REQUIRE(!residual_blocked_or_null || (Cout % 8 == 0 && (reinterpret_cast<uintptr_t>(residual_blocked_or_null) & 15) == 0),
             "ss_conv2d_tc: the residual is a 16-byte aligned bf16 blocked tensor (Cout %% 8 == 0)");
  SS_REQUIRE(in0_blocked && weight_packed && out, "ss_conv2d_tc: null pointer");
  SS_REQUIRE(B > 0 && H > 0 && W > 0 && Cout > 0 && C0 > 0 && C1 >= 0, "ss_conv2d_tc: non-positive dimension");
  SS_REQUIRE((in1_blocked_or_null != nullptr) == (C1 > 0), "ss_conv2d_tc: second input and its channel count must come together");
  SS_UNSUPPORTED(C0 % CB || C1 % CB, "ss_conv2d_tc: input channel counts (%d, %d) must be multiples of %d", C0, C1, CB);
  const int N = ss_conv2d_tc_ntile(mode, C0 + C1, Cout);
  SS_UNSUPPORTED(N == 0, "ss_conv2d_tc: mode %d with (Cin=%d, Cout=%d) has no tensor-core configuration", mode, C0 + C1, Cout);
  SS_REQUIRE(out_mode == 0 || out_mode == 1, "ss_conv2d_tc: out_mode must be 0 (bf16 blocked) or 1 (fp32 NCHW)");
  SS_REQUIRE(out_mode == 1 || Cout % 8 == 0, "ss_conv2d_tc: a bf16 blocked output needs Cout %% 8 == 0");
  SS_REQUIRE(((reinterpret_cast<uintptr_t>(in0_blocked) | reinterpret_cast<uintptr_t>(in1_blocked_or_null) |
               reinterpret_cast<uintptr_t>(weight_packed) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
             "ss_conv2d_tc: pointers must be 16-byte aligned");
  C2P p;
  p.w = reinterpret_cast<const __nv_bfloat16*>(weight_packed);
  p.scale = scale_or_null; p.shift = shift_or_null; p.out = out; p.out_mode = out_mode; p.cout = Cout;
  p.act = act; p.residual = reinterpret_cast<const __nv_bfloat16*>(residual_blocked_or_null);
  p.B = B; p.H = H; p.W = W; p.relu = relu;
  p.ncb0 = C0 / CB; p.ncb = (C0 + C1) / CB;
  p.n_tiles = ceil_div(Cout, N);
  p.HT = ceil_div(H, TH); p.WT = ceil_div(W, TW);
  SS_UNSUPPORTED((long long)B * p.HT * p.WT * p.n_tiles > 0x7fffffffLL, "ss_conv2d_tc: too many tiles");
  p.items = B * p.HT * p.WT;
  CUtensorMap t0, t1;
  int rc = make_tmap2d(&t0, in0_blocked, W, H, (long long)B * (C0 / 8));
  if (rc != SS_OK) return rc;
  t1 = t0;
  if (C1 > 0 && (rc = make_tmap2d(&t1, in1_blocked_or_null, W, H, (long long)B * (C1 / 8))) != SS_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 2) return N == 16 ? launch_deconv<16>(t0, t1, p, st) : launch_deconv<64>(t0, t1, p, st);
  if (mode == 0) {
    switch (N) {
      case 16: return launch_conv<16, 9>(t0, t1, p, st);
      case 32: return launch_conv<32, 9>(t0, t1, p, st);
      case 64: return launch_conv<64, 9>(t0, t1, p, st);
      default: return launch_conv<128, 9>(t0, t1, p, st);
    }
  }
  switch (N) {
    case 16: return launch_conv<16, 1>(t0, t1, p, st);
    case 32: return launch_conv<32, 1>(t0, t1, p, st);
    case 64: return launch_conv<64, 1>(t0, t1, p, st);
    default: return launch_conv<128, 1>(t0, t1, p, st);
  }
}

extern "C" int ss_bilinear_up2(const float* in, float* out, int planes, int h, int w, void* stream) {
  SS_REQUIRE(in && out && planes > 0 && h > 0 && w > 0, "ss_bilinear_up2: bad argument");
  SS_UNSUPPORTED(2 * h > 65535 || planes > 65535, "ss_bilinear_up2: grid dimension exceeds 65535");
  bilinear_up2_kernel<<<dim3(ceil_div(2 * w, 256), 2 * h, planes), 256, 0, (cudaStream_t)stream>>>(in, out, h, w);
  SS_CHECK_LAUNCH("ss_bilinear_up2");
  return SS_OK;
}
