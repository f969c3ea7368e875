// K4, tensor-core mode: the Cout = 1 classifier heads, nn.Conv3d(32, 1, 3, padding=1, bias=False)
// (reference models/SemStereo.py:230,234 -- classif_att_.2 / classif.2).
//
// Run as an ordinary implicit GEMM this layer wastes the tensor core: N would be 1 (padded to 32) and every one of the 27 taps
// re-reads a 128 x 32 A tile.  Here the roles are swapped ("taps as N") and the depth taps are folded as in the s1f kernel:
//     Q_dout[v][t9] = sum_kd sum_c X[dout+kd-1, v][c] * W[kd, t9][c]     for EVERY voxel v of the staged halo tile,
// built by ONE small GEMM per INPUT slice: M = 180 halo voxels (two overlapping 128-row blocks), K = 32, N = 48 = 3 depth taps
// x 16 columns (9 in-plane taps, 7 zero).  Column block j of the product belongs to output depth d_in - 1 + j; the
// accumulators of consecutive output depths are consecutive 16-column blocks of a TMEM ring (16 depths per row block), so
// the three partial products land where they belong (always accumulating; the epilogue zeroes a block after draining it).
// The epilogue then only has the in-plane part left:
//     out[d, h, w] = sum_{kh,kw} Q_d[(h+kh, w+kw)][kh*3 + kw]
// a 9-term shifted sum from a shared-memory copy of Q (9 floats per halo voxel instead of 27 per slice and 27 reads).
// The halo tile of a slice in shared memory ([4 chunks][18][10][16 B]) is a K-major no-swizzle A operand whose rows are the
// 180 halo voxels in linear order (SBO = 128 B, LBO = chunk pitch), so the same TMA box as the other conv kernels feeds it.
#include "tc_common.cuh"

namespace {

constexpr int TH = 16, TW = 8, HH = TH + 2, WW = TW + 2, NV = HH * WW;      // 180 halo voxels
constexpr uint32_t TILE_B = NV * 16;                                        // bytes of one channel chunk of a halo tile
constexpr int CIN = 32, C8 = CIN / 8, NT = 16, NROW = 3 * NT;               // 9 in-plane taps padded to 16 columns, x 3 depth taps
constexpr uint32_t SLICE = C8 * TILE_B;                                     // 11520 B
constexpr uint32_t WBYTES = C8 * NROW * 16;                                 // 3072 B of weights
constexpr int NS = 4;                                                       // input-slice ring
constexpr uint32_t NB = 8;                                                  // accumulator ring: output depths in flight
constexpr uint32_t RB_COLS = NB * 16;                                       // TMEM columns of one 128-row block (NB depths x 16 columns)
constexpr int QSTRIDE = 9;                                                  // floats per halo voxel in the smem copy of Q
constexpr int ROWB = NV - 128;                                              // first halo voxel of the second 128-row block (52)

struct HeadP {
  const __nv_bfloat16* w;   // [C8][48][8] bf16: w[chunk][j*16 + t9][c8] = weight[0][chunk*8+c8][kd = 2-j][t9]
  const float* acc_in;      // (B,1,D,H,W) fp32 partial sums added to the result (bf16x3 split route), or null
  const float* acc_in2;
  float* out;               // (B,1,D,H,W) fp32
  int B, D, H, W;
  int HT, WT, DC, n_dc, items;
};

__device__ __forceinline__ void decode_item(const HeadP& p, int s, int& b, int& h0, int& w0, int& dlo, int& dhi) {
  const int dc = s % p.n_dc;  s /= p.n_dc;
  const int wt = s % p.WT;    s /= p.WT;
  const int ht = s % p.HT;
  b = s / p.HT;
  h0 = ht * TH; w0 = wt * TW;
  dlo = dc * p.DC; dhi = min(p.D, dlo + p.DC);
}

// SP (in-kernel bf16x3 split): split input tensor (hi batches | lo batches), slice = [hi chunks][lo chunks], weights [hi][lo],
// three MMAs per K step: x_hi*w_hi + x_lo*w_hi + x_hi*w_lo.
template <bool SP>
__global__ void __launch_bounds__(256, 2) conv3d_tc_head_kernel(const __grid_constant__ CUtensorMap tmA, const HeadP p) {
  constexpr uint32_t HALF_A = C8 * TILE_B, HALF_B = C8 * NROW * 16;
  constexpr uint32_t SLICE = (SP ? 2 : 1) * HALF_A, WBYTES = (SP ? 2 : 1) * HALF_B;      // shadow the single-operand sizes
  constexpr uint32_t LBO_A = TILE_B, SBO_A = 128, LBO_B = NROW * 16, SBO_B = 128;
  // 2 row blocks x NB depths x 16 columns = 256 of the SM's 512 TMEM columns: TWO CTAs per SM, so that one CTA's epilogue chain
  // (TMEM -> smem copy of Q -> barrier -> 9-term gather -> store, ~0.45 us per depth) overlaps the other's (round 1: 1 CTA/SM).
  constexpr uint32_t TMEM_COLS = 2 * RB_COLS;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t a_full[NS], a_empty[NS], acc_full[NB], acc_empty[NB], w_full;
  __shared__ uint32_t tmem_base_s;
  uint8_t* Abase = smem;
  uint8_t* Wbase = smem + NS * SLICE;
  float* Qs = reinterpret_cast<float*>(Wbase + WBYTES);             // [2][NV][QSTRIDE]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmA);
    for (int i = 0; i < NS; ++i) { tc::mbar_init(&a_full[i], 1); tc::mbar_init(&a_empty[i], 1); }
    for (uint32_t i = 0; i < NB; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 128); }
    tc::mbar_init(&w_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  if (warp >= 4) {                                                  // all accumulator blocks start out zero
#pragma unroll 1
    for (uint32_t c = 0; c < TMEM_COLS; c += 32) tc::tmem_zero32(tmem_base + ((uint32_t)((warp - 4) * 32) << 16) + c);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const int cta_s = blockIdx.x, cta_stride = gridDim.x;

  if (warp == 0 && lane == 0) {
    // ===== producer: weights once, then the input slices =====
    if (cta_s < p.items) {
      tc::mbar_expect_tx(&w_full, WBYTES);
      tc::bulk_load(Wbase, p.w, WBYTES, &w_full);
    }
    uint32_t g = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int din0 = max(dlo - 1, 0), din1 = min(dhi, p.D - 1);
      for (int d_in = din0; d_in <= din1; ++d_in, ++g) {
        const uint32_t slot = g % NS;
        tc::mbar_wait(&a_empty[slot], ((g / NS) & 1) ^ 1);
        tc::mbar_expect_tx(&a_full[slot], SLICE);
        tc::tma_load_4d(Abase + slot * SLICE, &tmA, &a_full[slot], (w0 - 1) * 8, h0 - 1, d_in, b * C8);
        if (SP) tc::tma_load_4d(Abase + slot * SLICE + HALF_A, &tmA, &a_full[slot], (w0 - 1) * 8, h0 - 1, d_in, (b + p.B) * C8);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: per input slice, [Q_{d-1} | Q_d | Q_{d+1}] += X_d * W^T for the two 128-row blocks of the halo tile =====
    const bool leader = tc::elect_one();
    const uint32_t a_lo0 = tc::desc_lo(tc::smem_u32(Abase), LBO_A), a_hi = tc::desc_hi(SBO_A);
    const uint32_t b_lo0 = tc::desc_lo(tc::smem_u32(Wbase), LBO_B), b_hi = tc::desc_hi(SBO_B);
    if (cta_s < p.items) tc::mbar_wait(&w_full, 0);
    uint32_t g = 0, acc_base = 0, acquired = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int din0 = max(dlo - 1, 0), din1 = min(dhi, p.D - 1);
#pragma unroll 1
      for (int d_in = din0; d_in <= din1; ++d_in, ++g) {
        const int j0 = max(0, dlo - d_in + 1), j1 = min(3, dhi - d_in + 1);       // column blocks [j0, j1) exist
        const uint32_t u0 = acc_base + (uint32_t)(d_in - 1 + j0 - dlo), nb = (uint32_t)(j1 - j0);
        while (acquired < u0 + nb) {
          tc::mbar_wait(&acc_empty[acquired % NB], ((acquired / NB) & 1) ^ 1);
          ++acquired;
        }
        const uint32_t slot = g % NS;
        tc::mbar_wait(&a_full[slot], (g / NS) & 1);
        tc::fence_after_sync();
        const uint32_t blk = u0 % NB, n1 = min(nb, NB - blk), n2 = nb - n1;
        const uint32_t id1 = n1 == 1 ? tc::make_idesc_bf16(128, NT) : n1 == 2 ? tc::make_idesc_bf16(128, 2 * NT) : tc::make_idesc_bf16(128, 3 * NT);
        const uint32_t id2 = n2 == 1 ? tc::make_idesc_bf16(128, NT) : tc::make_idesc_bf16(128, 2 * NT);
        const uint32_t brow1 = (uint32_t)j0 * (NT / 8) * (SBO_B >> 4), brow2 = (uint32_t)(j0 + n1) * (NT / 8) * (SBO_B >> 4);
        if (leader) {
          const uint32_t a_lo = a_lo0 + slot * (SLICE >> 4);
#pragma unroll
          for (int rb = 0; rb < 2; ++rb)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint32_t a = a_lo + (uint32_t)(rb * ROWB) + (uint32_t)(ks * 2 * LBO_A) / 16;
              const uint32_t bb = b_lo0 + (uint32_t)(ks * 2 * LBO_B) / 16;
              tc::mma_bf16_lohi(tmem_base + rb * RB_COLS + blk * NT, a, a_hi, bb + brow1, b_hi, id1, 1u);
              if (n2) tc::mma_bf16_lohi(tmem_base + rb * RB_COLS, a, a_hi, bb + brow2, b_hi, id2, 1u);
              if (SP) {
                tc::mma_bf16_lohi(tmem_base + rb * RB_COLS + blk * NT, a + (HALF_A >> 4), a_hi, bb + brow1, b_hi, id1, 1u);
                if (n2) tc::mma_bf16_lohi(tmem_base + rb * RB_COLS, a + (HALF_A >> 4), a_hi, bb + brow2, b_hi, id2, 1u);
                tc::mma_bf16_lohi(tmem_base + rb * RB_COLS + blk * NT, a, a_hi, bb + (HALF_B >> 4) + brow1, b_hi, id1, 1u);
                if (n2) tc::mma_bf16_lohi(tmem_base + rb * RB_COLS, a, a_hi, bb + (HALF_B >> 4) + brow2, b_hi, id2, 1u);
              }
            }
          tc::mma_commit(&a_empty[slot]);
          if (d_in - 1 >= dlo) tc::mma_commit(&acc_full[(acc_base + (uint32_t)(d_in - 1 - dlo)) % NB]);
          if (d_in == din1 && din1 == dhi - 1) tc::mma_commit(&acc_full[(acc_base + (uint32_t)(d_in - dlo)) % NB]);
        }
        __syncwarp();
      }
      acc_base += (uint32_t)(dhi - dlo);
    }
  } else if (warp >= 4) {
    // ===== epilogue: per completed output depth, TMEM -> smem copy of Q (zeroing the block), then the 9-term shifted sum =====
    const int e = warp - 4, m = e * 32 + lane, hh = m >> 3, ww = m & 7;
    uint32_t u = 0;
    const size_t HWs = (size_t)p.H * p.W;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int h = h0 + hh, w = w0 + ww;
      const bool valid = h < p.H && w < p.W;
      for (int d_out = dlo; d_out < dhi; ++d_out, ++u) {
        const uint32_t blk = u % NB;
        tc::mbar_wait(&acc_full[blk], (u / NB) & 1);
        tc::fence_after_sync();
        float* dst = Qs + (size_t)(u & 1) * NV * QSTRIDE;
        const uint32_t ta = tmem_base + ((uint32_t)(e * 32) << 16) + blk * NT;
        float v[16];
        tc::tmem_ld16(ta, v);                                     // row block A: halo voxel m
        tc::tmem_zero16(ta);
#pragma unroll
        for (int t = 0; t < 9; ++t) dst[m * QSTRIDE + t] = v[t];
        tc::tmem_ld16(ta + RB_COLS, v);                           // row block B: halo voxel ROWB + m
        tc::tmem_zero16(ta + RB_COLS);
        tc::fence_before_sync();
        tc::mbar_arrive(&acc_empty[blk]);
        if (ROWB + m >= 128) {
#pragma unroll
          for (int t = 0; t < 9; ++t) dst[(ROWB + m) * QSTRIDE + t] = v[t];
        }
        // the 4 epilogue warps: Q copy complete.  Qs is double-buffered and a thread is at most one barrier ahead of the
        // slowest one, so the copy of depth u+2 cannot overtake a gather of depth u.
        asm volatile("bar.sync 1, 128;" ::: "memory");
        float acc = 0.0f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) acc += dst[((hh + kh) * WW + ww + kw) * QSTRIDE + kh * 3 + kw];
        if (valid) {
          const size_t o = ((size_t)b * p.D + d_out) * HWs + (size_t)h * p.W + w;
          if (p.acc_in) acc += __ldg(p.acc_in + o);
          if (p.acc_in2) acc += __ldg(p.acc_in2 + o);
          p.out[o] = acc;
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

// in_blocked: bf16 (B,4,D,H,W,8); weight_packed: bf16 [4][48][8] (rows j*16 + t9, kd = 2 - j, 9 real in-plane taps); out fp32 (B,1,D,H,W)
extern "C" int ss_conv3d_tc_head(const void* in_blocked, const void* weight_packed, float* out, int B, int Cin, int D, int H, int W,
                                 void* stream) {
  return ss_conv3d_tc_head_ex(in_blocked, weight_packed, nullptr, nullptr, out, 0, B, Cin, D, H, W, stream);
}

extern "C" int ss_conv3d_tc_head_ex(const void* in_blocked, const void* weight_packed, const float* acc_in_or_null,
                                    const float* acc_in2_or_null, float* out, int in_split, int B, int Cin, int D, int H, int W,
                                    void* stream) {
  SS_REQUIRE(in_blocked && weight_packed && out, "ss_conv3d_tc_head: null pointer");
  SS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "ss_conv3d_tc_head: non-positive dimension");
  SS_UNSUPPORTED(Cin != CIN, "ss_conv3d_tc_head: only Cin = 32 is supported (got %d)", Cin);
  SS_REQUIRE((reinterpret_cast<uintptr_t>(in_blocked) & 15) == 0 && (reinterpret_cast<uintptr_t>(weight_packed) & 15) == 0,
             "ss_conv3d_tc_head: pointers must be 16-byte aligned");
  ss_encode_tiled_fn enc = ss_get_encode_tiled();
  if (!enc) return SS_ERR_CUDA;
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)(in_split ? 2 : 1) * B * C8};
  cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
  cuuint32_t box[4] = {(cuuint32_t)WW * 8, (cuuint32_t)HH, 1u, (cuuint32_t)C8};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in_blocked), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ss_set_error("cuTensorMapEncodeTiled(head) failed with CUresult %d", (int)r);
    return SS_ERR_CUDA;
  }
  HeadP p;
  p.w = reinterpret_cast<const __nv_bfloat16*>(weight_packed);
  p.out = out; p.acc_in = acc_in_or_null; p.acc_in2 = acc_in2_or_null;
  p.B = B; p.D = D; p.H = H; p.W = W;
  p.HT = ceil_div(H, TH); p.WT = ceil_div(W, TW);
  int grid = 2 * ss_num_sms();             // two persistent CTAs per SM (256 TMEM columns each)
  const int spatial = B * p.HT * p.WT;
  int best = D;
  double best_cost = 1e30;
  for (int dc = 1; dc <= D; ++dc) {
    if (D % dc) continue;
    const long long items = (long long)spatial * (D / dc);
    const double cost = (double)ceil_div64(items, grid) * (dc + 1.0);     // every chunk re-reads 2 halo slices (cheap: one pass each)
    if (cost < best_cost - 1e-9) { best_cost = cost; best = dc; }
  }
  p.DC = best; p.n_dc = D / best; p.items = spatial * p.n_dc;
  if (p.items < grid) grid = p.items;
  const size_t smem = (in_split ? 2 : 1) * ((size_t)NS * SLICE + WBYTES) + (size_t)2 * NV * QSTRIDE * sizeof(float);
  if (in_split) {
    SS_CUDA(ss_allow_smem(conv3d_tc_head_kernel<true>, smem));
    conv3d_tc_head_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(tm, p);
  } else {
    SS_CUDA(ss_allow_smem(conv3d_tc_head_kernel<false>, smem));
    conv3d_tc_head_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(tm, p);
  }
  SS_CHECK_LAUNCH("ss_conv3d_tc_head");
  return SS_OK;
}
