// K4, tensor-core mode: the Cout = 1 classifier heads, nn.Conv3d(32, 1, 3, padding=1, bias=False)
// (reference models/SemStereo.py:230,234 -- classif_att_.2 / classif.2).
//
// Run as an ordinary implicit GEMM this layer wastes the tensor core: N would be 1 (padded to 32) and every one of the 27 taps
// re-reads a 128 x 32 A tile.  Here the roles are swapped ("taps as N"):
//     P_d[v][t] = sum_c X[d, v][c] * W[t][c]          for EVERY voxel v of the staged halo tile and all 27 taps t at once,
// i.e. ONE small GEMM per INPUT slice (M = 180 halo voxels as two overlapping 128-row blocks, N = 32 (27 taps), K = 32 ->
// 4 MMAs instead of 54 per output slice), followed by the shifted gather-sum
//     out[d, h, w] = sum_{kd,kh,kw} P_{d+kd-1}[(h+kh, w+kw)][kd*9 + kh*3 + kw]
// which the epilogue warps do from a shared-memory copy of the last three P slices.  The halo tile of a slice in shared
// memory ([4 chunks][18][10][16 B]) is a K-major no-swizzle A operand whose rows are simply the 180 halo voxels in
// linear order (SBO = 128 B, LBO = chunk pitch), so the same TMA box as the other conv kernels feeds it.
#include "tc_common.cuh"

namespace {

constexpr int TH = 16, TW = 8, HH = TH + 2, WW = TW + 2, NV = HH * WW;      // 180 halo voxels
constexpr uint32_t TILE_B = NV * 16;                                        // bytes of one channel chunk of a halo tile
constexpr int CIN = 32, C8 = CIN / 8, NTAP = 32;                            // 27 taps padded to 32 columns
constexpr uint32_t SLICE = C8 * TILE_B;                                     // 11520 B
constexpr int NS = 4;                                                       // input-slice ring
constexpr int PR = 4;                                                       // P ring (TMEM column blocks and smem copies)
constexpr int PSTRIDE = 29;                                                 // floats per halo voxel in the smem copy (odd: no conflicts)
constexpr int ROWB = NV - 128;                                              // first halo voxel of the second 128-row block (52)

struct HeadP {
  const __nv_bfloat16* w;   // [C8][NTAP][8] bf16: w[chunk][tap][c8] = weight[0][chunk*8+c8][tap]
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

__global__ void __launch_bounds__(256, 1) conv3d_tc_head_kernel(const __grid_constant__ CUtensorMap tmA, const HeadP p) {
  constexpr uint32_t LBO_A = TILE_B, SBO_A = 128, LBO_B = NTAP * 16, SBO_B = 128;
  constexpr uint32_t TMEM_COLS = 256;                               // PR slots x 2 blocks x 32 columns
  constexpr uint32_t IDESC = tc::make_idesc_bf16(128, NTAP);
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t a_full[NS], a_empty[NS], p_full[PR], p_empty[PR], w_full;
  __shared__ uint32_t tmem_base_s;
  uint8_t* Abase = smem;
  uint8_t* Wbase = smem + NS * SLICE;                               // 2 KB of weights
  float* Ps = reinterpret_cast<float*>(Wbase + 2048);               // [PR][NV][PSTRIDE]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmA);
    for (int i = 0; i < NS; ++i) { tc::mbar_init(&a_full[i], 1); tc::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < PR; ++i) { tc::mbar_init(&p_full[i], 1); tc::mbar_init(&p_empty[i], 128); }
    tc::mbar_init(&w_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const int cta_s = blockIdx.x, cta_stride = gridDim.x;

  if (warp == 0 && lane == 0) {
    // ===== producer: weights once, then the input slices =====
    if (cta_s < p.items) {
      tc::mbar_expect_tx(&w_full, 2048);
      tc::bulk_load(Wbase, p.w, 2048, &w_full);
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
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: per input slice, P = X * W^T for the two 128-row blocks of the halo tile =====
    const bool leader = tc::elect_one();
    const uint32_t a_lo0 = tc::desc_lo(tc::smem_u32(Abase), LBO_A), a_hi = tc::desc_hi(SBO_A);
    const uint32_t b_lo0 = tc::desc_lo(tc::smem_u32(Wbase), LBO_B), b_hi = tc::desc_hi(SBO_B);
    if (cta_s < p.items) tc::mbar_wait(&w_full, 0);
    uint32_t g = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int din0 = max(dlo - 1, 0), din1 = min(dhi, p.D - 1);
      for (int d_in = din0; d_in <= din1; ++d_in, ++g) {
        const uint32_t slot = g % NS, ps = g % PR;
        tc::mbar_wait(&p_empty[ps], ((g / PR) & 1) ^ 1);
        tc::mbar_wait(&a_full[slot], (g / NS) & 1);
        tc::fence_after_sync();
        if (leader) {
          const uint32_t a_lo = a_lo0 + slot * (SLICE >> 4);
#pragma unroll
          for (int blk = 0; blk < 2; ++blk)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              tc::mma_bf16_lohi(tmem_base + ps * 64 + blk * 32, a_lo + (uint32_t)(blk * ROWB) + (uint32_t)(ks * 2 * LBO_A) / 16, a_hi,
                                b_lo0 + (uint32_t)(ks * 2 * LBO_B) / 16, b_hi, IDESC, ks);
          tc::mma_commit(&p_full[ps]);
          tc::mma_commit(&a_empty[slot]);        // the slice is not needed again: every tap was taken in this one pass
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> smem copy of P, then the shifted 27-term sum for the output slice that just became complete =====
    const int e = warp - 4, m = e * 32 + lane, hh = m >> 3, ww = m & 7;
    uint32_t g = 0;
    const size_t HWs = (size_t)p.H * p.W;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode_item(p, s, b, h0, w0, dlo, dhi);
      const int din0 = max(dlo - 1, 0), din1 = min(dhi, p.D - 1);
      const int h = h0 + hh, w = w0 + ww;
      const bool valid = h < p.H && w < p.W;
      const uint32_t g0 = g;                                      // ring index of slice din0
      for (int d_in = din0; d_in <= din1; ++d_in, ++g) {
        const uint32_t ps = g % PR;
        tc::mbar_wait(&p_full[ps], (g / PR) & 1);
        tc::fence_after_sync();
        float* dst = Ps + (size_t)ps * NV * PSTRIDE;
        {
          float v[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(e * 32) << 16) + ps * 64, v);            // block A: halo voxel m
#pragma unroll
          for (int t = 0; t < 27; ++t) dst[m * PSTRIDE + t] = v[t];
          tc::tmem_ld32(tmem_base + ((uint32_t)(e * 32) << 16) + ps * 64 + 32, v);       // block B: halo voxel ROWB + m
          tc::fence_before_sync();
          tc::mbar_arrive(&p_empty[ps]);
          if (ROWB + m >= 128) {
#pragma unroll
            for (int t = 0; t < 27; ++t) dst[(ROWB + m) * PSTRIDE + t] = v[t];
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");           // the 4 epilogue warps: P copy complete / previous gather done
        // slice d_in completes output d_in - 1; the last slice of the item also completes output d_in when it is the volume's end
        for (int d_out = d_in - 1; d_out <= d_in; ++d_out) {
          if (d_out < dlo || d_out >= dhi) continue;
          if (d_out == d_in && !(d_in == din1 && d_in == dhi - 1)) continue;     // d_out == d_in only when no slice d_in+1 will come
          float acc = 0.0f;
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) {
            const int ds = d_out + kd - 1;
            if (ds < din0 || ds > din1) continue;                  // outside the volume: zero padding
            const float* src = Ps + (size_t)((g0 + (uint32_t)(ds - din0)) % PR) * NV * PSTRIDE;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) acc += src[((hh + kh) * WW + ww + kw) * PSTRIDE + kd * 9 + kh * 3 + kw];
          }
          if (valid) p.out[((size_t)b * p.D + d_out) * HWs + (size_t)h * p.W + w] = acc;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");             // all gathers of this item done before its P copies are overwritten
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

// in_blocked: bf16 (B,4,D,H,W,8); weight_packed: bf16 [4][32][8] (tap-major rows, 27 real taps); out fp32 (B,1,D,H,W)
extern "C" int ss_conv3d_tc_head(const void* in_blocked, const void* weight_packed, float* out, int B, int Cin, int D, int H, int W,
                                 void* stream) {
  SS_REQUIRE(in_blocked && weight_packed && out, "ss_conv3d_tc_head: null pointer");
  SS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "ss_conv3d_tc_head: non-positive dimension");
  SS_UNSUPPORTED(Cin != CIN, "ss_conv3d_tc_head: only Cin = 32 is supported (got %d)", Cin);
  SS_REQUIRE((reinterpret_cast<uintptr_t>(in_blocked) & 15) == 0 && (reinterpret_cast<uintptr_t>(weight_packed) & 15) == 0,
             "ss_conv3d_tc_head: pointers must be 16-byte aligned");
  ss_encode_tiled_fn enc = ss_get_encode_tiled();
  if (!enc) return SS_ERR_CUDA;
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B * C8};
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
  p.out = out;
  p.B = B; p.D = D; p.H = H; p.W = W;
  p.HT = ceil_div(H, TH); p.WT = ceil_div(W, TW);
  int grid = ss_num_sms();
  const int spatial = B * p.HT * p.WT;
  int best = D;
  double best_cost = 1e30;
  for (int dc = 1; dc <= D; ++dc) {
    if (D % dc) continue;
    const long long items = (long long)spatial * (D / dc);
    const double cost = (double)ceil_div64(items, grid) * (dc + 2.0);     // every chunk re-reads (and re-multiplies) 2 halo slices
    if (cost < best_cost - 1e-9) { best_cost = cost; best = dc; }
  }
  p.DC = best; p.n_dc = D / best; p.items = spatial * p.n_dc;
  if (p.items < grid) grid = p.items;
  const size_t smem = (size_t)NS * SLICE + 2048 + (size_t)PR * NV * PSTRIDE * sizeof(float);
  SS_CUDA(ss_allow_smem(conv3d_tc_head_kernel, smem));
  conv3d_tc_head_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(tm, p);
  SS_CHECK_LAUNCH("ss_conv3d_tc_head");
  return SS_OK;
}
