// K1 (fast path): group-wise correlation volume with TMA-staged rows.
// Same math as gwc_volume_kernel (volumes.cu; reference models/submodule.py:190-255, models/submodule_.py:180-237), but:
//   * persistent CTAs; a work item = (sample, image row, group chunk, x-tile);
//   * the left row tile [CC][TX] and the right row tile [CC][TX+E8] (zero halo for every shift) are fetched by two TMA boxes
//     (out-of-bounds zero fill = the invalid wedge of the volume) into a 2-stage ring, so the loads of item i+1 overlap the
//     FMAs and stores of item i and no thread spends issue slots on global loads;
//   * the per-group L2 norms are reduced once per staged column into a small table and applied to the accumulators
//     (acc * invL[x] * invR[x-d]) instead of rescaling the staged rows in place;
//   * 4(x) x 8(shift) register tiles, 128-bit shared loads, 128-bit streaming stores.
// Used when W % 4 == 0 (TMA needs 16-byte row strides); other shapes take the generic kernel in volumes.cu.
#include "tc_common.cuh"

namespace {

struct GwcT {
  float* out;
  int B, C, H, W, G, cg, D, dmax, norm;
  int GC, TX, E8, RW, n_xt, n_gc;
  long long items;
  uint32_t stage_bytes;
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                   tc::smem_u32(smem_dst)),
               "l"(tmap), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// PADL = columns between the 16-byte-aligned start of the staged right tile and column x0 - dmax (TMA boxes must start on a
// 16-byte boundary in the innermost dimension); it only shifts the register-tile indexing.
template <int THREADS, int PADL>
__global__ void __launch_bounds__(THREADS) gwc_volume_tma_kernel(const __grid_constant__ CUtensorMap tmL,
                                                                 const __grid_constant__ CUtensorMap tmR, const GwcT p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[2];
  const int CC = p.GC * p.cg;
  const uint32_t l_bytes = (uint32_t)CC * p.TX * 4, r_bytes = (uint32_t)CC * p.RW * 4;
  float* inv = reinterpret_cast<float*>(smem_raw + 2 * p.stage_bytes);      // invL [GC][TX], invR [GC][RW]

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmL);
    tc::prefetch_tmap(&tmR);
    tc::mbar_init(&full[0], 1);
    tc::mbar_init(&full[1], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();

  // NOTE: the descriptor addresses are taken here, at kernel scope: taking them inside a by-reference lambda makes the
  // compiler spill the __grid_constant__ parameters to local memory, and TMA faults on a descriptor that is not in param space.
  const CUtensorMap* const pL = &tmL;
  const CUtensorMap* const pR = &tmR;
  auto issue = [=](long long item, int stage) {       // one thread
    long long t = item;
    const int xt = (int)(t % p.n_xt);  t /= p.n_xt;
    const int gc = (int)(t % p.n_gc);  t /= p.n_gc;
    const int y = (int)(t % p.H);
    const int b = (int)(t / p.H);
    uint8_t* st = smem_raw + (size_t)stage * p.stage_bytes;
    tc::mbar_expect_tx(&full[stage], l_bytes + r_bytes);
    tma_load_3d(st, pL, &full[stage], xt * p.TX, y, b * p.C + gc * CC);
    tma_load_3d(st + l_bytes, pR, &full[stage], xt * p.TX - p.dmax - PADL, y, b * p.C + gc * CC);
  };

  long long item = blockIdx.x;
  if (threadIdx.x == 0 && item < p.items) issue(item, 0);
  const int nxq = p.TX >> 2, nec = p.E8 >> 3;
  const int ntiles = p.GC * nec * nxq;
  const float inv_cg = 1.0f / (float)p.cg;
  uint32_t it = 0;
  for (; item < p.items; item += gridDim.x, ++it) {
    const int stage = it & 1;
    const long long next = item + gridDim.x;
    if (threadIdx.x == 0 && next < p.items) issue(next, stage ^ 1);     // stage^1 was released by the barrier ending the last item
    tc::mbar_wait(&full[stage], (it >> 1) & 1);
    const float* Ls = reinterpret_cast<const float*>(smem_raw + (size_t)stage * p.stage_bytes);
    const float* Rs = Ls + (size_t)CC * p.TX;
    long long t = item;
    const int xt = (int)(t % p.n_xt);  t /= p.n_xt;
    const int gc = (int)(t % p.n_gc);  t /= p.n_gc;
    const int y = (int)(t % p.H);
    const int b = (int)(t / p.H);
    const int x0 = xt * p.TX;

    if (p.norm) {      // 1 / (||column||_2 + eps) per (group, staged column)   (groupwise_correlation_norm, submodule.py:218)
      for (int i = threadIdx.x; i < p.GC * (p.TX + p.RW); i += THREADS) {
        const int g = i / (p.TX + p.RW), j = i - g * (p.TX + p.RW);
        const float* col;
        int stride;
        if (j < p.TX) { col = Ls + (size_t)g * p.cg * p.TX + j; stride = p.TX; }
        else          { col = Rs + (size_t)g * p.cg * p.RW + (j - p.TX); stride = p.RW; }
        float ss = 0.0f;
        for (int c = 0; c < p.cg; ++c) { const float v = col[c * stride]; ss = fmaf(v, v, ss); }
        inv[i] = 1.0f / (sqrtf(ss) + 1e-5f);
      }
      __syncthreads();
    }

    for (int tl = threadIdx.x; tl < ntiles; tl += THREADS) {
      const int xq = tl % nxq;
      const int ec = (tl / nxq) % nec;
      const int gl = tl / (nxq * nec);
      float acc[8][4];
#pragma unroll
      for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[e][i] = 0.0f;
      const float* lp = Ls + (size_t)gl * p.cg * p.TX + 4 * xq;
      const float* rp = Rs + (size_t)gl * p.cg * p.RW + 4 * xq + 8 * ec;
      for (int c = 0; c < p.cg; ++c) {
        const float4 l4 = *reinterpret_cast<const float4*>(lp + (size_t)c * p.TX);
        const float4 r0 = *reinterpret_cast<const float4*>(rp + (size_t)c * p.RW);
        const float4 r1 = *reinterpret_cast<const float4*>(rp + (size_t)c * p.RW + 4);
        const float4 r2 = *reinterpret_cast<const float4*>(rp + (size_t)c * p.RW + 8);
        float4 r3 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (PADL > 1) r3 = *reinterpret_cast<const float4*>(rp + (size_t)c * p.RW + 12);
        const float l[4] = {l4.x, l4.y, l4.z, l4.w};
        const float r[16] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w};
#pragma unroll
        for (int e = 0; e < 8; ++e)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[e][i] = fmaf(l[i], r[e + i + PADL], acc[e][i]);
      }
      const int x = x0 + 4 * xq;
      if (x >= p.W) continue;
      float il[4] = {inv_cg, inv_cg, inv_cg, inv_cg}, ir[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) ir[i] = 1.0f;
      if (p.norm) {
        const float* ilp = inv + (size_t)gl * (p.TX + p.RW) + 4 * xq;
        const float* irp = inv + (size_t)gl * (p.TX + p.RW) + p.TX + 4 * xq + 8 * ec;
#pragma unroll
        for (int i = 0; i < 4; ++i) il[i] = ilp[i] * inv_cg;
#pragma unroll
        for (int i = 0; i < 12; ++i) ir[i] = irp[i + PADL];
      }
      const int g = gc * p.GC + gl;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int ee = 8 * ec + e;          // shift index: d = dmax - ee, bin k = D-1-ee
        if (ee >= p.D) break;
        float4 v = make_float4(acc[e][0] * il[0] * ir[e], acc[e][1] * il[1] * ir[e + 1], acc[e][2] * il[2] * ir[e + 2],
                               acc[e][3] * il[3] * ir[e + 3]);
        __stcs(reinterpret_cast<float4*>(p.out + ((((size_t)b * p.G + g) * p.D + (p.D - 1 - ee)) * p.H + y) * p.W + x), v);
      }
    }
    __syncthreads();      // every thread is done with this stage (and with `inv`) before it is refilled
  }
}

int make_row_tmap(CUtensorMap* tm, const float* base, int W, int H, long long BC, int box_w, int box_c) {
  ss_encode_tiled_fn enc = ss_get_encode_tiled();
  if (!enc) return SS_ERR_CUDA;
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)BC};
  cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_w, 1u, (cuuint32_t)box_c};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ss_set_error("cuTensorMapEncodeTiled(rows) failed with CUresult %d (W=%d H=%d BC=%lld box=%dx%d)", (int)r, W, H, BC, box_w, box_c);
    return SS_ERR_CUDA;
  }
  return SS_OK;
}

}  // namespace

// Returns SS_OK when the TMA path ran, 1 when the shape is not eligible (caller falls through to the generic kernel), <0 on error.
int ss_gwc_volume_tma(const float* left, const float* right, float* volume, int B, int C, int H, int W, int maxdisp, int num_groups,
                      int flags, cudaStream_t stream) {
  const bool sgn = flags & 1;
  GwcT p;
  p.out = volume;
  p.B = B; p.C = C; p.H = H; p.W = W; p.G = num_groups; p.cg = C / num_groups;
  p.D = sgn ? 2 * maxdisp : maxdisp;
  p.dmax = maxdisp - 1;
  p.norm = (flags >> 1) & 1;
  p.E8 = ceil_div(p.D, 8) * 8;
  if (W % 4 || (reinterpret_cast<uintptr_t>(left) & 15) || (reinterpret_cast<uintptr_t>(right) & 15) ||
      (reinterpret_cast<uintptr_t>(volume) & 15))
    return 1;
  int TX = W <= 32 ? 32 : (W <= 64 ? 64 : 128);
  while (TX > 32 && TX + p.E8 + 4 > 256) TX >>= 1;
  if (TX + p.E8 + 4 > 256) return 1;
  const int padl = (4 - (p.dmax & 3)) & 3;     // x0 is a multiple of 4, so (x0 - dmax - padl) % 4 == 0
  auto stage_bytes = [&](int gcount) {
    size_t b = (size_t)gcount * p.cg * (2 * TX + p.E8 + 4) * sizeof(float);
    return (b + 127) / 128 * 128;
  };
  if (p.cg > 256 || 2 * stage_bytes(1) > 160 * 1024) return 1;
  int GC = 1;     // ~16-24 KB per stage keeps 4+ CTAs resident per SM
  for (int g = 1; g <= num_groups; ++g)
    if (num_groups % g == 0 && g * p.cg <= 256 && stage_bytes(g) <= 24 * 1024) GC = g;
  p.GC = GC; p.TX = TX; p.RW = TX + p.E8 + 4;
  p.n_xt = ceil_div(W, TX); p.n_gc = num_groups / GC;
  p.items = (long long)p.n_xt * p.n_gc * H * B;
  p.stage_bytes = (uint32_t)stage_bytes(GC);
  const size_t smem = 2 * (size_t)p.stage_bytes + (size_t)GC * (TX + p.RW) * sizeof(float);
  CUtensorMap tmL, tmR;
  int rc = make_row_tmap(&tmL, left, W, H, (long long)B * C, TX, GC * p.cg);
  if (rc != SS_OK) return rc;
  rc = make_row_tmap(&tmR, right, W, H, (long long)B * C, p.RW, GC * p.cg);
  if (rc != SS_OK) return rc;
  constexpr int THREADS = 128;
  void (*k)(const CUtensorMap, const CUtensorMap, const GwcT) =
      padl == 0 ? gwc_volume_tma_kernel<THREADS, 0> : padl == 1 ? gwc_volume_tma_kernel<THREADS, 1>
      : padl == 2 ? gwc_volume_tma_kernel<THREADS, 2> : gwc_volume_tma_kernel<THREADS, 3>;
  SS_CUDA(ss_allow_smem(k, smem));
  int per_sm = (int)((200 * 1024) / (smem + 1024));
  if (per_sm > 12) per_sm = 12;
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)ss_num_sms() * per_sm;
  if (grid > p.items) grid = p.items;
  k<<<(unsigned)grid, THREADS, smem, stream>>>(tmL, tmR, p);
  SS_CHECK_LAUNCH("ss_gwc_volume(tma)");
  return SS_OK;
}
