// K4, tensor-core mode: Conv3d k3 s1 p1 (+ folded eval-BatchNorm, ReLU, channelAtt gate) as an implicit GEMM on the
// 5th-generation tensor cores: TMA -> shared memory -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld epilogue.
// Reference layers: convbn_3d / BasicConv(is_3d) k3 s1 (models/submodule_other.py:845-848, models/submodule.py:89-116)
// as instantiated at models/SemStereo.py:109-119, 228-236 (concat_stem, classif*.0, hourglass*.conv2/conv4).
//
// Data layout ("blocked channels", bf16): activations are [B][C/8][D][H][W][8]: the 8 channels of a chunk are the 16 bytes
// one UMMA core-matrix row needs, and voxels that are neighbours along W are neighbours in memory.  A (TH+2)x(TW+2) halo
// tile of one depth slice therefore lands in shared memory (one TMA box, zero-filled outside the volume = the conv padding)
// as [C/8][TH+2][TW+2][16 B], which IS the canonical K-major no-swizzle UMMA layout for every one of the 9 in-plane taps:
// a tap (kh,kw) is just a different start address (SBO = (TW+2)*16 B between 8-voxel row groups, LBO = chunk pitch).
// So an input slice is fetched ONCE per output tile column and reused by 9 taps x 3 output slices (depth sliding ring),
// instead of 27 shifted re-loads.  GEMM tile: M = 128 voxels (16 h x 8 w of one depth slice), N = Cout tile, K = 27*Cin.
//
// Warp roles (256 threads, 1 CTA/SM, persistent over work items): warp 0 = TMA producer of input slices, warp 3 = weight
// producer (resident: all 27 taps once; streamed: per-tap ring), warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator,
// warps 4-7 = epilogue (TMEM -> registers -> scale/shift/ReLU/gate -> bf16 blocked or fp32 NCDHW stores).
// Accumulators are double-buffered in TMEM so the epilogue of slice d overlaps the MMAs of slice d+1.
#include "tc_common.cuh"

namespace {

constexpr int TH = 16, TW = 8, HH = TH + 2, WW = TW + 2;

struct TcP {
  const __nv_bfloat16* w;   // [n_tiles][27][CIN/8][N][8]
  const float* scale;       // [Cout] or null
  const float* shift;       // [Cout] or null
  const float* gate;        // (B,Cout,H,W) logits or null
  void* out;
  int out_f32;              // 0: bf16 blocked, 1: fp32 NCDHW
  int B, D, H, W, Cout, relu;
  int n_tiles, HT, WT, DC, n_dc, items;   // items = B*HT*WT*n_dc per n-tile
};

template <int CIN, int N, int NS, int NWS>
__global__ void __launch_bounds__(256, 1) conv3d_tc_s1_kernel(const __grid_constant__ CUtensorMap tmA, const TcP p) {
  constexpr bool kResident = (NWS == 27);
  constexpr uint32_t SLICE = (CIN / 8) * HH * WW * 16;     // bytes of one staged input slice
  constexpr uint32_t TAPB = CIN * N * 2;                   // bytes of one weight tap
  constexpr int KS = CIN / 16;                             // UMMA_K = 16 steps per tap
  constexpr uint32_t LBO_A = HH * WW * 16, SBO_A = WW * 16, LBO_B = N * 16, SBO_B = 128;
  constexpr uint32_t TMEM_COLS = (2 * N <= 32) ? 32 : (2 * N <= 64) ? 64 : (2 * N <= 128) ? 128 : (2 * N <= 256) ? 256 : 512;
  constexpr uint32_t IDESC = tc::make_idesc_bf16(128, N);

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* Abase = smem;
  uint8_t* Wbase = smem + NS * SLICE;
  __shared__ __align__(8) uint64_t a_full[NS], a_empty[NS], w_full[NWS], w_empty[NWS], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_scale[N], s_shift[N];     // folded BatchNorm of this CTA's Cout tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = blockIdx.x % p.n_tiles;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    s_scale[i] = p.scale ? __ldg(p.scale + nt * N + i) : 1.0f;
    s_shift[i] = p.shift ? __ldg(p.shift + nt * N + i) : 0.0f;
  }
  const int cta_s = blockIdx.x / p.n_tiles, cta_stride = gridDim.x / p.n_tiles;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmA);
    for (int i = 0; i < NS; ++i) { tc::mbar_init(&a_full[i], 1); tc::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < (kResident ? 1 : NWS); ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 128); }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  // work item s -> (b, h-tile, w-tile, depth chunk)
  auto decode = [&](int s, int& b, int& h0, int& w0, int& dlo, int& dhi) {
    const int dc = s % p.n_dc;  s /= p.n_dc;
    const int wt = s % p.WT;    s /= p.WT;
    const int ht = s % p.HT;
    b = s / p.HT;
    h0 = ht * TH; w0 = wt * TW;
    dlo = dc * p.DC; dhi = min(p.D, dlo + p.DC);
  };

  if (warp == 0 && lane == 0) {
    // ===== input-slice producer =====
    uint32_t g = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode(s, b, h0, w0, dlo, dhi);
      const int din0 = max(dlo - 1, 0), din1 = min(dhi, p.D - 1);
      for (int d_in = din0; d_in <= din1; ++d_in, ++g) {
        const uint32_t slot = g % NS;
        tc::mbar_wait(&a_empty[slot], ((g / NS) & 1) ^ 1);
        tc::mbar_expect_tx(&a_full[slot], SLICE);
        tc::tma_load_4d(Abase + slot * SLICE, &tmA, &a_full[slot], (w0 - 1) * 8, h0 - 1, d_in, b * (CIN / 8));
      }
    }
  } else if (warp == 3 && lane == 0) {
    // ===== weight producer =====
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
        decode(s, b, h0, w0, dlo, dhi);
        for (int d_out = dlo; d_out < dhi; ++d_out)
          for (int kd = 0; kd < 3; ++kd) {
            const int d_in = d_out + kd - 1;
            if (d_in < 0 || d_in >= p.D) continue;
            for (int t9 = 0; t9 < 9; ++t9, ++wc) {
              const uint32_t slot = wc % NWS;
              tc::mbar_wait(&w_empty[slot], ((wc / NWS) & 1) ^ 1);
              tc::mbar_expect_tx(&w_full[slot], TAPB);
              tc::bulk_load(Wbase + slot * TAPB, wsrc + (size_t)(kd * 9 + t9) * TAPB, TAPB, &w_full[slot]);
            }
          }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the (uniform) control flow so that descriptors live in uniform registers;
    //       one elected lane issues the tcgen05 instructions =====
    const bool leader = tc::elect_one();
    const uint32_t a_lo0 = tc::desc_lo(tc::smem_u32(Abase), LBO_A), a_hi = tc::desc_hi(SBO_A);
    const uint32_t b_lo0 = tc::desc_lo(tc::smem_u32(Wbase), LBO_B), b_hi = tc::desc_hi(SBO_B);
    if (kResident && cta_s < p.items) tc::mbar_wait(&w_full[0], 0);
    uint32_t g_base = 0, acc_it = 0, wc = 0;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode(s, b, h0, w0, dlo, dhi);
      const int din0 = max(dlo - 1, 0), din1 = min(dhi, p.D - 1);
      for (int d_out = dlo; d_out < dhi; ++d_out, ++acc_it) {
        const uint32_t as = acc_it & 1;
        tc::mbar_wait(&acc_empty[as], ((acc_it >> 1) & 1) ^ 1);
        tc::fence_after_sync();
        const uint32_t tmem_d = tmem_base + as * N;
        uint32_t accumulate = 0;
#pragma unroll 1
        for (int kd = 0; kd < 3; ++kd) {
          const int d_in = d_out + kd - 1;
          if (d_in < 0 || d_in >= p.D) continue;
          const uint32_t gs = g_base + (uint32_t)(d_in - din0), slot = gs % NS;
          tc::mbar_wait(&a_full[slot], (gs / NS) & 1);
          tc::fence_after_sync();
          const uint32_t a_lo = a_lo0 + slot * (SLICE >> 4);
#pragma unroll
          for (int t9 = 0; t9 < 9; ++t9) {
            const int kh = t9 / 3, kw = t9 - 3 * kh;
            uint32_t b_lo;
            uint32_t wslot = 0;
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
          // input slices no output slice of this item needs any more
          if (d_out - 1 >= din0) tc::mma_commit(&a_empty[(g_base + (uint32_t)(d_out - 1 - din0)) % NS]);
          if (d_out == dhi - 1) {
            tc::mma_commit(&a_empty[(g_base + (uint32_t)(d_out - din0)) % NS]);
            if (d_out + 1 <= din1) tc::mma_commit(&a_empty[(g_base + (uint32_t)(d_out + 1 - din0)) % NS]);
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
    const size_t HWs = (size_t)p.H * p.W;
    for (int s = cta_s; s < p.items; s += cta_stride) {
      int b, h0, w0, dlo, dhi;
      decode(s, b, h0, w0, dlo, dhi);
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
          if (!valid) continue;
          const int co0 = nt * N + j * 32;
          const float lo = p.relu ? 0.0f : -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(fmaf(v[i], s_scale[j * 32 + i], s_shift[j * 32 + i]), lo);
          if (p.gate) {                     // branch hoisted out of the channel loop: 32 independent loads in flight
            const float* gp = p.gate + ((size_t)b * p.Cout + co0) * HWs + (size_t)h * p.W + w;
            float gl[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) gl[i] = __ldg(gp + (size_t)i * HWs);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= sigmoidf_(gl[i]);
          }
          if (p.out_f32) {
            float* o = reinterpret_cast<float*>(p.out) + (((size_t)b * p.Cout + co0) * p.D + d_out) * HWs + (size_t)h * p.W + w;
#pragma unroll
            for (int i = 0; i < 32; ++i) o[(size_t)i * p.D * HWs] = v[i];
          } else {
            uint4* o = reinterpret_cast<uint4*>(p.out) + (((size_t)b * (p.Cout / 8) + co0 / 8) * p.D + d_out) * HWs + (size_t)h * p.W + w;
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              uint4 q;
              q.x = tc::pack_bf16x2(v[8 * c8 + 0], v[8 * c8 + 1]);
              q.y = tc::pack_bf16x2(v[8 * c8 + 2], v[8 * c8 + 3]);
              q.z = tc::pack_bf16x2(v[8 * c8 + 4], v[8 * c8 + 5]);
              q.w = tc::pack_bf16x2(v[8 * c8 + 6], v[8 * c8 + 7]);
              o[(size_t)c8 * p.D * HWs] = q;
            }
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---- layout converters: fp32 NCDHW <-> bf16 blocked [B][C/8][D][H][W][8] -------------------------------------------
__global__ void __launch_bounds__(256) to_blocked_kernel(const float* __restrict__ in, uint4* __restrict__ out, int C, size_t S) {
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
  out[((size_t)b * (C / 8) + chunk) * S + v] = q;
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

int make_act_tmap(CUtensorMap* tm, const void* base, int B, int C, int D, int H, int W) {
  ss_encode_tiled_fn enc = ss_get_encode_tiled();
  if (!enc) return SS_ERR_CUDA;
  cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B * (C / 8)};
  cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
  cuuint32_t box[4] = {(cuuint32_t)WW * 8, (cuuint32_t)HH, 1u, (cuuint32_t)(C / 8)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ss_set_error("cuTensorMapEncodeTiled failed with CUresult %d (B=%d C=%d D=%d H=%d W=%d)", (int)r, B, C, D, H, W);
    return SS_ERR_CUDA;
  }
  return SS_OK;
}

template <int CIN, int N, int NS, int NWS>
int launch_s1(const CUtensorMap& tm, TcP p, cudaStream_t st) {
  constexpr size_t smem = (size_t)NS * (CIN / 8) * HH * WW * 16 + (size_t)NWS * CIN * N * 2;
  static_assert(smem <= 227 * 1024 - 2048, "shared memory budget");
  auto k = conv3d_tc_s1_kernel<CIN, N, NS, NWS>;
  SS_CUDA(ss_allow_smem(k, smem));
  int grid = ss_num_sms();
  grid -= grid % p.n_tiles;
  // depth chunking: balance waves against the 2 extra halo slices every chunk re-reads
  const int spatial = p.B * p.HT * p.WT;
  int best = p.D;
  double best_cost = 1e30;
  for (int dc = 2; dc <= p.D; ++dc) {
    if (p.D % dc && dc != p.D) continue;
    const long long items = (long long)spatial * ceil_div(p.D, dc) * p.n_tiles;
    const double cost = (double)ceil_div64(items, grid) * (dc + 0.35);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = dc; }
  }
  if (p.D < 2) best = p.D;
  p.DC = best;
  p.n_dc = ceil_div(p.D, best);
  p.items = spatial * p.n_dc;
  const long long total = (long long)p.items * p.n_tiles;
  if (total < grid) { grid = (int)total; grid -= grid % p.n_tiles; if (grid < p.n_tiles) grid = p.n_tiles; }
  k<<<grid, 256, smem, st>>>(tm, p);
  SS_CHECK_LAUNCH("ss_conv3d_tc");
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

// Cout tile the kernel uses for (Cin, Cout): the weight must be packed as [Cout/N][27][Cin/8][N][8] bf16. 0 = unsupported.
extern "C" int ss_conv3d_tc_ntile(int Cin, int Cout) {
  if ((Cin == 32 || Cin == 64) && (Cout == 32 || Cout == 64)) return 32;
  if (Cin == 128 && Cout == 128) return 128;
  return 0;
}

extern "C" int ss_conv3d_tc(const void* in_blocked, const void* weight_packed, const float* scale_or_null, const float* shift_or_null,
                            const float* gate_logits_or_null, void* out, int out_is_f32, int B, int Cin, int Cout, int D, int H,
                            int W, int relu, void* stream) {
  SS_REQUIRE(in_blocked && weight_packed && out, "ss_conv3d_tc: null pointer");
  SS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "ss_conv3d_tc: non-positive dimension");
  const int N = ss_conv3d_tc_ntile(Cin, Cout);
  SS_UNSUPPORTED(N == 0, "ss_conv3d_tc: (Cin=%d, Cout=%d) has no tensor-core configuration", Cin, Cout);
  SS_REQUIRE((reinterpret_cast<uintptr_t>(in_blocked) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(weight_packed) & 15) == 0, "ss_conv3d_tc: pointers must be 16-byte aligned");
  SS_UNSUPPORTED((long long)B * (Cin / 8) > 0x7fffffffLL, "ss_conv3d_tc: too many channel chunks");
  CUtensorMap tm;
  int rc = make_act_tmap(&tm, in_blocked, B, Cin, D, H, W);
  if (rc != SS_OK) return rc;
  TcP p;
  p.w = reinterpret_cast<const __nv_bfloat16*>(weight_packed);
  p.scale = scale_or_null; p.shift = shift_or_null; p.gate = gate_logits_or_null;
  p.out = out; p.out_f32 = out_is_f32;
  p.B = B; p.D = D; p.H = H; p.W = W; p.Cout = Cout; p.relu = relu;
  p.n_tiles = Cout / N; p.HT = ceil_div(H, TH); p.WT = ceil_div(W, TW);
  p.DC = D; p.n_dc = 1; p.items = 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 32) return launch_s1<32, 32, 4, 27>(tm, p, st);
  if (Cin == 64) return launch_s1<64, 32, 4, 27>(tm, p, st);
  return launch_s1<128, 128, 3, 2>(tm, p, st);
}

extern "C" int ss_to_blocked_bf16(const float* in_ncdhw, void* out_blocked, int B, int C, int D, int H, int W, void* stream) {
  SS_REQUIRE(in_ncdhw && out_blocked && B > 0 && C > 0 && D > 0 && H > 0 && W > 0, "ss_to_blocked_bf16: bad argument");
  SS_REQUIRE(C % 8 == 0, "ss_to_blocked_bf16: C=%d must be a multiple of 8", C);
  SS_UNSUPPORTED(C / 8 > 65535 || B > 65535, "ss_to_blocked_bf16: grid dimension exceeds 65535");
  const size_t S = (size_t)D * H * W;
  to_blocked_kernel<<<dim3((unsigned)ceil_div64(S, 256), C / 8, B), 256, 0, (cudaStream_t)stream>>>(
      in_ncdhw, reinterpret_cast<uint4*>(out_blocked), C, S);
  SS_CHECK_LAUNCH("ss_to_blocked_bf16");
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
