// K1 / K2: cost-volume builders.
//   gwc_volume     : build_gwc_volume, build_gwc_volume_norm, build_norm_correlation_volume
//                    (reference models/submodule.py:190-255, models/submodule_.py:180-237)
//   concat_volume  : build_concat_volume (models/submodule.py:173-187, models/submodule_.py:166-178)
//
// Both are HBM-bound.  One CTA stages the left/right feature rows of one image row (a chunk of
// channel groups, one x-tile) in shared memory ONCE; the right row carries a zero halo of D-1
// columns, so every disparity shift is a plain shared-memory offset and the out-of-range region
// of the volume comes out as exact zeros without branches.  Each thread owns a 4(x) x 8(shift)
// register tile: one 128-bit load of the left row and three of the right row feed 32 FMAs.
#include "common.cuh"

// volumes_tma.cu: 0 = done, 1 = shape not eligible for the TMA path, <0 = error
int ss_gwc_volume_tma(const float* left, const float* right, float* volume, int B, int C, int H, int W, int maxdisp, int num_groups,
                      int flags, cudaStream_t stream);

namespace {

struct GwcParams {
  const float* L;
  const float* R;
  float* out;
  int B, C, H, W, G, cg, D, dmax, norm;
  int GC;   // groups per CTA
  int TX;   // x-tile width (multiple of 4)
  int E8;   // D rounded up to a multiple of 8
  int RW;   // staged right-row width = TX + E8
  int n_xt, n_gc;
  int vec4;  // W % 4 == 0 -> 128-bit stores
};

__global__ void __launch_bounds__(256) gwc_volume_kernel(const GwcParams p) {
  extern __shared__ __align__(16) float smem[];
  const int CC = p.GC * p.cg;
  float* Ls = smem;                 // [CC][TX]
  float* Rs = smem + CC * p.TX;     // [CC][RW]   Rs[c][j] = R[c][x0 - dmax + j] (0 outside the row)

  int bid = blockIdx.x;
  const int xt = bid % p.n_xt;  bid /= p.n_xt;
  const int gc = bid % p.n_gc;  bid /= p.n_gc;
  const int y = bid % p.H;
  const int b = bid / p.H;
  const int x0 = xt * p.TX;
  const int c0 = gc * CC;
  const size_t row_stride = (size_t)p.H * p.W;
  const float* Lrow = p.L + ((size_t)b * p.C + c0) * row_stride + (size_t)y * p.W;
  const float* Rrow = p.R + ((size_t)b * p.C + c0) * row_stride + (size_t)y * p.W;

  // ---- stage rows (coalesced along x) ------------------------------------------------
  for (int i = threadIdx.x; i < CC * p.TX; i += blockDim.x) {
    int c = i / p.TX, j = i - c * p.TX;
    int x = x0 + j;
    Ls[i] = (x < p.W) ? __ldg(Lrow + (size_t)c * row_stride + x) : 0.0f;
  }
  for (int i = threadIdx.x; i < CC * p.RW; i += blockDim.x) {
    int c = i / p.RW, j = i - c * p.RW;
    int x = x0 - p.dmax + j;
    Rs[i] = (x >= 0 && x < p.W) ? __ldg(Rrow + (size_t)c * row_stride + x) : 0.0f;
  }
  __syncthreads();

  // ---- per-group L2 normalisation, once per staged column (groupwise_correlation_norm) ----
  if (p.norm) {
    for (int i = threadIdx.x; i < p.GC * (p.TX + p.RW); i += blockDim.x) {
      int g = i / (p.TX + p.RW), j = i - g * (p.TX + p.RW);
      float* col;
      int stride;
      if (j < p.TX) { col = Ls + (size_t)g * p.cg * p.TX + j; stride = p.TX; }
      else          { col = Rs + (size_t)g * p.cg * p.RW + (j - p.TX); stride = p.RW; }
      float ss = 0.0f;
      for (int c = 0; c < p.cg; ++c) { float v = col[c * stride]; ss = fmaf(v, v, ss); }
      float den = sqrtf(ss) + 1e-5f;          // eps is added to the norm (submodule.py:218)
      for (int c = 0; c < p.cg; ++c) col[c * stride] = col[c * stride] / den;
    }
    __syncthreads();
  }

  // ---- correlation: thread tile = 4 x-columns x 8 shifts -------------------------------
  const int nxq = p.TX >> 2, nec = p.E8 >> 3;
  const int ntiles = p.GC * nec * nxq;
  const float denom = (float)p.cg;   // mean over the channel group = sum / cg (ATen mean)
  for (int t = threadIdx.x; t < ntiles; t += blockDim.x) {
    const int xq = t % nxq;
    const int ec = (t / nxq) % nec;
    const int gl = t / (nxq * nec);
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
      const float l[4] = {l4.x, l4.y, l4.z, l4.w};
      const float r[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
#pragma unroll
      for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[e][i] = fmaf(l[i], r[e + i], acc[e][i]);
    }
    const int x = x0 + 4 * xq;
    if (x >= p.W) continue;
    const int g = gc * p.GC + gl;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ee = 8 * ec + e;          // shift index: d = dmax - ee
      if (ee >= p.D) break;
      const int k = p.D - 1 - ee;         // bin index of disparity d (bins ascend with d)
      float* o = p.out + ((((size_t)b * p.G + g) * p.D + k) * p.H + y) * p.W + x;
      if (p.vec4) {
        float4 v = make_float4(acc[e][0] / denom, acc[e][1] / denom, acc[e][2] / denom, acc[e][3] / denom);
        __stcs(reinterpret_cast<float4*>(o), v);      // streaming store: the volume is write-once
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (x + i < p.W) o[i] = acc[e][i] / denom;
      }
    }
  }
}

struct ConcatParams {
  const float* L;
  const float* R;
  float* out;
  int B, C, H, W, D, dmin, mask_left;
};

// One CTA per (b, channel of the 2C, y): the source row is staged once, D shifted copies are streamed out.
__global__ void __launch_bounds__(128) concat_volume_kernel(const ConcatParams p) {
  extern __shared__ __align__(16) float row[];   // [W]
  int bid = blockIdx.x;
  const int y = bid % p.H;  bid /= p.H;
  const int c2 = bid % (2 * p.C);
  const int b = bid / (2 * p.C);
  const bool left = c2 < p.C;
  const float* src = (left ? p.L : p.R) + (((size_t)b * p.C + (left ? c2 : c2 - p.C)) * p.H + y) * p.W;
  for (int x = threadIdx.x; x < p.W; x += blockDim.x) row[x] = __ldg(src + x);
  __syncthreads();
  float* obase = p.out + ((size_t)b * 2 * p.C + c2) * p.D * p.H * p.W + (size_t)y * p.W;
  const size_t kstride = (size_t)p.H * p.W;
  const bool vec4 = (p.W & 3) == 0;
  const int nq = (p.W + 3) >> 2;
  for (int i = threadIdx.x; i < p.D * nq; i += blockDim.x) {
    const int k = i / nq, xq = i - k * nq;
    const int d = p.dmin + k;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = 4 * xq + j;
      const int xs = x - d;
      const bool ok = xs >= 0 && xs < p.W && x < p.W;
      if (left) v[j] = (x < p.W && (ok || !p.mask_left)) ? row[x] : 0.0f;
      else      v[j] = ok ? row[xs] : 0.0f;
    }
    float* o = obase + k * kstride + 4 * xq;
    if (vec4) __stcs(reinterpret_cast<float4*>(o), make_float4(v[0], v[1], v[2], v[3]));
    else
      for (int j = 0; j < 4; ++j)
        if (4 * xq + j < p.W) o[j] = v[j];
  }
}

}  // namespace

// flags: bit0 = signed disparity range (-M..M-1, depth 2M) else unsigned (0..M-1, depth M); bit1 = normalise
extern "C" int ss_gwc_volume(const float* left, const float* right, float* volume, int B, int C, int H, int W,
                             int maxdisp, int num_groups, int flags, void* stream) {
  SS_REQUIRE(left && right && volume, "ss_gwc_volume: null pointer");
  SS_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && maxdisp > 0 && num_groups > 0, "ss_gwc_volume: non-positive dimension");
  SS_REQUIRE(C % num_groups == 0, "ss_gwc_volume: C=%d not divisible by num_groups=%d", C, num_groups);
  {  // TMA-staged fast path (volumes_tma.cu) for 16-byte-aligned rows; everything else takes the generic kernel below
    const int rc = ss_gwc_volume_tma(left, right, volume, B, C, H, W, maxdisp, num_groups, flags, (cudaStream_t)stream);
    if (rc <= 0) return rc;
  }
  const bool sgn = flags & 1;
  GwcParams p;
  p.L = left; p.R = right; p.out = volume;
  p.B = B; p.C = C; p.H = H; p.W = W; p.G = num_groups; p.cg = C / num_groups;
  p.D = sgn ? 2 * maxdisp : maxdisp;
  p.dmax = maxdisp - 1;
  p.norm = (flags >> 1) & 1;
  p.E8 = ceil_div(p.D, 8) * 8;
  p.vec4 = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(volume) & 15) == 0);
  // x-tile: as wide as the row allows (<=128); shrink while one group does not fit in shared memory
  int TX = W <= 32 ? 32 : (W <= 64 ? 64 : 128);
  const size_t kMaxSmem = 200 * 1024, kTarget = 72 * 1024;
  auto bytes = [&](int gcount, int tx) { return (size_t)gcount * p.cg * (2 * tx + p.E8) * sizeof(float); };
  while (TX > 4 && bytes(1, TX) > kMaxSmem) TX >>= 1;
  SS_UNSUPPORTED(bytes(1, TX) > kMaxSmem, "ss_gwc_volume: channels-per-group=%d with depth %d does not fit shared memory", p.cg, p.D);
  int GC = 1;
  for (int g = 1; g <= num_groups; ++g)
    if (num_groups % g == 0 && bytes(g, TX) <= kTarget) GC = g;
  p.GC = GC; p.TX = TX; p.RW = TX + p.E8;
  p.n_xt = ceil_div(W, TX); p.n_gc = num_groups / GC;
  const size_t smem = bytes(GC, TX);
  const int64_t nblk = (int64_t)p.n_xt * p.n_gc * H * B;
  SS_UNSUPPORTED(nblk > 0x7fffffffLL, "ss_gwc_volume: problem too large for one launch");
  SS_CUDA(ss_allow_smem(gwc_volume_kernel, smem));
  gwc_volume_kernel<<<(unsigned)nblk, 256, smem, (cudaStream_t)stream>>>(p);
  SS_CHECK_LAUNCH("ss_gwc_volume");
  return SS_OK;
}

extern "C" int ss_concat_volume(const float* left, const float* right, float* volume, int B, int C, int H, int W,
                                int maxdisp, int flags, void* stream) {
  SS_REQUIRE(left && right && volume, "ss_concat_volume: null pointer");
  SS_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && maxdisp > 0, "ss_concat_volume: non-positive dimension");
  const bool sgn = flags & 1;
  ConcatParams p;
  p.L = left; p.R = right; p.out = volume;
  p.B = B; p.C = C; p.H = H; p.W = W;
  p.D = sgn ? 2 * maxdisp : maxdisp;
  p.dmin = sgn ? -maxdisp : 0;
  p.mask_left = sgn ? 1 : 0;   // the unsigned variant leaves the left half unmasked (submodule_.py:171)
  const int64_t nblk = (int64_t)B * 2 * C * H;
  SS_UNSUPPORTED(nblk > 0x7fffffffLL, "ss_concat_volume: problem too large for one launch");
  SS_UNSUPPORTED((size_t)W * 4 > 160 * 1024, "ss_concat_volume: row too wide");
  SS_REQUIRE((reinterpret_cast<uintptr_t>(volume) & 15) == 0 || (W & 3), "ss_concat_volume: volume must be 16-byte aligned");
  SS_CUDA(ss_allow_smem(concat_volume_kernel, (size_t)W * 4));
  concat_volume_kernel<<<(unsigned)nblk, 128, (size_t)W * 4, (cudaStream_t)stream>>>(p);
  SS_CHECK_LAUNCH("ss_concat_volume");
  return SS_OK;
}
