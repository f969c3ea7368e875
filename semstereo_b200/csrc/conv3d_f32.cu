// K4, fp32-accurate mode: 3-D convolutions of the hourglass stack as an implicit GEMM on the FP32 pipe.
// This is the precision-reference mode that keeps the attention branch's top-k indices bit-comparable
// with the fp32 reference; the tensor-core (tcgen05, bf16) mode lives in conv3d_tc.cu.
//
// Covers every 3-D layer type on the path (reference models/SemStereo.py:106-182, 228-236;
// convbn_3d models/submodule_other.py:845-848; BasicConv is_3d models/submodule.py:89-116):
//   Conv3d k3 s1 p1, Conv3d k3 s2 p1, Conv3d k1, ConvTranspose3d k3 s2 p1 op1 (as 8 sub-pixel phase GEMMs,
//   no zero insertion), and the Cout=1 classifier head (dedicated kernel).
// Epilogue: y = acc*scale[co] + shift[co] (eval BatchNorm / bias folded), + residual, ReLU, * sigmoid(gate[b,co,h,w]).
//
// GEMM view: M = output voxels (tile 128, w fastest), N = Cout (tile 32 or 64), K = taps x Cin (chunks of 16 Cin per tap).
// A is gathered on the fly (coalesced along w), B is the pre-packed weight [tap][Cin][Cout].
#include "common.cuh"

namespace {

struct ConvP {
  const float* in;
  const float* w;         // [taps][Cin][Cout]
  const float* scale;     // [Cout] or null
  const float* shift;     // [Cout] or null
  const float* residual;  // like out, or null
  const float* gate;      // (B,Cout,Ho,Wo) logits, or null
  float* out;
  int B, Cin, Cout, Di, Hi, Wi, Do, Ho, Wo;
  int K, stride, pad, transposed, relu;
  int Mo_d, Mo_h, Mo_w;   // extent of the M index space per dim (= Do,Ho,Wo; halved for the transposed phases)
  long long M;            // B*Mo_d*Mo_h*Mo_w
};

constexpr int BM = 128, CK = 16, NT = 256;

template <int BN>
__global__ void __launch_bounds__(NT) conv3d_igemm_f32_kernel(const ConvP p) {
  constexpr int TN = BN / 8;                  // couts per thread (warp-uniform), 4 voxels per thread
  __shared__ __align__(16) float As[2][CK][BM];
  __shared__ __align__(16) float Bs[2][CK][BN];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int co0 = blockIdx.y * BN;
  const int phase = blockIdx.z;               // transposed only: (pd,ph,pw) output parity
  const int pd = (phase >> 2) & 1, ph = (phase >> 1) & 1, pw = phase & 1;

  // ---- the voxel this thread gathers for (fixed for the whole kernel) ----
  const int gm = tid & (BM - 1);              // A-tile row gathered by this thread
  const int gk0 = tid >> 7;                   // first k row (0/1), step 2
  long long mg = (long long)blockIdx.x * BM + gm;
  const bool m_ok = mg < p.M;
  int mw = 0, mh = 0, md = 0, mb = 0;
  if (m_ok) {
    mw = (int)(mg % p.Mo_w); mg /= p.Mo_w;
    mh = (int)(mg % p.Mo_h); mg /= p.Mo_h;
    md = (int)(mg % p.Mo_d); mb = (int)(mg / p.Mo_d);
  }
  const size_t in_cs = (size_t)p.Di * p.Hi * p.Wi;     // channel stride
  const float* in_b = p.in + (size_t)mb * p.Cin * in_cs;

  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  const int ntaps = p.K * p.K * p.K;
  const int nchunk = (p.Cin + CK - 1) / CK;
  float a_reg[CK / 2];
  float b_reg[(CK * BN) / NT];

  // tap -> input coordinate of this thread's voxel; returns validity
  auto tap_src = [&](int tap, long long& off) -> bool {
    const int kd = tap / (p.K * p.K), kh = (tap / p.K) % p.K, kw = tap % p.K;
    int di, hi, wi;
    if (!p.transposed) {
      di = md * p.stride - p.pad + kd; hi = mh * p.stride - p.pad + kh; wi = mw * p.stride - p.pad + kw;
    } else {   // out o = 2*m + parity ; in i = (o + 1 - k) / 2 (parity of k already matches)
      di = md + ((pd + 1 - kd) >> 1); hi = mh + ((ph + 1 - kh) >> 1); wi = mw + ((pw + 1 - kw) >> 1);
    }
    off = ((long long)di * p.Hi + hi) * p.Wi + wi;
    return m_ok && di >= 0 && di < p.Di && hi >= 0 && hi < p.Hi && wi >= 0 && wi < p.Wi;
  };
  auto tap_live = [&](int tap) -> bool {      // CTA-uniform: transposed phases use only parity-matching taps
    if (!p.transposed) return true;
    const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
    return ((pd + 1 - kd) & 1) == 0 && ((ph + 1 - kh) & 1) == 0 && ((pw + 1 - kw) & 1) == 0;
  };
  auto load_regs = [&](int tap, int chunk) {
    long long off;
    const bool ok = tap_src(tap, off);
    const int c0 = chunk * CK;
#pragma unroll
    for (int j = 0; j < CK / 2; ++j) {
      const int c = c0 + gk0 + 2 * j;
      a_reg[j] = (ok && c < p.Cin) ? __ldg(in_b + (size_t)c * in_cs + off) : 0.0f;
    }
    const float* wt = p.w + ((size_t)tap * p.Cin + c0) * p.Cout + co0;
#pragma unroll
    for (int j = 0; j < (CK * BN) / NT; ++j) {
      const int i = tid + j * NT, k = i / BN, n = i - k * BN;
      b_reg[j] = (c0 + k < p.Cin && co0 + n < p.Cout) ? __ldg(wt + (size_t)k * p.Cout + n) : 0.0f;
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int j = 0; j < CK / 2; ++j) As[buf][gk0 + 2 * j][gm] = a_reg[j];
#pragma unroll
    for (int j = 0; j < (CK * BN) / NT; ++j) {
      const int i = tid + j * NT, k = i / BN, n = i - k * BN;
      Bs[buf][k][n] = b_reg[j];
    }
  };

  // ---- software pipeline over (tap, chunk) ----
  int tap = 0, chunk = 0;
  while (tap < ntaps && !tap_live(tap)) ++tap;
  int buf = 0;
  if (tap < ntaps) { load_regs(tap, chunk); store_smem(0); }
  __syncthreads();
  while (tap < ntaps) {
    int ntap = tap, nchk = chunk + 1;
    if (nchk == nchunk) { nchk = 0; ++ntap; while (ntap < ntaps && !tap_live(ntap)) ++ntap; }
    const bool more = ntap < ntaps;
    if (more) load_regs(ntap, nchk);          // global loads in flight during the FMAs below
#pragma unroll
    for (int k = 0; k < CK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][4 * lane]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      float bv[TN];
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][warp * TN + j]);
        bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) store_smem(buf ^ 1);
    __syncthreads();
    buf ^= 1; tap = ntap; chunk = nchk;
  }

  // ---- epilogue ----
  const size_t out_cs = (size_t)p.Do * p.Ho * p.Wo;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = (long long)blockIdx.x * BM + 4 * lane + i;
    if (m >= p.M) continue;
    int ow = (int)(m % p.Mo_w); m /= p.Mo_w;
    int oh = (int)(m % p.Mo_h); m /= p.Mo_h;
    int od = (int)(m % p.Mo_d);
    const int ob = (int)(m / p.Mo_d);
    if (p.transposed) { od = 2 * od + pd; oh = 2 * oh + ph; ow = 2 * ow + pw; }
    const size_t sp = ((size_t)od * p.Ho + oh) * p.Wo + ow;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = co0 + warp * TN + j;
      if (co >= p.Cout) continue;
      float v = acc[i][j];
      if (p.scale) v *= __ldg(p.scale + co);
      if (p.shift) v += __ldg(p.shift + co);
      const size_t o = ((size_t)ob * p.Cout + co) * out_cs + sp;
      if (p.residual) v += __ldg(p.residual + o);
      if (p.relu) v = fmaxf(v, 0.0f);
      if (p.gate) v *= sigmoidf_(__ldg(p.gate + (((size_t)ob * p.Cout + co) * p.Ho + oh) * p.Wo + ow));
      p.out[o] = v;
    }
  }
}

// Cout = 1, k3 s1 p1 (classif.2 / classif_att_.2): memory-bound; thread = 4 consecutive w, weights in smem.
__global__ void __launch_bounds__(128) conv3d_cout1_f32_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                              float* __restrict__ out, int Cin, int D, int H, int W) {
  extern __shared__ float ws[];                 // [Cin][27]
  for (int i = threadIdx.x; i < Cin * 27; i += blockDim.x) ws[i] = __ldg(w + i);
  __syncthreads();
  const int xq = blockIdx.x * blockDim.x + threadIdx.x;
  const int x0 = 4 * xq;
  if (x0 >= W) return;
  const int y = blockIdx.y % H, d = blockIdx.y / H, b = blockIdx.z;
  const size_t cs = (size_t)D * H * W;
  const float* ib = in + (size_t)b * Cin * cs;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = 0; c < Cin; ++c) {
    const float* ic = ib + (size_t)c * cs;
    const float* wc = ws + c * 27;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const int dd = d + kd - 1;
      if (dd < 0 || dd >= D) continue;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int yy = y + kh - 1;
        if (yy < 0 || yy >= H) continue;
        const float* row = ic + ((size_t)dd * H + yy) * W;
        float v[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const int xx = x0 - 1 + j;
          v[j] = (xx >= 0 && xx < W) ? __ldg(row + xx) : 0.0f;
        }
        const float w0 = wc[kd * 9 + kh * 3], w1 = wc[kd * 9 + kh * 3 + 1], w2 = wc[kd * 9 + kh * 3 + 2];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = fmaf(w0, v[i], fmaf(w1, v[i + 1], fmaf(w2, v[i + 2], acc[i])));
      }
    }
  }
  float* o = out + ((size_t)b * D + d) * H * W + (size_t)y * W + x0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (x0 + i < W) o[i] = acc[i];
}

}  // namespace

// in (B,Cin,Di,Hi,Wi) fp32; weight_packed [K^3][Cin][Cout] fp32 (tap = (kd*K + kh)*K + kw; for the transposed layer the
// tap indexes the ConvTranspose3d weight directly, no flip); out (B,Cout,Do,Ho,Wo).
// mode: 0 = Conv3d(K, stride, pad = K/2); 1 = ConvTranspose3d(k3, s2, p1, op1) (Do = 2*Di ...).
extern "C" int ss_conv3d_f32(const float* in, const float* weight_packed, const float* scale_or_null, const float* shift_or_null,
                             const float* residual_or_null, const float* gate_logits_or_null, float* out, int B, int Cin, int Cout,
                             int Di, int Hi, int Wi, int K, int stride, int mode, int relu, void* stream) {
  SS_REQUIRE(in && weight_packed && out, "ss_conv3d_f32: null pointer");
  SS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && Di > 0 && Hi > 0 && Wi > 0, "ss_conv3d_f32: non-positive dimension");
  SS_UNSUPPORTED(!(K == 1 || K == 3), "ss_conv3d_f32: kernel size %d unsupported (1 or 3)", K);
  SS_UNSUPPORTED(!(stride == 1 || stride == 2), "ss_conv3d_f32: stride %d unsupported (1 or 2)", stride);
  SS_UNSUPPORTED(mode == 1 && !(K == 3 && stride == 2), "ss_conv3d_f32: transposed mode needs k=3, stride=2");
  SS_UNSUPPORTED(mode != 0 && mode != 1, "ss_conv3d_f32: unknown mode %d", mode);
  ConvP p;
  p.in = in; p.w = weight_packed; p.scale = scale_or_null; p.shift = shift_or_null;
  p.residual = residual_or_null; p.gate = gate_logits_or_null; p.out = out;
  p.B = B; p.Cin = Cin; p.Cout = Cout; p.Di = Di; p.Hi = Hi; p.Wi = Wi;
  p.K = K; p.stride = stride; p.pad = K / 2; p.transposed = mode; p.relu = relu;
  if (mode == 0) {
    p.Do = (Di + 2 * p.pad - K) / stride + 1; p.Ho = (Hi + 2 * p.pad - K) / stride + 1; p.Wo = (Wi + 2 * p.pad - K) / stride + 1;
    p.Mo_d = p.Do; p.Mo_h = p.Ho; p.Mo_w = p.Wo;
  } else {
    p.Do = 2 * Di; p.Ho = 2 * Hi; p.Wo = 2 * Wi;
    p.Mo_d = Di; p.Mo_h = Hi; p.Mo_w = Wi;
  }
  p.M = (long long)B * p.Mo_d * p.Mo_h * p.Mo_w;
  const long long mt = (p.M + BM - 1) / BM;
  SS_UNSUPPORTED(mt > 0x7fffffffLL, "ss_conv3d_f32: problem too large for one launch");
  const int phases = mode ? 8 : 1;
  if (Cout >= 64) {
    dim3 grid((unsigned)mt, ceil_div(Cout, 64), phases);
    conv3d_igemm_f32_kernel<64><<<grid, NT, 0, (cudaStream_t)stream>>>(p);
  } else {
    dim3 grid((unsigned)mt, ceil_div(Cout, 32), phases);
    conv3d_igemm_f32_kernel<32><<<grid, NT, 0, (cudaStream_t)stream>>>(p);
  }
  SS_CHECK_LAUNCH("ss_conv3d_f32");
  return SS_OK;
}

// weight (1,Cin,3,3,3) in PyTorch layout; out (B,1,D,H,W)
extern "C" int ss_conv3d_cout1_f32(const float* in, const float* weight, float* out, int B, int Cin, int D, int H, int W, void* stream) {
  SS_REQUIRE(in && weight && out, "ss_conv3d_cout1_f32: null pointer");
  SS_REQUIRE(B > 0 && Cin > 0 && D > 0 && H > 0 && W > 0, "ss_conv3d_cout1_f32: non-positive dimension");
  SS_UNSUPPORTED((long long)D * H > 65535 || B > 65535, "ss_conv3d_cout1_f32: grid dimension exceeds 65535");
  SS_UNSUPPORTED(Cin * 27 * 4 > 48 * 1024, "ss_conv3d_cout1_f32: Cin=%d too large", Cin);
  dim3 grid(ceil_div(ceil_div(W, 4), 128), D * H, B);
  conv3d_cout1_f32_kernel<<<grid, 128, Cin * 27 * sizeof(float), (cudaStream_t)stream>>>(in, weight, out, Cin, D, H, W);
  SS_CHECK_LAUNCH("ss_conv3d_cout1_f32");
  return SS_OK;
}
