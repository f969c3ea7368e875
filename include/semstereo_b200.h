/* semstereo_b200.h — C-ABI of the B200-native disparity hot path of SemStereo.
 *
 * Every entry point replaces one function / module of the reference's Python operator surface
 * (cited per function as file:line under /root/reference).  Conventions:
 *   - all pointers are DEVICE pointers to contiguous fp32 NCHW / NCDHW tensors unless stated otherwise;
 *   - the library never allocates, frees or retains caller memory; outputs are fully overwritten;
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream),
 *     never synchronises, and is CUDA-graph capturable; entry points are re-entrant (one Python thread per GPU
 *     as nn.DataParallel does is fine);
 *   - return value: 0 on success, <0 on error (SS_ERR_*), with a thread-local message from ss_last_error();
 *     no exceptions cross the ABI.  Unsupported shapes return SS_ERR_UNSUPPORTED: there is no CPU / cuDNN fallback;
 *   - `flags` bit 0 (SS_SIGNED): disparity range -maxdisp..maxdisp-1, depth 2*maxdisp  (models/submodule.py)
 *                     else       0..maxdisp-1, depth maxdisp                       (models/submodule_.py).
 */
#ifndef SEMSTEREO_B200_H
#define SEMSTEREO_B200_H

#ifdef __cplusplus
#define SS_API extern "C" __attribute__((visibility("default")))
#else
#define SS_API __attribute__((visibility("default")))
#endif

#define SS_SIGNED 1
#define SS_NORM 2

#define SS_ERR_BAD_ARG (-1)
#define SS_ERR_UNSUPPORTED (-2)
#define SS_ERR_CUDA (-3)

/* ---- library ---------------------------------------------------------------------------------------- */
SS_API int ss_version(void);                /* major*10000 + minor*100 + patch */
SS_API const char* ss_last_error(void);     /* thread-local, valid until the next failing call of this thread */
SS_API int ss_sm_count(void);               /* SM count of the current device */

/* ---- K1/K2: cost-volume builders -------------------------------------------------------------------- */
/* build_gwc_volume (submodule.py:198-211, submodule_.py:188-198), build_gwc_volume_norm (submodule.py:224-238,
 * submodule_.py:211-221; flags |= SS_NORM) and build_norm_correlation_volume (submodule.py:244-255; num_groups = 1, SS_NORM).
 * left/right (B,C,H,W) -> volume (B,num_groups,D,H,W), D = 2*maxdisp (signed) or maxdisp. */
SS_API int ss_gwc_volume(const float* left, const float* right, float* volume, int B, int C, int H, int W, int maxdisp,
                         int num_groups, int flags, void* stream);
/* build_concat_volume (submodule.py:173-187; unsigned submodule_.py:166-178 leaves the left half unmasked).
 * -> volume (B,2C,D,H,W). */
SS_API int ss_concat_volume(const float* left, const float* right, float* volume, int B, int C, int H, int W, int maxdisp,
                            int flags, void* stream);

/* ---- K3: patch depthwise conv + channel gate -------------------------------------------------------- */
/* `patch` = Conv3d(G,G,(1,3,3),groups=G,bias=False) (SemStereo.py:219,274), weight (G,1,1,3,3) as (G,9), fused with the
 * channelAtt product sigmoid(gate_logits (B,G,H,W))[:, :, None] * cv (SemStereo.py:98-103).  Either may be NULL. */
SS_API int ss_patch_gate(const float* volume, const float* patch_w_or_null, const float* gate_logits_or_null, float* out, int B,
                         int G, int D, int H, int W, void* stream);
/* 1x1 Conv2d of channelAtt.im_att (SemStereo.py:93-95): out = act(scale*W.x + shift); weight (Cout,Cin); P = H*W. */
SS_API int ss_pointwise_conv2d(const float* in, const float* weight, const float* scale_or_null, const float* shift_or_null,
                               float* out, int B, int Cin, int Cout, int P, int relu, void* stream);

/* ---- K4: 3-D convolutions, fp32-accurate mode ------------------------------------------------------- */
/* convbn_3d (submodule_other.py:845-848), BasicConv(is_3d) (submodule.py:89-116), nn.ConvTranspose3d(k3,s2,p1,op1)
 * (SemStereo.py:124-130), with eval BatchNorm folded into scale/shift and the hourglass residual/ReLU (SemStereo.py:141-142)
 * and the channelAtt gate (SemStereo.py:320) as epilogue.  weight_packed is [K^3][Cin][Cout] (tap = (kd*K+kh)*K+kw; for
 * mode 1 the tap indexes the ConvTranspose3d weight (Cin,Cout,kd,kh,kw) directly).  mode 0: Conv3d(K,stride,pad=K/2);
 * mode 1: ConvTranspose3d(3,2,1,1), out dims = 2x in dims.  residual: same shape as out.  gate: (B,Cout,Ho,Wo) logits. */
SS_API int ss_conv3d_f32(const float* in, const float* weight_packed, const float* scale_or_null, const float* shift_or_null,
                         const float* residual_or_null, const float* gate_logits_or_null, float* out, int B, int Cin, int Cout,
                         int Di, int Hi, int Wi, int K, int stride, int mode, int relu, void* stream);
/* nn.Conv3d(Cin,1,3,padding=1,bias=False) classifier head (SemStereo.py:230,234); weight in PyTorch layout (1,Cin,3,3,3). */
SS_API int ss_conv3d_cout1_f32(const float* in, const float* weight, float* out, int B, int Cin, int D, int H, int W, void* stream);

/* ---- K4: 3-D convolutions, tensor-core mode (tcgen05 / TMEM / TMA, bf16 operands, fp32 accumulation) ------ */
/* Activations in the "blocked channels" layout: bf16 [B][C/8][D][H][W][8], or its phase-split ("s2d") form
 * [B][8 = (d&1,h&1,w&1)][C/8][D/2][H/2][W/2][8] that the stride-2 layer reads and the transposed layer's residual uses.
 * Converters: fp32 NCDHW -> blocked (s2d = 1: phase-split), blocked -> fp32 NCDHW, blocked -> phase-split. */
SS_API int ss_to_blocked_bf16(const float* in_ncdhw, void* out_blocked, int B, int C, int D, int H, int W, int s2d, void* stream);
/* bf16 -> fp32 widening of n elements (n % 8 == 0): host boundary helper, see pipeline.HostPipeline(host_dtype=bfloat16). */
SS_API int ss_widen_bf16(const void* in_bf16, float* out_f32, long long n, void* stream);
SS_API int ss_from_blocked_bf16(const void* in_blocked, float* out_ncdhw, int B, int C, int D, int H, int W, void* stream);
SS_API int ss_blocked_to_s2d(const void* in_blocked, void* out_s2d, int B, int C, int D, int H, int W, void* stream);
/* The 3-D layers of the hourglass stack (same layers as ss_conv3d_f32) + folded eval-BN + residual + ReLU + channelAtt gate.
 * kind 4: Conv2d k3 s1 p1 on a depth-1 volume (D = 1, 9 taps) -- the two convs of `concat_feature` (SemStereo.py:221-223, 314-315).
 * kind 5: Conv3d k3 s1 p1 for narrow layers (Cout == 32 or 64, Cin 32 / 64) with the three depth taps folded into the GEMM N:
 *         weight_packed is [9 in-plane taps][Cin/8][3*Cout (j*Cout + co, kd = 2 - j)][8]; same results as kind 0.
 * kind 0: Conv3d k3 s1 p1;  1: Conv3d k1;  2: Conv3d k3 s2 p1 (input tensor phase-split);  3: ConvTranspose3d k3 s2 p1 op1
 * (residual_s2d: the skip tensor in phase-split layout at OUTPUT resolution, added before the ReLU; SemStereo.py:141-142).
 * D,H,W: input dims of the layer (kind 2: of the un-split input).  weight_packed: bf16 [ceil(Cout/N)][taps][Cin/8][N][8],
 * N = ss_conv3d_tc_ntile(kind,Cin,Cout) (0 = unsupported), Cout zero-padded to a multiple of N, tap = (kd*3+kh)*3+kw (kind 3:
 * of the ConvTranspose3d weight (Cin,Cout,kd,kh,kw)).  out: bf16 blocked (Cout % 8 == 0) or fp32 NCDHW with Cout channels. */
SS_API int ss_conv3d_tc_ntile(int kind, int Cin, int Cout);
/* gate_blocked: sigmoid(channelAtt logits) as fp32 (B,Cout/8,Ho,Wo,8) from ss_gate_sigmoid_blocked.
 * out_mode 0: bf16 blocked, 1: fp32 NCDHW, 2: bf16 phase-split blocked (kinds 0-2, even output dims).
 * skip_weight (kind 3 only): when given, residual_s2d is the INPUT of the hourglass' 1x1 redir conv (phase-split, Cout
 * channels) and skip_weight its weight as bf16 [Cout/8][Cout][8] with the redir BatchNorm scale folded in; the redir conv then
 * runs as one more GEMM tap inside the layer (the caller folds the layer's own BN scale into weight_packed, passes no scale
 * and the sum of the two BN shifts as shift). */
SS_API int ss_conv3d_tc(int kind, const void* in_blocked, const void* weight_packed, const float* scale_or_null,
                        const float* shift_or_null, const float* gate_blocked_or_null, const void* residual_s2d_or_null,
                        const void* skip_weight_or_null, void* out, int out_mode, int B, int Cin, int Cout, int D, int H, int W,
                        int relu, void* stream);
/* bf16x3 split route (fp32-accurate products on the bf16 tensor cores; used for the attention branch, SemStereo.py:273-278, whose
 * logits alone decide the top-k sample selection).  A value x is carried as the pair hi = bf16(x), lo = bf16(x - hi), the two
 * halves STACKED ON THE BATCH AXIS: a split tensor is an ordinary blocked / phase-split tensor of batch 2B, [hi batches | lo batches].
 * x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo is then two launches of the same layer: (1) the 2B-batch tensor with w_hi, raw fp32
 * output (no affine) -> partial sums P (2B,Cout,...); (2) the hi half with w_lo and the two hooks below:
 *   acc_in / acc_in2: fp32 NCDHW tensors shaped like an out_mode-1 output, added to the accumulator before scale/shift/ReLU/gate;
 *   out_split = 1  : the bf16 output (out_mode 0 / 2) is itself written as a split tensor of batch 2B (lo half B batches further).
 * Linear in the operands, so the fused redir skip conv of kind 3 splits the same way (residual = the split skip tensor). */
SS_API int ss_conv3d_tc_ex(int kind, const void* in_blocked, const void* weight_packed, const float* scale_or_null,
                           const float* shift_or_null, const float* gate_blocked_or_null, const void* residual_s2d_or_null,
                           const void* skip_weight_or_null, const float* acc_in_or_null, const float* acc_in2_or_null, void* out,
                           int out_mode, int out_split, int in_split, int B, int Cin, int Cout, int D, int H, int W, int relu,
                           void* stream);
/* in_split = 1: the whole split product in ONE launch (layers for which ss_conv3d_tc_split_supported() != 0: shared memory holds the
 * hi and lo halves of the slice ring and of the weights): in_blocked (and residual_s2d) are split tensors of batch 2B,
 * weight_packed = per tap [hi: Cin/8 chunks][lo: Cin/8 chunks] (the two single packings concatenated along the chunk axis; the
 * skip weight likewise [hi: Cout/8][lo: Cout/8] chunks), and every K step issues the three MMAs into one TMEM accumulator. */
SS_API int ss_conv3d_tc_split_supported(int kind, int Cin, int Cout);
/* ss_to_blocked_bf16 / ss_patch_gate_blocked / ss_conv3d_tc_head with the split hooks: split = 1 writes a split tensor (batch 2B);
 * the head adds fp32 partial sums (B,1,D,H,W) to its result.  ss_to_blocked_bf16_ex split = 2: the channel-stacked K-concat form
 * [hi | lo | hi], bf16 (B, 3C/8, D, H, W, 8): a 1x1 conv with the weight [w_hi | w_hi | w_lo] (K = 3C) over it IS the bf16x3
 * product in one GEMM -- used with ss_conv2d_tc mode 1 for the channelAtt gate convs (SemStereo.py:93-95) and the qkv / final
 * 1x1x1 projections of attention_block (submodule_other.py:795-803; a 3-D volume is a 2-D image of D*H rows for a 1x1 conv). */
SS_API int ss_to_blocked_bf16_ex(const float* in_ncdhw, void* out_blocked, int B, int C, int D, int H, int W, int s2d, int split,
                                 void* stream);
SS_API int ss_patch_gate_blocked_ex(const float* volume, const float* patch_w, const float* gate_logits, void* out_s2d, int B, int G,
                                    int D, int H, int W, int split, void* stream);
SS_API int ss_conv3d_tc_head_ex(const void* in_blocked, const void* weight_packed, const float* acc_in_or_null,
                                const float* acc_in2_or_null, float* out, int in_split, int B, int Cin, int D, int H, int W,
                                void* stream);     /* in_split: split input (batch 2B), weight_packed [2][4][48][8] = hi | lo */
/* concat_volume_generator * att_topk -> concat_stem -> * sigmoid(gate) (SemStereo.py:241-244, 316-320) in ONE kernel: the sparse
 * concat volume is produced tile by tile in shared memory as the GEMM's A operand and never written to HBM.
 * cf_l / cf_r: bf16 blocked (B,4,H,W,8) = concat_feature(f4_*) (32 channels); disp_topk, att_topk: fp32 (B,K,H,W);
 * samples must be the integer disparity bins dmin .. dmin+31 (disparity_sample_topk, :305; dmin = -(maxdisp/4) signed, 0 unsigned);
 * weight_packed: concat_stem's Conv3d(64,32,3) weight in the kind-5 packing of ss_conv3d_tc; scale/shift: folded BN;
 * gate_blocked: fp32 (B,4,H,W,8) pre-activated gate or NULL; out: (B,32,K,H,W) in out_mode 0 / 1 / 2 as ss_conv3d_tc. */
SS_API int ss_concat_stem_fused(const void* cf_l_blocked, const void* cf_r_blocked, const float* disp_topk, const float* att_topk,
                                const void* weight_packed, const float* scale_or_null, const float* shift_or_null,
                                const float* gate_blocked_or_null, void* out, int out_mode, int B, int K, int H, int W, int dmin,
                                int relu, void* stream);
/* nn.Conv3d(32, 1, 3, padding=1, bias=False) classifier heads (SemStereo.py:230,234) with the taps as the GEMM's N dimension:
 * in_blocked bf16 (B,4,D,H,W,8); weight_packed bf16 [4][48][8]: row j*16 + t9 of chunk c = weight[0][c*8+c8][kd = 2-j][t9]
 * (t9 = kh*3+kw; rows with t9 >= 9 zero) -- the three depth taps are folded into N like kind 5 of ss_conv3d_tc;
 * out fp32 (B,1,D,H,W). */
SS_API int ss_conv3d_tc_head(const void* in_blocked, const void* weight_packed, float* out, int B, int Cin, int D, int H, int W,
                             void* stream);
/* 2-D decoder convolutions around the path (FeatUp / Conv2x, models/submodule.py:119-161; segmenthead,
 * :31-52; chal_*, spx*, models/SemStereo.py:207-216) on tcgen05 tensor cores.  Activations bf16 blocked (B,C/8,H,W,8).
 * mode 0: Conv2d 3x3 s1 p1; mode 1: Conv2d 1x1; mode 2: ConvTranspose2d k4 s2 p1 (H,W = input dims, output 2H x 2W).
 * The input is the channel concat [in0 (C0) | in1 (C1)] (torch.cat((x, rem), 1), submodule.py:155) without materialising it;
 * in1 may be NULL with C1 = 0; C0 and C1 must be multiples of 64.  y = conv * scale[co] + shift[co] (folded BN / bias) -> ReLU.
 * N = ss_conv2d_tc_ntile(mode, C0+C1, Cout) (0 = unsupported).  weight_packed (bf16), cb = 64-channel block of the concat:
 *   modes 0/1: [ceil(Cout/N)][ncb][taps][8][N][8] = w[nt*N+n][cb*64+chunk*8+c][tap]            (Cout zero-padded)
 *   mode 2   : [ceil(Cout/N)][ncb][9 slabs], slab s = [8][rows_s][8] with rows = (phase, n) over the output phases the input
 *              shift feeds: shifts (0,0),(-1,0),(+1,0),(0,-1),(0,+1),(-1,-1),(-1,+1),(+1,-1),(+1,+1) feed phases
 *              {0,1,2,3},{0,1},{2,3},{0,2},{1,3},{0},{1},{2},{3} (phase = 2*(oh&1) + (ow&1)); tap k(phase bit, shift):
 *              k(0,0)=1, k(0,-1)=3, k(1,0)=2, k(1,+1)=0; value w_convT[cb*64+chunk*8+c][nt*N+n][k_h][k_w].
 * out_mode 0: bf16 blocked, 1: fp32 NCHW. */
SS_API int ss_conv2d_tc_ntile(int mode, int Cin, int Cout);
SS_API int ss_conv2d_tc(int mode, const void* in0_blocked, int C0, const void* in1_blocked_or_null, int C1, const void* weight_packed,
                        const float* scale_or_null, const float* shift_or_null, void* out, int out_mode, int B, int Cout, int H, int W,
                        int relu, void* stream);
/* ss_conv2d_tc with the epilogue the MobileViTv2 backbone needs (mode 1 only): act 0 none / 1 ReLU / 2 SiLU, and a bf16 blocked
 * residual (B,Cout/8,H,W,8) added after the activation (inverted-residual and transformer skip connections). */
SS_API int ss_conv2d_tc_ex(int mode, const void* in0_blocked, int C0, const void* in1_blocked_or_null, int C1, const void* weight_packed,
                           const float* scale_or_null, const float* shift_or_null, const void* residual_blocked_or_null, void* out,
                           int out_mode, int B, int Cout, int H, int W, int act, void* stream);
/* ---- backbone `Feature` (SemStereo.py:33-56 = timm mobilevitv2_100; architecture restated in semstereo_b200/backbone.py) --------
 * Everything that is not a 1x1 convolution, on the bf16 blocked layout (B,C/8,H,W,8):
 * stem: Conv2d(3,32,3,s2,p1) + folded BN + SiLU from the fp32 NCHW image (weight fp32 (32,3,3,3)); the output has Cout_padded
 *   channels (>= 32, the rest zero) so that the next 1x1 conv sees a multiple of 64 input channels.
 * dwconv: depthwise Conv2d 3x3 p1, stride 1 or 2 (weight fp32 (C,9)) + folded BN + act (0 none, 1 ReLU, 2 SiLU).
 * groupnorm1: nn.GroupNorm(1, C) (mean / variance over C*H*W per sample, fp32 statistics, deterministic two-pass);
 *   workspace: ss_groupnorm1_workspace_floats(B) floats.
 * linear_attention: MobileViTv2 separable self-attention core between qkv_proj and out_proj.  qkv blocked with 2d/8 + 1 chunks:
 *   [0,d/8) key, [d/8,2d/8) value, chunk 2d/8 lane 0 = query.  The softmax runs over the 2x2-patch index for each of the 4 patch
 *   positions = over the pixels of one (y&1, x&1) parity class, so no unfold/fold is materialised.  out = relu(value) * context,
 *   (B,d/8,H,W,8).  workspace: ss_linear_attention_workspace_floats(B, d) floats. */
SS_API int ss_stem_conv3x3_s2(const float* image, const float* weight, const float* scale, const float* shift, void* out_blocked, int B,
                              int H, int W, int Cout_padded, void* stream);
SS_API int ss_dwconv3x3_blocked(const void* in_blocked, const float* weight, const float* scale, const float* shift, void* out_blocked,
                                int B, int C, int H, int W, int stride, int act, void* stream);
SS_API int ss_groupnorm1_workspace_floats(int B);
SS_API int ss_groupnorm1_blocked(const void* in_blocked, const float* gamma, const float* beta, void* out_blocked, float* workspace, int B,
                                 int C, int H, int W, float eps, void* stream);
SS_API int ss_linear_attention_workspace_floats(int B, int d);
SS_API int ss_linear_attention_blocked(const void* qkv_blocked, void* out_blocked, float* workspace, int B, int d, int H, int W, void* stream);
/* F.interpolate(scale 2, bilinear, align_corners=False) of `planes` fp32 (h,w) planes (segmenthead, submodule.py:46-51). */
SS_API int ss_bilinear_up2(const float* in, float* out, int planes, int h, int w, void* stream);
/* segmenthead.conv2 (submodule.py:36,44): Conv2d 1x1 + bias, Cout <= 8, from blocked bf16 (B,C/8,H,W,8) to fp32 (B,Cout,H,W);
 * weight fp32 [Cout][C]. */
SS_API int ss_pointwise_blocked_small(const void* in_blocked, const float* weight, const float* bias_or_null, float* out, int B, int C,
                                      int Cout, int H, int W, void* stream);
/* Producers of the blocked layouts (bf16 mode never materialises the fp32 volumes):
 * sigmoid(gate logits (B,C,H,W)) -> fp32 (B,C/8,H,W,8);  `patch` conv * gate (SemStereo.py:274-276) -> phase-split bf16;
 * concat_volume_generator * att_topk (SemStereo.py:241-244,318) -> blocked bf16 (B,2C/8,K,H,W,8). */
SS_API int ss_gate_sigmoid_blocked(const float* gate_logits, float* out_blocked, int B, int C, int H, int W, void* stream);
SS_API int ss_patch_gate_blocked(const float* volume, const float* patch_w, const float* gate_logits, void* out_s2d, int B, int G, int D,
                                 int H, int W, void* stream);
SS_API int ss_sparse_concat_volume_blocked(const float* cf_l, const float* cf_r, const float* disp_topk, const float* att_topk_or_null,
                                           void* volume_blocked, int B, int C, int K, int H, int W, void* stream);

/* ---- K5: windowed 3-D attention --------------------------------------------------------------------- */
/* attention_block.forward (submodule_other.py:805-837) for window-divisible D,H,W.  wqkv_t = qkv_3d.weight^T [C][3C],
 * wo_t = final1x1.weight^T [C][C] (wo_t[c][co]).  Only C=128, 16 heads, windows of 64 or 96 tokens. */
SS_API int ss_window_attention3d(const float* x, const float* wqkv_t, const float* bqkv, const float* wo_t, const float* bo,
                                 float* out, int B, int C, int D, int H, int W, int bd, int bh, int bw, int num_heads,
                                 void* stream);

/* bf16x3 split route: the fp32 softmax(q k^T * scale) v core alone.  qkv fp32 (B,3C,D,H,W) (channel = which*C + head*8 + j, what
 * qkv_3d produces, submodule_other.py:813-816) -> out_tri bf16 (B, 3*C/8, D, H, W, 8) in the [hi | lo | hi] K-concat form of
 * ss_to_blocked_bf16_ex(split = 2), ready for the final 1x1x1 conv as an fp32-accurate GEMM. */
SS_API int ss_window_attention_core_f32(const float* qkv, void* out_tri, int B, int C, int D, int H, int W, int bd, int bh, int bw,
                                        int num_heads, void* stream);

/* Tensor-core mode: the qkv Linear and the final 1x1x1 conv run as ss_conv3d_tc kind 1 layers (Cin=128 -> 384 / 128, bias as
 * shift); this is the softmax(q k^T * scale) v core in between on the blocked layout, where head h is channel chunk h:
 * qkv (B,48,D,H,W,8) bf16 -> out (B,16,D,H,W,8) bf16. */
SS_API int ss_window_attention_core_blocked(const void* qkv_blocked, void* out_blocked, int B, int C, int D, int H, int W, int bd,
                                            int bh, int bw, int num_heads, void* stream);

/* ---- K6/K7/K8: attention statistics, sample strength, top-k selection -------------------------------- */
/* F.interpolate(trilinear, x2) -> softmax(dim=1) -> disparity_regression -> disparity_variance -> sigmoid(beta+gamma*var)
 * (SemStereo.py:279-287).  cost_att (B,1,D8,H8,W8) -> att_up (B,2*D8,2*H8,2*W8) logits, mu (B,2H8,2W8), gate (B,1,2H8,2W8).
 * beta/gamma: device pointers to the 1-element parameters.  dmin = disparity value of bin 0. */
SS_API int ss_att_stats(const float* cost_att, const float* beta, const float* gamma, float* att_up, float* mu, float* gate, int B,
                        int D8, int H8, int W8, float dmin, void* stream);
/* Propagation x2 + SpatialTransformer_grid + (L*Rwarp).mean(C) + softmax(strength*var) (SemStereo.py:288-293)
 * -> strength (B,5,H,W). */
SS_API int ss_sample_strength(const float* feat_l, const float* feat_r, const float* mu, const float* gate, float* strength, int B,
                              int C, int H, int W, void* stream);
/* Propagation_prob, mix, softmax, top-K (ascending bin order; ties toward the lower bin), gathers, renormalised expectation
 * (SemStereo.py:295-310).  ind_k (B,1,K,H,W) int64 (may be NULL), att_topk/disp_topk (B,K,H,W), pred_att (B,H,W),
 * prob_or_null (B,nbins,H,W) = the full softmax.  disp_topk = bin - disp_offset. */
SS_API int ss_topk_select(const float* att_up, const float* strength, long long* ind_k_or_null, float* att_topk, float* disp_topk,
                          float* pred_att, float* prob_or_null, int B, int nbins, int K, int H, int W, float disp_offset,
                          void* stream);

/* ---- K9/K10/K11/K12 ----------------------------------------------------------------------------------- */
/* concat_volume_generator (SemStereo.py:241-244) * att_topk (:318) -> volume (B,2C,K,H,W). att_topk may be NULL (= 1). */
SS_API int ss_sparse_concat_volume(const float* cf_l, const float* cf_r, const float* disp_topk, const float* att_topk_or_null,
                                   float* volume, int B, int C, int K, int H, int W, void* stream);
/* regression_topk (submodule.py:434-442): cost, disp_samples (B,D,H,W) -> pred (B,1,H,W); ties toward the lower index. */
SS_API int ss_regression_topk(const float* cost, const float* disp_samples, float* pred, int B, int D, int K, int H, int W,
                              void* stream);
/* SSR_upsample.forward (submodule.py:421-431), eval BatchNorm.  packed_host: HOST array of ss_ssr_param_count(nc) floats:
 * a0,b0 | wc[nc][9], bc[nc] | s2[nc], t2[nc] | w1[nc][nc], b1[nc], s1[nc], t1[nc] | w2[nc][nc], b2[nc], s2b[nc], t2b[nc] |
 * w3[nc], b3   (BN layers as scale/shift).  depth_low (B,1,h,w); spx, label (B,nc,4h,4w) -> out (B,4h,4w). */
SS_API int ss_ssr_param_count(int num_classes);
SS_API int ss_ssr_upsample(const float* depth_low, const float* spx, const float* label, float* out, const float* packed_host, int B,
                           int h, int w, int num_classes, void* stream);
/* The model's two SSR_upsample calls (SemStereo.py:312 pred_att, :324 pred; same spx / label) in one pass: the class gate is
 * computed once per pixel.  out_a / out_b are bit-identical to two ss_ssr_upsample calls. */
SS_API int ss_ssr_upsample2(const float* depth_low_a, const float* depth_low_b, const float* spx, const float* label, float* out_a,
                            float* out_b, const float* packed_host, int B, int h, int w, int num_classes, void* stream);
/* context_upsample (submodule_.py:311-323): depth_low (B,1,h,w), up_weights (B,9,4h,4w) -> out (B,4h,4w). */
SS_API int ss_context_upsample(const float* depth_low, const float* up_weights, float* out, int B, int h, int w, void* stream);
/* disparity_regression (submodule.py:164-170) / disparity_variance (:257-263): prob (B,D,H,W); dmin = value of bin 0. */
SS_API int ss_disparity_regression(const float* prob, float* out, int B, int D, int H, int W, float dmin, void* stream);
SS_API int ss_disparity_variance(const float* prob, const float* disparity, float* out, int B, int D, int H, int W, float dmin,
                                 void* stream);
/* Propagation (submodule.py:290-307; D=1) and Propagation_prob (:361-377): in (B,1,D,H,W) -> out (B,5,D,H,W). */
SS_API int ss_propagation(const float* in, float* out, int B, int D, int H, int W, void* stream);
/* SpatialTransformer_grid (submodule.py:265-288): x,y (B,C,H,W), disp_samples (B,K,H,W) -> y_warped, x_rep (B,C,K,H,W). */
SS_API int ss_spatial_transformer_grid(const float* x, const float* y, const float* disp_samples, float* y_warped,
                                       float* x_rep_or_null, int B, int C, int K, int H, int W, void* stream);

/* ---- backward (vector-Jacobian products) of the volume / regression operators: BASELINE config #5, first step ------------
 * Gather-form, deterministic, fp32.  Shapes as in the forward entry points; every grad_* output is fully overwritten. */
/* build_gwc_volume(_norm): grad_volume (B,G,D,H,W) -> grad_left, grad_right (B,C,H,W).  flags as ss_gwc_volume. */
SS_API int ss_gwc_volume_backward(const float* left, const float* right, const float* grad_volume, float* grad_left, float* grad_right,
                                  int B, int C, int H, int W, int maxdisp, int num_groups, int flags, void* stream);
/* build_concat_volume: grad_volume (B,2C,D,H,W) -> grad_left, grad_right (B,C,H,W). */
SS_API int ss_concat_volume_backward(const float* grad_volume, float* grad_left, float* grad_right, int B, int C, int H, int W,
                                     int maxdisp, int flags, void* stream);
/* disparity_regression: grad_out (B,H,W) -> grad_prob (B,D,H,W). */
SS_API int ss_disparity_regression_backward(const float* grad_out, float* grad_prob, int B, int D, int H, int W, float dmin, void* stream);
/* regression_topk: grad_pred (B,1,H,W) -> grad_cost, grad_samples (B,D,H,W) (zero outside the selected K; the sort carries no gradient). */
SS_API int ss_regression_topk_backward(const float* cost, const float* disp_samples, const float* grad_pred, float* grad_cost,
                                       float* grad_samples, int B, int D, int K, int H, int W, void* stream);
/* context_upsample: grad_out (B,4h,4w) -> grad_depth (B,1,h,w), grad_weights (B,9,4h,4w). */
SS_API int ss_context_upsample_backward(const float* depth_low, const float* up_weights, const float* grad_out, float* grad_depth,
                                        float* grad_weights, int B, int h, int w, void* stream);

/* Propagation / Propagation_prob: grad_out (B,5,[D,]H,W) -> grad_in (B,1,[D,]H,W) (D = 1 for the 2-D module). */
SS_API int ss_propagation_backward(const float* grad_out, float* grad_in, int B, int D, int H, int W, void* stream);
/* disparity_variance: grad_out (B,1,H,W) -> grad_prob (B,D,H,W), grad_disparity (B,1,H,W). */
SS_API int ss_disparity_variance_backward(const float* prob, const float* disparity, const float* grad_out, float* grad_prob,
                                          float* grad_disparity, int B, int D, int H, int W, float dmin, void* stream);
/* SpatialTransformer_grid: grad_y_warped, grad_x_rep (B,C,K,H,W) -> grad_x, grad_y (B,C,H,W), grad_disp (B,K,H,W).
 * grad_y is accumulated with atomics and must be ZERO on entry; grad_x_rep / grad_x may both be NULL. */
SS_API int ss_spatial_transformer_grid_backward(const float* y, const float* disp_samples, const float* grad_y_warped,
                                                const float* grad_x_rep_or_null, float* grad_x_or_null, float* grad_y, float* grad_disp,
                                                int B, int C, int K, int H, int W, void* stream);

/* ---- training closure of the 3-D conv stack (BASELINE config #5; convbn_3d / BasicConv(is_3d) in training mode) -----------------
 * Conv3d weight gradient: x (B,Cin,Di,Hi,Wi), grad_out (B,Cout,Do,Ho,Wo) -> grad_weight_packed [K^3][Cin][Cout] (the packing of
 * ss_conv3d_f32; MUST be zero on entry, accumulated with atomics).  K in {1,3}, pad K/2, stride in {1,2}.  The INPUT gradient
 * needs no entry point of its own: it is ss_conv3d_f32 on grad_out with a re-packed weight (flipped k3 s1 / transposed-conv k3 s2 /
 * W^T k1; semstereo_b200/train_ops.py). */
SS_API int ss_conv3d_wgrad_f32(const float* x, const float* grad_out, float* grad_weight_packed, int B, int Cin, int Cout, int Di, int Hi,
                               int Wi, int K, int stride, void* stream);

/* Small-channel 2-D convolutions in training mode (Cin, Cout <= 8, k in {1, 3}, stride 1, zero padding k/2): the convs of
 * SSR_upsample (models/submodule.py:394-431: Conv2d(1,6,3,1,1), Conv2d(6,6,1), Conv2d(6,1,1)) at full image resolution.
 * in (B,Cin,H,W), weight (Cout,Cin,k,k), out (B,Cout,H,W), all fp32; the input gradient is the same call on dY with the flipped,
 * channel-transposed weight.  ss_conv2d_small_wgrad_f32 ACCUMULATES into grad_weight (Cout,Cin,k,k) (zero it first); it exists for
 * the three SSR shapes (ss_conv2d_small_wgrad_supported). */
SS_API int ss_conv2d_small_f32(const float* in, const float* weight, float* out, int B, int Cin, int Cout, int H, int W, int k, void* stream);
SS_API int ss_conv2d_small_wgrad_supported(int Cin, int Cout, int k);
SS_API int ss_conv2d_small_wgrad_f32(const float* x, const float* grad_out, float* grad_weight, int B, int Cin, int Cout, int H, int W, int k,
                                     void* stream);
/* BatchNorm{2,3}d with BATCH statistics over (B, S = spatial size) per channel (training mode of nn.BatchNorm, appendix C of
 * SURVEY.md): batch_mean / batch_var (biased) are outputs of the forward (the caller updates running_mean / running_var with
 * momentum and the unbiased variance) and inputs of the backward.  out = (x - mean) * rsqrt(var + eps) * weight + bias (-> ReLU).
 * backward: grad_weight = sum(dy * xhat), grad_bias = sum(dy), grad_x = w * rstd * (dy - (grad_bias + xhat * grad_weight) / N);
 * with relu = 1 in the forward the caller masks grad_out with (out > 0) first. */
SS_API int ss_bn_workspace_bytes(int C);          /* device scratch for the per-channel reductions (64 partial sums per channel) */
SS_API int ss_bn_train_forward(const float* x, const float* weight_or_null, const float* bias_or_null, float* out, float* batch_mean,
                               float* batch_var, void* workspace, int B, int C, long long S, float eps, int relu, void* stream);
SS_API int ss_bn_train_backward(const float* x, const float* grad_out, const float* batch_mean, const float* batch_var,
                                const float* weight_or_null, float* grad_x, float* grad_weight, float* grad_bias, void* workspace, int B, int C,
                                long long S, float eps, void* stream);

/* attention_block in training mode (submodule_other.py:805-837): the qkv Linear and the final 1x1x1 conv are k = 1 layers of the
 * differentiable conv above; in between, the fp32 softmax core with an fp32 NCDHW output (qkv (B,3C,D,H,W) -> (B,C,D,H,W)) and its
 * backward (grad_out (B,C,D,H,W) -> grad_qkv (B,3C,D,H,W)).  C = 128, 16 heads, windows <= 96 tokens (forward: 64 / 96, bw = 4). */
/* The masked softmax core of attention_block's padded branch (models/submodule_other.py:809-829) for volumes whose H AND W were both
 * zero-padded to the window: qkv (B,3C,D,H,W) fp32 is the qkv Linear of the PADDED volume, tokens with h >= H0 or w >= W0 are padding,
 * a score between a padded and a real token gets -1000 before the softmax.  Output fp32 (B,C,D,H,W), channel = head*hd + j; the caller
 * crops to (H0, W0) and applies the final 1x1x1 conv.  (With one padded axis the reference masks nothing: the unmasked entry points on
 * the padded volume are exact.) */
SS_API int ss_window_attention_core_f32_masked(const float* qkv, float* out_f32, int B, int C, int D, int H, int W, int bd, int bh, int bw,
                                               int num_heads, int H0, int W0, void* stream);
SS_API int ss_window_attention_core_f32_out(const float* qkv, float* out_f32, int B, int C, int D, int H, int W, int bd, int bh, int bw,
                                            int num_heads, void* stream);
SS_API int ss_window_attention_core_backward(const float* qkv, const float* grad_out, float* grad_qkv, int B, int C, int D, int H, int W,
                                             int bd, int bh, int bw, int num_heads, void* stream);
/* The same backward for the masked core (ss_window_attention_core_f32_masked): qkv / grad_out / grad_qkv of the PADDED volume, tokens with
 * h >= H0 or w >= W0 are padding.  H0 = H and W0 = W is the unmasked backward. */
SS_API int ss_window_attention_core_backward_masked(const float* qkv, const float* grad_out, float* grad_qkv, int B, int C, int D, int H, int W,
                                                    int bd, int bh, int bw, int num_heads, int H0, int W0, void* stream);
/* F.interpolate(scale_factor 4, bilinear, align_corners=False) of `planes` fp32 (h,w) planes and its VJP (SSR_upsample in training
 * mode, submodule.py:424, composed from differentiable kernels because its BatchNorm layers then need batch statistics). */
SS_API int ss_bilinear_up4(const float* in, float* out, int planes, int h, int w, void* stream);
SS_API int ss_bilinear_up4_backward(const float* grad_out, float* grad_in, int planes, int h, int w, void* stream);

#endif /* SEMSTEREO_B200_H */
