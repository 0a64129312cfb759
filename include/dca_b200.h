/* dca_b200.h -- C ABI of the B200 (sm_100a) DCANet cost-volume hot path.
 *
 * Drop-in boundary for everything between the 1/4-resolution feature maps and the final disparity
 * of the reference model (cocowy1/Cost-Volume-Aggregation-in-Stereo-Matching-Revisited,
 * models/gwcnet_dca_g.py:216-240).  The reference has no native code on this path (pure PyTorch
 * eager); its own precedent for this shape of boundary is models/libs/GANet/src/GANet_cuda.cpp:5-64
 * (`extern "C" int f(...)`).  Every entry point here
 *   - takes raw DEVICE pointers, sizes and a cudaStream_t (passed as void*),
 *   - never allocates, never synchronises, never throws,
 *   - returns DCA_OK (0) or a negative error code.
 * The caller (Python via ctypes) owns all buffers.
 *
 * "cost planes" = channels-last 16-bit tensor [planes][B][D][H][W][C]; plane 0 = h16(x),
 * plane 1 = h16(x - plane0) (only when planes == 2, the parity precision mode); h16 = IEEE fp16 unless the library
 * was built with -DDCA_F16_PLANES=0 (bf16), see dca_plane_format().
 */
#ifndef DCA_B200_H
#define DCA_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define DCA_OK 0
#define DCA_ERR_ARG (-1)         /* bad pointer / shape */
#define DCA_ERR_LAUNCH (-2)      /* CUDA launch failed */
#define DCA_ERR_UNSUPPORTED (-3) /* shape outside what the kernels were built for */

#define DCA_ACT_NONE 0
#define DCA_ACT_RELU 1
#define DCA_ACT_LEAKY 2 /* LeakyReLU(0.1) */

#define DCA_CONV_K3S1 0  /* Conv3d k3 s1 p1          (submodule.py:121-124) */
#define DCA_CONV_K3S2 1  /* Conv3d k3 s2 p1          (cva.py:16)            */
#define DCA_CONV_T3S2 2  /* ConvTranspose3d k3 s2 p1 op1 (cva.py:20-22)     */
#define DCA_CONV_K1 3    /* Conv3d 1x1x1             (cva.py:24,55)         */
#define DCA_CONV_2D3 4   /* Conv2d 3x3 s1 p1, Di==1  (gwcnet_dca_g.py:112-115) */

int dca_version(void);
/* 16-bit element format of the cost planes / tensor-core operand packs of this build: 1 = IEEE fp16 (default; hi + lo
 * = 22 significand bits), 0 = bf16 (compile with -DDCA_F16_PLANES=0; 16 bits, 8 more exponent bits). */
int dca_plane_format(void);

/* (1) volume construction -------------------------------------------------------------------- */
/* Fused replacement of build_gwc_volume (submodule.py:157-167) + build_concat_volume (:134-145)
 * + torch.cat (gwcnet_dca_g.py:220).  gwc_* [B,C,H,W] fp32, cat_* [B,Cc,H,W] fp32 (NCHW, as the
 * extractor emits them); vol = cost planes [planes][B][D][H][W][Cv], channels = G groups, Cc left,
 * Cc right, zero pad up to Cv (multiple of 8, <= 64). */
int dca_volume_gwc_concat(const float* gwc_l, const float* gwc_r, const float* cat_l, const float* cat_r, void* vol,
                          int B, int C, int G, int Cc, int D, int H, int W, int Cv, int planes, void* stream);
/* Reference-API forms: fp32 NCDHW outputs [B,G,D,H,W] / [B,2C,D,H,W]. */
int dca_build_gwc_volume_f32(const float* l, const float* r, float* vol, int B, int C, int G, int D, int H, int W,
                             void* stream);
int dca_build_concat_volume_f32(const float* l, const float* r, float* vol, int B, int C, int D, int H, int W,
                                void* stream);

/* (3) convolution family ----------------------------------------------------------------------- */
/* y = act(scale * conv(x, w) + shift + res_pre) + res_post.   CUDA-core fp32 kernel, any mode.
 * w_packed fp32 [taps][Cin][CoutPad] (dca_pack_weights); scale/shift [CoutPad] or NULL;
 * res_* cost planes shaped like y or NULL; out_kind 0 = cost planes, 1 = fp32 channels-last. */
int dca_conv3d_direct(int mode, const void* x, int planes_in, const float* w_packed, const float* scale,
                      const float* shift, const void* res_pre, const void* res_post, int planes_res, void* y,
                      int planes_out, int out_kind, int act, int B, int Cin, int Cout, int CoutPad, int Di, int Hi,
                      int Wi, int Do, int Ho, int Wo, void* stream);
/* Conv3d k3 s1 p1, Cin=32 -> 1 channel, no BN: fp32 logits [B,D,H,W] (cva.py:53, gwcnet_dca_g.py:168).
 * w_host = HOST pointer, fp32 [27][Cin]: the 864 weights travel as launch parameters (constant bank). */
int dca_conv3d_cout1(const void* x, int planes_in, const float* w_host, float* y, int B, int Cin, int D, int H, int W,
                     void* stream);
/* tcgen05/TMEM/TMA implicit-GEMM member of the family (conv_tc.cu), modes K3S1 / K3S2 / T3S2 / K1, Cin,Cout in {32,64}:
 *   y = act(scale * (conv(x, w) + trilinear_x2(up)) + shift + res_pre) + res_post
 * w_tc = bf16 operand pack from dca_pack_weights_tc; `up` (optional, Cout == 32) is a cost-plane tensor at half the
 * output resolution: the trilinear x2 + cat + 1x1x1 fuse of cva.py:64,69 folded into the epilogue.
 * `side` (optional, T3S2 only) is a cost-plane tensor [.., Do,Ho,Wo, side_c] whose 1x1x1 conv (weights = tap 27 of
 * w_tc, zero padded to Cin) joins the same GEMM: Multi_Aggregation's conv3 + redir (cva.py:20-29) in one kernel. */
int dca_conv3d_tc(int mode, const void* x, int planes_in, const void* w_tc, const float* scale, const float* shift,
                  const void* res_pre, const void* res_post, int planes_res, const void* up, int planes_up,
                  const void* side, int side_c, void* y, int planes_out, int act, int B, int Cin, int Cout, int Di,
                  int Hi, int Wi, int Do, int Ho, int Wo, void* stream);
int dca_pack_weights_tc(const float* w, int transposed, int Co, int Ci, int taps, void* out, int planes, void* stream);
long long dca_pack_weights_tc_bytes(int Co, int Ci, int taps, int planes);
/* Outputs at twice the input resolution, 8 output parity classes per low-res tile fed from shared halo slabs:
 *   kind 0: y = act(scale * (ConvTranspose3d_k3s2p1op1(x) + side 1x1x1) + shift) + res_post     (cva.py:20-29)
 *           x [planes][B][Dl][Hl][Wl][Cin]; w_tc = 27 taps (+ tap 27: side weights zero-padded to Cin)
 *   kind 1: y = act(scale * (trilinear_x2(x) + side 1x1x1) + shift) + res_post                  (cva.py:64,55,69)
 *           x = border-replicated [planes][B][Dl+2][Hl+2][Wl+2][32]; w_tc = 5 taps: {27,9,3,1}/64*I, side weights
 *   kind 2: like kind 1 but x = [planes][B][Dl][Hl+2][Wl+2][32] is already at the output depth Dl (depth axis
 *           interpolated by the producer): bilinear x2 in (h, w) only; w_tc = 4 taps {9,3,1}/16*I, side weights;
 *           side / y are [planes][B][Dl][2Hl][2Wl][..]
 * side [planes][B][2Dl][2Hl][2Wl][side_c]; Cout = 32. */
int dca_up2_tc(int kind, const void* x, int planes, const void* side, int side_c, const void* w_tc, const float* scale,
               const float* shift, const void* res_post, int planes_res, void* y, int act, int B, int Cin, int Dl,
               int Hl, int Wl, void* stream);
/* Conv3d k3 s1 p1, 32 -> 1 channel on the tensor cores: per-tap products P[tap][v] = <x[v,:], w[tap,:]> as a 1x1x1
 * GEMM (fp32, tap-major [ntap][B*D*H*W]) followed by the 27-tap shifted sum (cva.py:53, gwcnet_dca_g.py:168). */
int dca_conv1_taps_tc(const void* x, int planes, const void* w_tc, float* P, int ntap, int B, int D, int H, int W,
                      void* stream);
int dca_tap_gather3d(const float* P, float* out, int B, int D, int H, int W, void* stream);
/* Fused tail, first half: Conv3d k3 s1 p1 (Cin in {32,64} -> 32) + BN + activation whose epilogue immediately applies
 * the Conv3d(32 -> 1, k3) that follows it in classif3 (gwcnet_dca_g.py:166-168) / cva.classify (cva.py:51-53) tap by tap:
 *   P[tap][v] = sum_c w27[tap][c] * act(scale * conv(x, w)[v][c] + shift)   (fp32, tap-major [27][B*D*H*W]).
 * The 32-channel intermediate never reaches HBM.  use_march = 1: depth-marching kernel, w = dca_pack_weights_tc_march
 * pack; 0: halo-slab kernel, w = dca_pack_weights_tc pack.  w27_host is a HOST pointer ([27][32] fp32, launch params). */
int dca_conv3d_tc_taps27(const void* x, int planes, const void* w, int use_march, const float* scale, const float* shift,
                         const float* w27_host, float* P, int act, int B, int Cin, int D, int H, int W, void* stream);
/* cva.classify.2 (cva.py:53) + the per-pixel class statistics of SemanticLevelContext (semantic_level.py:98-116) in one
 * launch: logits [B,D,H,W] = 27-tap shifted sum of P, and cls / e / S as dca_class_stats computes them from those logits
 * (scratch as there).  Bit-identical to dca_tap_gather3d followed by dca_class_stats. */
int dca_tap_gather_class_stats(const float* P, float* logits, int* cls, float* e, float* S, void* scratch, int B, int D,
                               int H, int W, void* stream);
/* Fused tail, second half: logits = 27-tap shifted sum of P, softmax over D, disparity regression
 * (gwcnet_dca_g.py:235-239, submodule.py:127-131): pred [B,H,W]; logits_out optional [B,D,H,W].  Neither the logits nor
 * the probability volume are materialised. */
int dca_tap_gather_softmax_regress(const float* P, float* pred, float* logits_out, int B, int D, int H, int W,
                                   void* stream);
/* Depth-marching member of the family: Conv3d k3 s1 p1, Cout = 32, Cin in {32,64}.  A CTA walks the input planes of a
 * chunk of <= 16 output planes; each halo slab is read once per (kh,kw) for all three kd (N = 96).  Same epilogue
 * contract as dca_conv3d_tc.  w_march from dca_pack_weights_tc_march ([kh*3+kw][plane][kd][32][Cin] bf16). */
int dca_conv3d_tc_march(const void* x, int planes, const void* w_march, const float* scale, const float* shift,
                        const void* res_pre, const void* res_post, int planes_res, void* y, int act, int B, int Cin,
                        int D, int H, int W, void* stream);
int dca_pack_weights_tc_march(const float* w, int Ci, void* out, int planes, void* stream);
long long dca_pack_weights_tc_march_bytes(int Ci, int planes);
/* 2-D 3x3 s1 p1 convs of the propagation net (gwcnet_dca_g.py:112-115) on the halo-slab tcgen05 kernel.
 * x [planes][B][1][H][W][Cin] (Cin 64/128); y cost planes (Cout % 64 == 0) or fp32 channels-last [B][H][W][Cout]. */
int dca_conv2d_tc(const void* x, int planes, const void* w_tc2d, const float* scale, const float* shift, void* y,
                  int out_f32, int act, int B, int Cin, int Cout, int H, int W, void* stream);
int dca_pack_weights_tc2d(const float* w, int Co, int Ci, void* out, int planes, void* stream);
long long dca_pack_weights_tc2d_bytes(int Co, int Ci, int planes);
/* The 2-D family member behind the 1/4-resolution part of the front end: feature_extraction.layer2[1:],
 * layer3, layer4 (dilation 2) and lastconv (gwcnet_dca_g.py:22-29, BasicBlock submodule.py:251-273), Guidance.layer2[1],
 * conv_g0 and guidance (submodule.py:395-460, ResidualBlock :305-347).
 *   y = act_post(act(scale * conv2d_3x3(x, dilation dil) + shift) + res)
 * Cin a multiple of 64 up to 320 (64-channel slabs), dil in {1, 2} (2: H and W even; four parity sub-images through
 * strided tensor maps), res = optional cost planes shaped like y.  1x1 convs are passed as 3x3 weights whose only
 * non-zero tap is the centre. */
int dca_conv2d_tc_ex(const void* x, int planes, const void* w_tc2d, const float* scale, const float* shift,
                     const void* res, int act_post, void* y, int out_f32, int act, int B, int Cin, int Cout, int H, int W,
                     int dil, void* stream);
/* The two 3-channel stems of the front end: Conv2d(3 -> 32, K x K, stride 2, pad K/2) (+ bias) + BN + act, K in {3, 7}
 * (feature_extraction.firstconv[0] gwcnet_dca_g.py:19; Guidance.conv_start submodule.py:413-414).  x fp32 NCHW [B,3,H,W],
 * w fp32 [32][3][K][K] (torch layout), scale/shift = folded BN (+ bias); y cost planes [planes][B][1][Ho][Wo][32]. */
int dca_conv2d_stem(const float* x, const float* w, const float* scale, const float* shift, void* y, int planes, int act,
                    int B, int H, int W, int K, void* stream);
/* The same conv over the channel concatenation cat(x0, x1, x2) (each a multiple of 64 channels, <= 320 in total; unused
 * sources NULL) without materialising it: feature_extraction.lastconv on cat(l2, l3, l4) (gwcnet_dca_g.py:60-65). */
int dca_conv2d_tc_cat(const void* x0, int C0, const void* x1, int C1, const void* x2, int C2, int planes,
                      const void* w_tc2d, const float* scale, const float* shift, void* y, int out_f32, int act, int B,
                      int Cout, int H, int W, void* stream);
/* disparity groups per block of dca_tap_gather_softmax_regress (16 / 8 / 6 (default) / 4; timing experiments). */
int dca_tap_gather_set_groups(int n);
/* planes per work item of the depth-marching kernel: 0 (default) = chosen per shape (halo overhead vs fill of the last
 * wave), n > 0 = forced (timing experiments). */
int dca_tc_set_march_n(int n);
/* k3 s1 main loop selector: 1 = halo'd slab reuse (default), 0 = one TMA box per tap. */
int dca_tc_set_halo(int on);
/* bit 0 (default 1): volumes with C = 320, Cc = 12, G in {8, 20, 40} and W % 8 == 0 use the TMA-staged group-pair kernel
   (DCANet's shape is G = 40); 0: the generic kernel everywhere (A/B timing and tests).  Bits 1-2: timing probes. */
int dca_volume_set_v2(int on);
/* dca_disp_attention at D/8 == 24: 1 (default) two warps per pixel + mma.sync attention core, 2 two warps + fp32 FMA
 * core, 0 one warp per pixel (A/B timing). */
int dca_attention_set_team(int on);
/* timing probes of the tcgen05 kernels: (flags >> 4) & 1 skips the epilogue math + stores, & 2 the MMAs of the halo
 * and up2 kernels; `reserved` must be 1. */
int dca_tc_set_tuning(int reserved, int flags);
/* The tensor core truncates its fp32 accumulator toward zero at every MMA (measured: a systematic -1.56e-8 relative per
 * accumulation step).  The tcgen05 conv epilogues multiply the main accumulator block by 1 + kappa * steps; this sets
 * kappa (default 1.56e-8f, 0 = off). */
int dca_tc_set_trunc_comp(float kappa);
/* Programmatic dependent launch: 1 (default) = every kernel of the forward is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization and waits (griddepcontrol.wait) after its prologue, so prologues
 * and launch latency overlap the previous kernel's tail; 0 = plain stream order.  Same results. */
int dca_set_pdl(int on);
int dca_pdl_enabled(void);
/* dca_up2_tc kind 0 with Cin = 64 and a side input: 1 (default) = two depth-adjacent tiles per weight fetch
 * (conv_tc_deconv_pair_kernel), 0 = one tile per fetch (conv_tc_up2_kernel); same results. */
int dca_tc_set_deconv_pair(int on);
/* dca_up2_tc kind 2: depth of the side-box ring, 2 (default) or 4 (timing experiments; measured: no difference). */
int dca_tc_set_up2_side_slots(int n);

/* (2) DCA module ------------------------------------------------------------------------------- */
/* AvgPool3d(3, 2, 1) (cva.py:39).  C == 32: TMA-staged depth-marching kernel (each input plane read once); other C, or
 * after dca_pool_set_march(0): the thread-per-output kernel dca_avgpool3d_simple.  Same results (sum order differs). */
int dca_avgpool3d(const void* x, void* y, int planes, int B, int C, int Di, int Hi, int Wi, void* stream);
int dca_avgpool3d_simple(const void* x, void* y, int planes, int B, int C, int Di, int Hi, int Wi, void* stream);
int dca_pool_set_march(int on);
/* logits fp32 [B,D,H,W] -> class map int32 [B,H,W], e = exp(P[k_p]) [B,H,W], S [B,D].  S is summed in 64-bit fixed
 * point (order-independent: two runs are bit-identical); scratch = (B*D + 1) x 8 bytes, 8-byte aligned, zeroed here. */
int dca_class_stats(const float* logits, int* cls, float* e, float* S, void* scratch, int B, int D, int H, int W,
                    void* stream);
/* weights: 7 x [32][32] transposed (q0,q1,k0,k1,v,o,Wa) then 6 x (scale[32],shift[32]). */
/* pad = 1: y is [planes][B][D+2][H+2][W+2][C] with a replicated 1-voxel border (input of dca_up2_tc kind 1).
 * pad = 2: y is [planes][B][2D][H+2][W+2][C]: already interpolated x2 along depth (align_corners=False), h/w borders
 *          replicated (input of dca_up2_tc kind 2). */
int dca_disp_attention(const void* x, const int* cls, const float* e, const float* S, const float* weights,
                       int has_wa, void* y, int pad, int planes, int B, int C, int D, int H, int W, void* stream);
/* SelfAttentionBlock.forward(query_feats, key_feats) (models/augment/SelfAttention_bn.py:62-98), the generic two-input
 * form: query, key, y = cost planes [planes][B][D][H][W][32]; weights as above (the 7th matrix is not used). */
int dca_self_attention(const void* query, const void* key, const float* weights, void* y, int planes, int B, int C,
                       int D, int H, int W, void* stream);
/* y = scale * (trilinear_x2(t) + WcT^T cost) + shift ; t at (Dl,Hl,Wl), cost / y at twice that. */
int dca_upsample_fuse(const void* t, const void* cost, const float* WcT, const float* scale, const float* shift,
                      void* y, int planes, int B, int C, int Dl, int Hl, int Wl, void* stream);

/* (4) regression + convex upsampling ------------------------------------------------------------ */
int dca_softmax_regress(const float* logits, float* pred, int B, int D, int H, int W, void* stream);
/* disparity_regression on its own (models/submodule.py:127-131): pred[b,h,w] = sum_d d * x[b,d,h,w], x not renormalised. */
int dca_regress_f32(const float* x, float* pred, int B, int D, int H, int W, void* stream);
int dca_convex_upsample(const float* mask, const float* disp, float* out, int B, int H, int W, void* stream);

/* (5) H-sharded single-pair mode: halo rows over NVLink peer memory ------------------------------ */
/* Replaces, for BASELINE configs[4] (one pair row-sharded over the GPUs of a box), the exchange a
 * multi-GPU port of the reference would do with NCCL send/recv after every 3x3x3 layer.  `t` is viewed as
 * [outer][rows][inner_bytes]; the rank owns rows [h, rows-h) and only the `live` <= h halo rows next to them are ever read.
 * dca_halo_push stores the owned rows [h,h+live) / [rows-h-live,rows-h) into the upper / lower neighbour's staging slot (peer-mapped addresses, 0 = no neighbour) and bumps the neighbour's
 * 64-bit arrival counter by dca_halo_push_ctas(); dca_halo_wait_unpack spins (bounded: dca_halo_set_timeout_ms, default 10 s, then *err = 1) until the
 * own counters reach `target` and copies the staged rows into the halo rows [h-live,h) / [rows-h,rows-h+live).  Needs 16-byte aligned rows
 * (DCA_ERR_UNSUPPORTED otherwise: use the NCCL transport for that tensor). */
int dca_halo_push_ctas(void);
int dca_halo_set_timeout_ms(int ms);
int dca_halo_push(const void* t, long long outer, long long rows, long long inner_bytes, int h, int live, void* peer_up_stage,
                  void* peer_down_stage, void* peer_up_flag, void* peer_down_flag, void* stream);
int dca_halo_wait_unpack(void* t, long long outer, long long rows, long long inner_bytes, int h, int live,
                         const void* stage_top,
                         const void* stage_bottom, const void* flag_top, const void* flag_bottom,
                         unsigned long long target, void* err, void* stream);

/* the two calls above as ONE launch (same arguments and semantics): what hshard.PeerHalo issues per exchange */
int dca_halo_exchange(void* t, long long outer, long long rows, long long inner_bytes, int h, int live,
                      void* peer_up_stage, void* peer_down_stage, void* peer_up_flag, void* peer_down_flag,
                      const void* stage_top, const void* stage_bottom, const void* flag_top, const void* flag_bottom,
                      unsigned long long target, void* err, void* stream);

/* (6) compositions under the family names of SURVEY.md section 8(b) -------------------------------- */
/* dca_conv3d_igemm: the implicit-GEMM family (convbn_3d / Conv3d s2 / ConvTranspose3d / 1x1x1, submodule.py:121-124,
 * cva.py:16-24): same contract as dca_conv3d_tc. */
int dca_conv3d_igemm(int mode, const void* x, int planes_in, const void* w_tc, const float* scale, const float* shift,
                     const void* res_pre, const void* res_post, int planes_res, const void* up, int planes_up,
                     const void* side, int side_c, void* y, int planes_out, int act, int B, int Cin, int Cout, int Di,
                     int Hi, int Wi, int Do, int Ho, int Wo, void* stream);
/* dca_pool_conv: cva.downsample (cva.py:39-41) = dca_avgpool3d into the caller's `pooled` scratch, then the 3x3x3 conv +
 * folded BN + act on it; y and pooled are [planes][B][(Di+1)/2][(Hi+1)/2][(Wi+1)/2][C]. */
int dca_pool_conv(const void* x, void* pooled, const void* w_tc, const float* scale, const float* shift, void* y,
                  int planes, int act, int B, int C, int Di, int Hi, int Wi, void* stream);
/* dca_softmax_regress_upsample: gwcnet_dca_g.py:238-239 + :117-124 in ONE kernel: softmax over D + disparity regression
 * of a 32 x 8 pixel tile and its 1-pixel ring into shared memory, convex 4x upsampling from there.  pred_q [B,1,H,W] =
 * the 1/4-res disparity, out [B,1,4H,4W].  Bit-identical to dca_softmax_regress + dca_convex_upsample. */
int dca_softmax_regress_upsample(const float* logits, const float* mask, float* pred_q, float* out, int B, int D, int H,
                                 int W, void* stream);

/* F.adaptive_avg_pool3d(x[:, :, h0:, :], (Do, Ho, Wo)) of an fp32 [B][D][H][W] tensor: the `vis_tsne1` head of the
 * plain-GwcNet baseline (models/gwcnet.py:186-190). */
int dca_adaptive_avgpool3d_rows(const float* x, float* y, int B, int D, int H, int W, int h0, int Do, int Ho, int Wo,
                                void* stream);

/* layout + parameter preparation ---------------------------------------------------------------- */
int dca_planes_from_ncdhw(const float* x, void* y, int planes, int B, int C, int Cp, int D, int H, int W,
                          void* stream);
int dca_planes_to_ncdhw(const void* x, int planes, float* y, int B, int C, int Cp, int D, int H, int W, void* stream);
/* cost planes [planes][B][1][H][W][Cp] -> a C-channel slice of a wider fp32 NCHW tensor (batch pitch y_batch_stride
 * elements): torch.cat((l2, l3, l4), dim=1) of feature_extraction.forward (gwcnet_dca_g.py:60) without the copy. */
int dca_planes_to_nchw_slice(const void* x, int planes, float* y, int B, int C, int Cp, int H, int W,
                             long long y_batch_stride, void* stream);
int dca_pack_weights(const float* w, int transposed, int Co, int Ci, int taps, float* out, int CoPad, void* stream);
int dca_fold_bn(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale,
                float* shift, int C, int Cpad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DCA_B200_H */
