// Entry points named as SURVEY.md section 8(b) lists them, for callers that bind "one symbol per kernel family".
// Each one is a thin composition of the entry points the Python driver calls (same kernels, same arguments); the
// caller provides the intermediate buffer, so nothing is allocated here either.
#include "dca_common.cuh"

extern "C" {
int dca_conv3d_tc(int mode, const void* x, int planes_in, const void* w_tc, const float* scale, const float* shift,
                  const void* res_pre, const void* res_post, int planes_res, const void* up, int planes_up,
                  const void* side, int side_c, void* y, int planes_out, int act, int B, int Cin, int Cout, int Di,
                  int Hi, int Wi, int Do, int Ho, int Wo, void* stream);
int dca_avgpool3d(const void* x, void* y, int planes, int B, int C, int Di, int Hi, int Wi, void* stream);
int dca_softmax_regress(const float* logits, float* pred, int B, int D, int H, int W, void* stream);
int dca_convex_upsample(const float* mask, const float* disp, float* out, int B, int H, int W, void* stream);

// The implicit-GEMM conv family under the survey's name (mode: 0 = 3x3x3 s1, 1 = 3x3x3 s2, 2 = transposed s2, 3 = 1x1x1;
// epilogue: folded BN scale/shift, ReLU / LeakyReLU, residual pointers): same contract as dca_conv3d_tc.
int dca_conv3d_igemm(int mode, const void* x, int planes_in, const void* w_tc, const float* scale, const float* shift,
                     const void* res_pre, const void* res_post, int planes_res, const void* up, int planes_up,
                     const void* side, int side_c, void* y, int planes_out, int act, int B, int Cin, int Cout, int Di,
                     int Hi, int Wi, int Do, int Ho, int Wo, void* stream) {
  return dca_conv3d_tc(mode, x, planes_in, w_tc, scale, shift, res_pre, res_post, planes_res, up, planes_up, side, side_c,
                       y, planes_out, act, B, Cin, Cout, Di, Hi, Wi, Do, Ho, Wo, stream);
}

// cva.downsample (cva.py:39-41): AvgPool3d(3, 2, 1) -> convbn_3d(C, C) -> act.  `pooled` = caller's scratch cost planes
// [planes][B][(Di+1)/2][(Hi+1)/2][(Wi+1)/2][C]; y has the same shape.
int dca_pool_conv(const void* x, void* pooled, const void* w_tc, const float* scale, const float* shift, void* y,
                  int planes, int act, int B, int C, int Di, int Hi, int Wi, void* stream) {
  if (!pooled) return DCA_ERR_ARG;
  const int rc = dca_avgpool3d(x, pooled, planes, B, C, Di, Hi, Wi, stream);
  if (rc != DCA_OK) return rc;
  const int Do = (Di + 1) / 2, Ho = (Hi + 1) / 2, Wo = (Wi + 1) / 2;
  return dca_conv3d_tc(0, pooled, planes, w_tc, scale, shift, nullptr, nullptr, 1, nullptr, 1, nullptr, 0, y, planes, act,
                       B, C, C, Do, Ho, Wo, Do, Ho, Wo, stream);
}

// softmax over disparity + disparity_regression (gwcnet_dca_g.py:238-239) + PropgationNet_4x's convex upsampling
// (gwcnet_dca_g.py:120-124).  logits fp32 [B,D,H,W], mask fp32 channels-last [B,H,W,144]; pred_q [B,1,H,W] is both the
// caller's scratch and the 1/4-res result; out [B,1,4H,4W].
int dca_softmax_regress_upsample(const float* logits, const float* mask, float* pred_q, float* out, int B, int D, int H,
                                 int W, void* stream) {
  if (!pred_q) return DCA_ERR_ARG;
  const int rc = dca_softmax_regress(logits, pred_q, B, D, H, W, stream);
  if (rc != DCA_OK) return rc;
  return dca_convex_upsample(mask, pred_q, out, B, H, W, stream);
}
}
