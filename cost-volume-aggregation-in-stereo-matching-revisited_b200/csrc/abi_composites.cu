// Two entry points named as SURVEY.md section 8(b) lists them, for callers that bind "one symbol per kernel family":
// dca_conv3d_igemm is the family dispatcher itself, dca_pool_conv runs the pooling kernel and the conv kernel back to
// back (the caller provides the intermediate buffer, so nothing is allocated here).  The third family name,
// dca_softmax_regress_upsample, is a fused kernel (dca_ops.cu).
#include "dca_common.cuh"

extern "C" {
int dca_conv3d_tc(int mode, const void* x, int planes_in, const void* w_tc, const float* scale, const float* shift,
                  const void* res_pre, const void* res_post, int planes_res, const void* up, int planes_up,
                  const void* side, int side_c, void* y, int planes_out, int act, int B, int Cin, int Cout, int Di,
                  int Hi, int Wi, int Do, int Ho, int Wo, void* stream);
int dca_avgpool3d(const void* x, void* y, int planes, int B, int C, int Di, int Hi, int Wi, void* stream);

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

// (dca_softmax_regress_upsample is a fused kernel of its own: dca_ops.cu)
}
