// placeholder until the tcgen05 implicit-GEMM lands (see include/dca_b200.h)
#include "dca_common.cuh"
extern "C" int dca_conv3d_tc(int, const void*, int, const void*, const float*, const float*, const void*, const void*,
                             int, void*, int, int, int, int, int, int, int, int, int, int, int, void*) {
  return DCA_ERR_UNSUPPORTED;
}
extern "C" int dca_pack_weights_tc(const float*, int, int, int, int, void*, int, void*) { return DCA_ERR_UNSUPPORTED; }
extern "C" long long dca_pack_weights_tc_bytes(int, int, int, int) { return 0; }
