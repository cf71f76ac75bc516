// placeholder until the tcgen05 kernel lands
#include "common.cuh"
#include "../../include/holo_b200.h"
extern "C" int holo_conv3d_tc(const void* x_hi, const void* x_lo, int Cin, int D, int H, int W, int ksize, const void* w_hi,
                   const void* w_lo, const float* bias, const float* residual, int Cout, float* out,
                   void* out_hi_bf16, void* out_lo_bf16, void* stream) {
    holo_set_error("holo_conv3d_tc: not built yet");
    return HOLO_ERR_UNSUPPORTED;
}
