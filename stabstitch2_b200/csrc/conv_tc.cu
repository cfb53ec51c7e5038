// tcgen05 tensor-core implicit-GEMM convolution (placeholder until the UMMA kernel lands).
#include "common.cuh"

int conv_tc_prepare(ss2_ctx* ctx, ConvLayer& L) {
  (void)ctx; (void)L;
  return SS2_OK;
}

int conv_tc_launch(ss2_ctx* ctx, const ConvLayer& L, const float* d_in, int B, int D, int H, int W, float* d_out,
                   const float* d_residual, int relu, cudaStream_t st, bool* handled) {
  (void)ctx; (void)L; (void)d_in; (void)B; (void)D; (void)H; (void)W; (void)d_out; (void)d_residual; (void)relu; (void)st;
  *handled = false;
  return SS2_OK;
}
