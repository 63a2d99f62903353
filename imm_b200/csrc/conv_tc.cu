// tcgen05 + TMA implicit-GEMM convolution engine (placeholder until the engine lands: nothing eligible).
#include "common.cuh"
namespace immb {
bool conv_tc_eligible(const immb_conv_desc*, int) { return false; }
int conv_tc_fwd(const immb_conv_desc*, const float*, const float*, const float*, const float*, const float*,
                float*, float*, cudaStream_t) { return set_error(IMMB_ERR_UNSUPPORTED, "tc engine not built"); }
int conv_tc_dgrad(const immb_conv_desc*, const float*, const float*, const float*, const float*, float*,
                  cudaStream_t) { return set_error(IMMB_ERR_UNSUPPORTED, "tc engine not built"); }
size_t conv_tc_wgrad_workspace(const immb_conv_desc*) { return 0; }
int conv_tc_wgrad(const immb_conv_desc*, const float*, const float*, const float*, const float*, float*, void*,
                  size_t, cudaStream_t) { return set_error(IMMB_ERR_UNSUPPORTED, "tc engine not built"); }
}  // namespace immb
