// tcgen05 implicit-GEMM convolution (placeholder until the UMMA pipeline lands: every layer is declined, so the
// caller uses the mma.sync pipeline).
#include "conv.cuh"

namespace dyf {
bool conv_umma_eligible(const ConvParams&) { return false; }
int launch_conv_umma(const ConvParams&, cudaStream_t) { return 0; }
}  // namespace dyf
