// tcgen05 GEMM launcher (gemm_tc.cu)
#pragma once
#include "kernels.h"
namespace sonic {
cudaError_t gemm_tc_init();
cudaError_t gemm_tc_configure();
// swap=false: tokens are the 128-row MMA operand (encoder / prefill). swap=true: weights are (decode, M <= 64 per tile).
cudaError_t launch_gemm_tc(const GemmArgs& g, bool swap, cudaStream_t st);
}  // namespace sonic
