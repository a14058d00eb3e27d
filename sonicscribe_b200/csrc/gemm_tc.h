// tcgen05 kernel launchers (gemm_tc.cu, attention_tc.cu)
#pragma once
#include <cuda.h>
#include "kernels.h"
namespace sonic {
cudaError_t gemm_tc_init();
cudaError_t gemm_tc_configure();
// swap=false: tokens are the 128-row MMA operand (encoder / prefill). swap=true: weights are (decode, M <= 64 per tile).
cudaError_t launch_gemm_tc(const GemmArgs& g, bool swap, cudaStream_t st);
cudaError_t make_tensor_map_2d(CUtensorMap* map, const void* ptr, long long cols, long long rows, long long ld, int box_cols, int box_rows);
// 2-D uint8 tensor map {cols bytes, rows} (row stride ld bytes), box {box_cols, box_rows}, no swizzle (int8 weight tiles)
cudaError_t make_tensor_map_2d_u8(CUtensorMap* map, const void* ptr, long long cols, long long rows, long long ld, int box_cols, int box_rows);
cudaError_t attention_tc_configure();
// fused-QKV encoder attention (head_dim 64, non-causal): qkv [segments*T, row_width] bf16 -> out [segments*T, out_stride]
cudaError_t launch_attention_tc(const bf16* qkv, int row_width, int q_col, int k_col, int v_col, bf16* out, int out_stride, int segments,
                                int T, int heads, float scale, cudaStream_t st);
cudaError_t attention_prefill_tc_configure();
// decoder prefill attention (causal, GQA, head_dim 128): q from the rotated fused rows, k/v from this layer's KV cache
cudaError_t launch_attention_prefill_tc(const bf16* qkv, int row_width, int total_rows, int q_col0, const bf16* kcache, const bf16* vcache,
                                        int max_batch, int kv_heads, int heads, int max_ctx, const int* tok_off, int batch, int max_q,
                                        bf16* out, int out_stride, float scale, cudaStream_t st);
}  // namespace sonic
