// Softmax attention on CUDA cores, fp32 math (T = storage type).
//   attention_simt_kernel : many queries per segment (encoder non-causal full attention, decoder causal GQA prefill).
//                           Arithmetic of the fp32 parity mode; also serves the (small) bf16 prefill attention.
//   attention_decode_kernel: one query per segment against the KV cache (greedy decode step), keys split over warps.
// Replaces F.scaled_dot_product_attention as reached from transformers/integrations/sdpa_attention.py:40-104 for
// GlmAsrAttention (modeling_glmasr.py:175-225, no mask, scale 64^-1/2) and LlamaAttention (modeling_llama.py:225-289,
// causal, 16 query / 4 kv heads, scale 128^-1/2).
#include "common.cuh"
#include "kernels.h"

namespace sonic {

static constexpr int QB = 16;        // queries per CTA (2 per warp)
static constexpr int KT = 32;        // keys per tile (one per lane)

template <typename T, int HD>
__global__ void __launch_bounds__(256) attention_simt_kernel(AttnArgs a) {
  constexpr int DPL = HD / 32;       // output dims per lane
  __shared__ float sK[KT][HD + 1];
  __shared__ float sV[KT][HD];
  __shared__ float sQ[QB][HD];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QB;
  const int q_first = a.q_off ? a.q_off[b] : b * a.q_len_fixed;
  const int q_len = a.q_off ? (a.q_off[b + 1] - a.q_off[b]) : a.q_len_fixed;
  const int kv_len = a.kv_len ? a.kv_len[b] : a.kv_len_fixed;
  if (q0 >= q_len) return;
  const int kvh = h / (a.heads / a.kv_heads);
  const T* Q = reinterpret_cast<const T*>(a.q);
  const T* K = reinterpret_cast<const T*>(a.k) + (size_t)b * a.k_seg_stride + (size_t)kvh * a.k_head_stride;
  const T* V = reinterpret_cast<const T*>(a.v) + (size_t)b * a.v_seg_stride + (size_t)kvh * a.v_head_stride;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < QB * HD; i += 256) {
    const int r = i / HD, d = i - r * HD;
    const int qi = q0 + r;
    sQ[r][d] = (qi < q_len) ? to_f32(Q[(size_t)(q_first + qi) * a.q_row_stride + h * HD + d]) * a.scale : 0.f;
  }
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  float o[2][DPL];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int d = 0; d < DPL; ++d) o[r][d] = 0.f;

  // last key any query of this CTA may see
  const int q_last = min(q0 + QB, q_len) - 1;
  const int k_end = a.causal ? min(kv_len, kv_len - q_len + q_last + 1) : kv_len;

  for (int k0 = 0; k0 < k_end; k0 += KT) {
    __syncthreads();
    for (int i = tid; i < KT * HD; i += 256) {
      const int j = i / HD, d = i - j * HD;
      const int kj = k0 + j;
      float kv = 0.f, vv = 0.f;
      if (kj < kv_len) {
        kv = to_f32(K[(size_t)kj * a.k_tok_stride + d]);
        vv = to_f32(V[(size_t)kj * a.v_tok_stride + d]);
      }
      sK[j][d] = kv;
      sV[j][d] = vv;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int qr = warp * 2 + r, qi = q0 + qr;
      if (qi >= q_len) continue;                         // warp-uniform
      const int kj = k0 + lane;
      const int limit = a.causal ? (kv_len - q_len + qi) : (kv_len - 1);
      float s = 0.f;
#pragma unroll 8
      for (int d = 0; d < HD; ++d) s = fmaf(sQ[qr][d], sK[lane][d], s);
      if (kj > limit) s = -INFINITY;
      const float mn = fmaxf(m[r], warp_max(s));
      const float p = (s == -INFINITY) ? 0.f : expf(s - mn);
      const float corr = (m[r] == -INFINITY) ? 0.f : expf(m[r] - mn);
      l[r] = l[r] * corr + warp_sum(p);
      m[r] = mn;
#pragma unroll
      for (int d = 0; d < DPL; ++d) o[r][d] *= corr;
      for (int j = 0; j < KT; ++j) {
        const float pj = __shfl_sync(0xffffffffu, p, j);
#pragma unroll
        for (int d = 0; d < DPL; ++d) o[r][d] = fmaf(pj, sV[j][lane + 32 * d], o[r][d]);
      }
    }
  }
  T* O = reinterpret_cast<T*>(a.o);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qi = q0 + warp * 2 + r;
    if (qi >= q_len) continue;
    const float inv = 1.0f / l[r];
#pragma unroll
    for (int d = 0; d < DPL; ++d)
      O[(size_t)(q_first + qi) * a.o_row_stride + h * HD + lane + 32 * d] = from_f32<T>(o[r][d] * inv);
  }
}

// one CTA (4 warps) per (segment, query head); keys strided over warps, lanes over keys for QK, lanes over dims for PV
template <typename T, int HD>
__global__ void __launch_bounds__(128) attention_decode_kernel(AttnArgs a) {
  constexpr int DPL = HD / 32;
  __shared__ float sQ[HD];
  __shared__ float sM[4], sL[4], sO[4][HD];
  const int b = blockIdx.y, h = blockIdx.x;
  const int kv_len = a.kv_len[b] + 1;                    // the step's own key was appended at index ctx_len
  const int kvh = h / (a.heads / a.kv_heads);
  const T* Q = reinterpret_cast<const T*>(a.q) + (size_t)b * a.q_row_stride + h * HD;
  const T* K = reinterpret_cast<const T*>(a.k) + (size_t)b * a.k_seg_stride + (size_t)kvh * a.k_head_stride;
  const T* V = reinterpret_cast<const T*>(a.v) + (size_t)b * a.v_seg_stride + (size_t)kvh * a.v_head_stride;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int d = tid; d < HD; d += 128) sQ[d] = to_f32(Q[d]) * a.scale;
  __syncthreads();
  float m = -INFINITY, l = 0.f, o[DPL];
#pragma unroll
  for (int d = 0; d < DPL; ++d) o[d] = 0.f;
  for (int k0 = warp * 32; k0 < kv_len; k0 += 128) {
    const int kj = k0 + lane;
    float s = -INFINITY;
    if (kj < kv_len) {
      const T* kr = K + (size_t)kj * a.k_tok_stride;
      float acc = 0.f;
#pragma unroll 8
      for (int d = 0; d < HD; ++d) acc = fmaf(sQ[d], to_f32(kr[d]), acc);
      s = acc;
    }
    const float mn = fmaxf(m, warp_max(s));
    const float p = (s == -INFINITY) ? 0.f : expf(s - mn);
    const float corr = (m == -INFINITY) ? 0.f : expf(m - mn);
    l = l * corr + warp_sum(p);
    m = mn;
#pragma unroll
    for (int d = 0; d < DPL; ++d) o[d] *= corr;
    const int nj = min(32, kv_len - k0);
    for (int j = 0; j < nj; ++j) {
      const float pj = __shfl_sync(0xffffffffu, p, j);
      const T* vr = V + (size_t)(k0 + j) * a.v_tok_stride;
#pragma unroll
      for (int d = 0; d < DPL; ++d) o[d] = fmaf(pj, to_f32(vr[lane + 32 * d]), o[d]);
    }
  }
  if (lane == 0) { sM[warp] = m; sL[warp] = l; }
#pragma unroll
  for (int d = 0; d < DPL; ++d) sO[warp][lane + 32 * d] = o[d];
  __syncthreads();
  // fixed-order combine of the 4 partial softmaxes
  float M = fmaxf(fmaxf(sM[0], sM[1]), fmaxf(sM[2], sM[3]));
  float L = 0.f;
  float w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { w[i] = (sM[i] == -INFINITY) ? 0.f : expf(sM[i] - M); L += sL[i] * w[i]; }
  T* O = reinterpret_cast<T*>(a.o) + (size_t)b * a.o_row_stride + h * HD;
  for (int d = tid; d < HD; d += 128) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc += sO[i][d] * w[i];
    O[d] = from_f32<T>(acc / L);
  }
}

template <typename T>
cudaError_t launch_attention_simt(const AttnArgs& a, cudaStream_t st) {
  if (a.batch <= 0) return cudaSuccess;
  if (a.decode) {
    dim3 grid(a.heads, a.batch);
    if (a.hd == 128) attention_decode_kernel<T, 128><<<grid, 128, 0, st>>>(a);
    else if (a.hd == 64) attention_decode_kernel<T, 64><<<grid, 128, 0, st>>>(a);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
  }
  dim3 grid(cdiv(a.max_q, QB), a.heads, a.batch);
  if (a.hd == 128) attention_simt_kernel<T, 128><<<grid, 256, 0, st>>>(a);
  else if (a.hd == 64) attention_simt_kernel<T, 64><<<grid, 256, 0, st>>>(a);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}
template cudaError_t launch_attention_simt<float>(const AttnArgs&, cudaStream_t);
template cudaError_t launch_attention_simt<bf16>(const AttnArgs&, cudaStream_t);

}  // namespace sonic
