// CUDA-core GEMM with fp32 accumulation in strictly ascending-K order per output element.
// This is the arithmetic of the fp32 parity mode (bit-stable, batch-invariant: the reduction order of an output
// element never depends on M or on the batch) and the on-device cross-check for the tcgen05 GEMM.  It is NOT a
// fallback of the bf16 product path.
//   C[b][M,N] = epi(A[b][M,K] . W[N,K]^T)  — replaces every nn.Linear / nn.Conv1d reached from
//   transformers/models/glmasr/modeling_glmasr.py:316-349 and transformers/models/llama/modeling_llama.py:171-289.
#include "common.cuh"
#include "kernels.h"

namespace sonic {

static constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
    float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};
template <> struct Vec4<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float (&o)[4]) {
    uint2 v = *reinterpret_cast<const uint2*>(p);
    o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
    o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
  }
};

template <typename T, typename TC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int b = blockIdx.z;
  const T* A = reinterpret_cast<const T*>(g.A) + (size_t)b * g.a_bstride;
  const T* W = reinterpret_cast<const T*>(g.W);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lr = tid >> 2, lc = (tid & 3) * 4;           // loader: row within tile, k offset

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const bool a_ok = (m0 + lr) < g.M, w_ok = (n0 + lr) < g.N;
  const T* ap = A + (size_t)(m0 + lr) * g.lda + lc;
  const T* wp = W + (size_t)(n0 + lr) * g.ldw + lc;

  for (int k0 = 0; k0 < g.K; k0 += BK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, wv[4] = {0.f, 0.f, 0.f, 0.f};
    if (a_ok) Vec4<T>::load(ap + k0, av);
    if (w_ok) Vec4<T>::load(wp + k0, wv);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) { As[lc + i][lr] = av[i]; Ws[lc + i][lr] = wv[i]; }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], w[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) w[j] = Ws[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
  }

  TC* C = reinterpret_cast<TC*>(g.C) + (size_t)b * g.c_bstride + (size_t)g.c_row0 * g.ldc;
  const TC* R = g.resid ? reinterpret_cast<const TC*>(g.resid) + (size_t)b * g.r_bstride : nullptr;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= g.M) continue;
    float v[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      float x = acc[i][j];
      if (g.bias && n < g.N) x += g.bias[n];
      if (g.act == ACT_GELU) x = gelu_erf(x);
      v[j] = x;
    }
    if (g.act == ACT_SWIGLU) {
#pragma unroll
      for (int j = 0; j < TN; j += 2) {
        const int n = n0 + tx * TN + j;
        if (n + 1 < g.N) C[(size_t)m * g.ldc + (n >> 1)] = from_f32<TC>(silu_exact(v[j]) * v[j + 1]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * TN + j;
        if (n < g.N) {
          float x = v[j];
          if (R) x += to_f32(R[(size_t)m * g.ldr + n]);
          C[(size_t)m * g.ldc + n] = from_f32<TC>(x);
        }
      }
    }
  }
}

template <typename T>
cudaError_t launch_gemm_simt(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return cudaSuccess;
  if (g.K % BK != 0 || g.lda % 4 != 0 || g.ldw % 4 != 0) return cudaErrorInvalidValue;
  dim3 grid(cdiv(g.N, BN), cdiv(g.M, BM), g.batch);
  if (g.out_f32)
    gemm_simt_kernel<T, float><<<grid, 256, 0, st>>>(g);
  else
    gemm_simt_kernel<T, T><<<grid, 256, 0, st>>>(g);
  return cudaGetLastError();
}

template cudaError_t launch_gemm_simt<float>(const GemmArgs&, cudaStream_t);
template cudaError_t launch_gemm_simt<bf16>(const GemmArgs&, cudaStream_t);

}  // namespace sonic
