// Greedy-decode attention step for the Llama decoder (bf16 storage, fp32 math), one launch per layer:
//   RoPE(q, k_new) + KV-cache append + split-KV attention over the cache + deterministic combine.
// Replaces, for one new token per segment, apply_rotary_pos_emb + DynamicCache.update + SDPA of
// transformers/models/llama/modeling_llama.py:225-289 (16 query heads, 4 KV heads, head_dim 128, scale 128^-1/2).
//
// grid = (key chunks of 64, 4 KV heads, segments), 128 threads: warp w serves query head 4*kvh + w, so the four query
// heads of a GQA group read each K/V chunk once (16 B coalesced loads into shared memory).  Every CTA writes an
// (m, l, o) partial; the CTA that arrives last at the group's counter merges the partials in chunk order.
#include "common.cuh"
#include "kernels.h"

namespace sonic {

static constexpr int DH = 128, DCH = 64, DG = 4;        // head dim, keys per chunk, query heads per KV head
static constexpr int kKRow = DH * 2 + 16;                // padded K row (bytes): conflict-free 16 B reads, one row per lane

__global__ void __launch_bounds__(128)
decode_attn_kernel(DecodeAttnArgs a) {
  __shared__ __align__(16) uint8_t sK[DCH * kKRow];
  __shared__ __align__(16) bf16 sV[DCH * DH];
  __shared__ __align__(16) float sQ[DG][DH];
  __shared__ float sP[DG][DCH];
  __shared__ float sM[DG], sL[DG];
  __shared__ int s_last;

  pdl_launch_dependents();
  pdl_wait();
  const int chunk = blockIdx.x, kvh = blockIdx.y, seg = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pos = a.ctx_len[seg];                       // position of the token being decoded
  const int kv_len = pos + 1;
  const int n_chunks = (kv_len + DCH - 1) / DCH;
  const int heads = a.kv_heads * DG;
  const int width = (heads + 2 * a.kv_heads) * DH;
  const bf16* row = a.qkv + (size_t)seg * width;
  bf16* kc = a.kcache + ((size_t)seg * a.kv_heads + kvh) * a.max_ctx * DH;
  bf16* vc = a.vcache + ((size_t)seg * a.kv_heads + kvh) * a.max_ctx * DH;
  const int group = seg * a.kv_heads + kvh;
  float* ws_o = a.ws + ((size_t)group * a.max_chunks + chunk) * DG * (DH + 2);

  if (chunk < n_chunks) {
    // ---- RoPE of the four query heads (fp32 math on bf16 inputs, bf16-rounded like the stored q of the prefill path)
    for (int i = tid; i < DG * (DH / 2); i += 128) {
      const int hq = i / (DH / 2), j = i - hq * (DH / 2);
      const float c = __bfloat162float(__float2bfloat16_rn(a.cos_t[(size_t)pos * (DH / 2) + j]));
      const float s = __bfloat162float(__float2bfloat16_rn(a.sin_t[(size_t)pos * (DH / 2) + j]));
      const bf16* q = row + (size_t)(kvh * DG + hq) * DH;
      const float x = __bfloat162float(q[j]), y = __bfloat162float(q[j + DH / 2]);
      sQ[hq][j] = __bfloat162float(__float2bfloat16_rn(x * c - y * s)) * a.scale;
      sQ[hq][j + DH / 2] = __bfloat162float(__float2bfloat16_rn(y * c + x * s)) * a.scale;
    }
    // ---- the CTA owning the chunk of `pos` appends the new (rotated) key and the new value to the cache
    if (chunk == pos / DCH) {
      const bf16* kn = row + (size_t)(heads + kvh) * DH;
      const bf16* vn = row + (size_t)(heads + a.kv_heads + kvh) * DH;
      if (tid < DH / 2) {
        const float c = __bfloat162float(__float2bfloat16_rn(a.cos_t[(size_t)pos * (DH / 2) + tid]));
        const float s = __bfloat162float(__float2bfloat16_rn(a.sin_t[(size_t)pos * (DH / 2) + tid]));
        const float x = __bfloat162float(kn[tid]), y = __bfloat162float(kn[tid + DH / 2]);
        kc[(size_t)pos * DH + tid] = __float2bfloat16_rn(x * c - y * s);
        kc[(size_t)pos * DH + tid + DH / 2] = __float2bfloat16_rn(y * c + x * s);
      } else {
        const int d = (tid - DH / 2) * 2;
        vc[(size_t)pos * DH + d] = vn[d];
        vc[(size_t)pos * DH + d + 1] = vn[d + 1];
      }
    }
    __syncthreads();
    // ---- stage the K / V chunk in shared memory
    const int k0 = chunk * DCH;
    const int nk = min(DCH, kv_len - k0);
    for (int i = tid; i < DCH * (DH / 8); i += 128) {          // 16 B pieces
      const int r = i / (DH / 8), c8 = i - r * (DH / 8);
      uint4 kvv = make_uint4(0, 0, 0, 0), vvv = make_uint4(0, 0, 0, 0);
      if (r < nk) {
        kvv = *reinterpret_cast<const uint4*>(kc + (size_t)(k0 + r) * DH + c8 * 8);
        vvv = *reinterpret_cast<const uint4*>(vc + (size_t)(k0 + r) * DH + c8 * 8);
      }
      *reinterpret_cast<uint4*>(sK + r * kKRow + c8 * 16) = kvv;
      *reinterpret_cast<uint4*>(sV + r * DH + c8 * 8) = vvv;
    }
    __syncthreads();
    // ---- scores: warp = query head, lane = keys (lane, lane + 32)
    float sc[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int r = lane + 32 * t;
      float acc = 0.f;
#pragma unroll 4
      for (int c8 = 0; c8 < DH / 8; ++c8) {
        const uint4 kv = *reinterpret_cast<const uint4*>(sK + r * kKRow + c8 * 16);
        const float4 q0 = *reinterpret_cast<const float4*>(&sQ[warp][c8 * 8]);
        const float4 q1 = *reinterpret_cast<const float4*>(&sQ[warp][c8 * 8 + 4]);
        acc = fmaf(__uint_as_float(kv.x << 16), q0.x, acc); acc = fmaf(__uint_as_float(kv.x & 0xffff0000u), q0.y, acc);
        acc = fmaf(__uint_as_float(kv.y << 16), q0.z, acc); acc = fmaf(__uint_as_float(kv.y & 0xffff0000u), q0.w, acc);
        acc = fmaf(__uint_as_float(kv.z << 16), q1.x, acc); acc = fmaf(__uint_as_float(kv.z & 0xffff0000u), q1.y, acc);
        acc = fmaf(__uint_as_float(kv.w << 16), q1.z, acc); acc = fmaf(__uint_as_float(kv.w & 0xffff0000u), q1.w, acc);
      }
      sc[t] = (r < nk) ? acc : -INFINITY;
    }
    const float m = warp_max(fmaxf(sc[0], sc[1]));
    const float p0 = (sc[0] == -INFINITY) ? 0.f : expf(sc[0] - m);
    const float p1 = (sc[1] == -INFINITY) ? 0.f : expf(sc[1] - m);
    const float l = warp_sum(p0 + p1);
    sP[warp][lane] = p0;
    sP[warp][lane + 32] = p1;
    if (lane == 0) { sM[warp] = m; sL[warp] = l; }
    __syncthreads();
    // ---- O partial: thread = (head pair, 2 dims)
    {
      const int hp = tid >> 6, d = (tid & 63) * 2;
      float o00 = 0.f, o01 = 0.f, o10 = 0.f, o11 = 0.f;
      for (int j = 0; j < nk; ++j) {
        const uint32_t vv = *reinterpret_cast<const uint32_t*>(sV + j * DH + d);
        const float v0 = __uint_as_float(vv << 16), v1 = __uint_as_float(vv & 0xffff0000u);
        const float pa = sP[2 * hp][j], pb = sP[2 * hp + 1][j];
        o00 = fmaf(pa, v0, o00); o01 = fmaf(pa, v1, o01);
        o10 = fmaf(pb, v0, o10); o11 = fmaf(pb, v1, o11);
      }
      float* oa = ws_o + (size_t)(2 * hp) * (DH + 2);
      float* ob = ws_o + (size_t)(2 * hp + 1) * (DH + 2);
      oa[d] = o00; oa[d + 1] = o01;
      ob[d] = o10; ob[d + 1] = o11;
    }
    if (tid < DG) { ws_o[(size_t)tid * (DH + 2) + DH] = sM[tid]; ws_o[(size_t)tid * (DH + 2) + DH + 1] = sL[tid]; }
  }
  // ---- arrival; the last CTA of the (segment, kv head) group merges the partials in chunk order
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int prev = atomicAdd(a.counters + group, 1);
    s_last = (prev == (int)gridDim.x - 1) ? 1 : 0;
    if (s_last) a.counters[group] = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* wg = a.ws + (size_t)group * a.max_chunks * DG * (DH + 2);
  for (int hq = 0; hq < DG; ++hq) {
    float M = -INFINITY;
    for (int c = 0; c < n_chunks; ++c) M = fmaxf(M, __ldcg(wg + ((size_t)c * DG + hq) * (DH + 2) + DH));
    float L = 0.f, acc = 0.f;
    for (int c = 0; c < n_chunks; ++c) {
      const float* pc = wg + ((size_t)c * DG + hq) * (DH + 2);
      const float w = expf(__ldcg(pc + DH) - M);
      L += __ldcg(pc + DH + 1) * w;
      acc += __ldcg(pc + tid) * w;
    }
    a.out[(size_t)seg * heads * DH + (size_t)(kvh * DG + hq) * DH + tid] = __float2bfloat16_rn(acc / L);
  }
}

cudaError_t launch_decode_attn(const DecodeAttnArgs& a, int batch, int n_chunks, cudaStream_t st, bool pdl) {
  if (batch <= 0) return cudaSuccess;
  if (n_chunks < 1 || n_chunks > a.max_chunks) return cudaErrorInvalidValue;
  dim3 grid(n_chunks, a.kv_heads, batch);
  return launch_ex(decode_attn_kernel, grid, dim3(128), 0, st, pdl, a);
}

}  // namespace sonic
