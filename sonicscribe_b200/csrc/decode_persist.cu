// Persistent greedy-decode step for the Llama decoder (bf16 weights/activations, fp32 accumulation): ONE cooperative
// kernel per generated token runs all 28 layers, the lm_head and the greedy pick, with grid-wide barriers between
// dependent phases instead of ~200 separate launches.
//
// Why: a decode step moves 2.9 GB of weights (0.45 ms at HBM speed) through ~200 tiny dependent operations; as separate
// kernels each one costs ~9-20 us of launch/setup/drain latency (measured).  Here every SM keeps 16 warps
// resident for the whole step; each warp streams 16 weight rows x a K-slice straight from global memory into mma.sync
// fragments (16 B loads, 8 in flight per lane => ~64 KB in flight per SM, which is what saturates HBM; the tensor pipe is
// irrelevant at <= 64 tokens), multiplies them with the (L1-resident) activation slice and writes an fp32 partial.  The
// consumer phase sums the partials in split order (deterministic, batch-invariant) while applying residual / RMSNorm /
// RoPE / SwiGLU, so those never exist as separate passes.
//
// Replaces, for one new token per segment: LlamaDecoderLayer x28 + final norm + lm_head + argmax/EOS bookkeeping
// (transformers/models/llama/modeling_llama.py:53-499, transformers/generation/utils.py:2743-2809).
#include <cooperative_groups.h>
#include "common.cuh"
#include "kernels.h"

namespace sonic {

static constexpr int PH = 2048, PQKV = 3072, PI = 6144, PV_ = 59264, PHD = 128, PKVH = 4, PG = 4;
static constexpr int kPThreads = 512, kPWarps = 16;
static constexpr int kSplitQkv = 16, kSplitO = 16, kSplitGu = 4, kSplitDown = 24, kSplitHead = 2;
static constexpr int AKEYS = 128;                        // keys per attention chunk
static constexpr int kAKRow = PHD * 2 + 16;              // padded K row in shared memory (bytes)

__device__ __forceinline__ uint4 ldg_stream(const void* p) {   // weights: read once, do not pollute L1
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float bf16r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// grid-wide barrier: monotonically increasing arrival counter (zeroed by the host before the launch).  The gpu-scope fences
// order every thread's global writes before the arrival and invalidate L1 after the wait, so plain loads see fresh data.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < epoch);
    __threadfence();
  }
  __syncthreads();
}

// ---- GEMM phase: part[s][tok][n] = sum_{k in slice s} W[n][k] * X[tok][k] ------------------------------------------------
// item = (K-slice s, 16-row block rb); the warps of one CTA take consecutive row blocks of the same slice so that the
// activation slice they all read stays in L1.
template <int NT>
__device__ __forceinline__ void gemm_phase(const bf16* __restrict__ W, int N, int K, int ksplit, const bf16* X, int B, int Bpad,
                                           float* __restrict__ part) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int gw = blockIdx.x * kPWarps + (threadIdx.x >> 5), GW = gridDim.x * kPWarps;
  const int nrb = N >> 4, Ks = K / ksplit, n_items = nrb * ksplit;
  for (int item = gw; item < n_items; item += GW) {
    const int s = item / nrb, rb = item - s * nrb;
    const bf16* w0 = W + (size_t)(rb * 16 + g) * K + (size_t)s * Ks + 8 * t;
    const bf16* w1 = w0 + (size_t)8 * K;
    const bf16* xb = X + (size_t)s * Ks + 8 * t;
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
    for (int k0 = 0; k0 < Ks; k0 += 128) {
      uint4 wa[4], wb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { wa[u] = ldg_stream(w0 + k0 + 32 * u); wb[u] = ldg_stream(w1 + k0 + 32 * u); }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int tok = nt * 8 + g;
          uint4 xv = make_uint4(0, 0, 0, 0);
          if (tok < B) xv = *reinterpret_cast<const uint4*>(xb + (size_t)tok * K + k0 + 32 * u);
          mma16816(acc[nt], wa[u].x, wb[u].x, wa[u].y, wb[u].y, xv.x, xv.y);
          mma16816(acc[nt], wa[u].z, wb[u].z, wa[u].w, wb[u].w, xv.z, xv.w);
        }
      }
    }
    float* p = part + (size_t)s * Bpad * N + rb * 16 + g;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int tok = nt * 8 + 2 * t;
      if (tok < B) { p[(size_t)tok * N] = acc[nt][0]; p[(size_t)tok * N + 8] = acc[nt][2]; }
      if (tok + 1 < B) { p[(size_t)(tok + 1) * N] = acc[nt][1]; p[(size_t)(tok + 1) * N + 8] = acc[nt][3]; }
    }
  }
}

__device__ __forceinline__ void gemm_dispatch(const bf16* W, int N, int K, int ksplit, const bf16* X, int B, int Bpad, float* part) {
  if (B <= 8) gemm_phase<1>(W, N, K, ksplit, X, B, Bpad, part);
  else if (B <= 16) gemm_phase<2>(W, N, K, ksplit, X, B, Bpad, part);
  else if (B <= 32) gemm_phase<4>(W, N, K, ksplit, X, B, Bpad, part);
  else gemm_phase<8>(W, N, K, ksplit, X, B, Bpad, part);
}

// fixed-order sum of the KS split-K partials of one element; fully unrolled so the KS L2 loads are in flight together
template <int KS>
__device__ __forceinline__ float sum_partials(const float* part, size_t stride, size_t idx) {
  float v[KS];
#pragma unroll
  for (int s = 0; s < KS; ++s) v[s] = __ldcg(part + (size_t)s * stride + idx);
  float a = 0.f;
#pragma unroll
  for (int s = 0; s < KS; ++s) a += v[s];
  return a;
}

// ---- row phase: x[b] (+)= sum of partials; u[b] = rmsnorm(x[b]) * gamma  (one CTA per token, 4 features per thread) -------
template <int KS>
__device__ __forceinline__ void residual_norm_phase(const float* part, int B, int Bpad, bf16* x, bf16* u,
                                                    const float* __restrict__ gamma, float eps, float* red) {
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const int c0 = threadIdx.x * 4;
    float v[4];
    const uint2 xr = *reinterpret_cast<const uint2*>(x + (size_t)b * PH + c0);
    v[0] = __uint_as_float(xr.x << 16); v[1] = __uint_as_float(xr.x & 0xffff0000u);
    v[2] = __uint_as_float(xr.y << 16); v[3] = __uint_as_float(xr.y & 0xffff0000u);
    float ss = 0.f;
    float add[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) add[i] = sum_partials<KS>(part, (size_t)Bpad * PH, (size_t)b * PH + c0 + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = bf16r(v[i] + add[i]);
      ss += v[i] * v[i];
    }
    __nv_bfloat162 o0 = __floats2bfloat162_rn(v[0], v[1]), o1 = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(x + (size_t)b * PH + c0) = make_uint2(*reinterpret_cast<uint32_t*>(&o0), *reinterpret_cast<uint32_t*>(&o1));
    const float rstd = rsqrtf(block_sum(ss, red) / PH + eps);
    const float4 gm = *reinterpret_cast<const float4*>(gamma + c0);
    __nv_bfloat162 u0 = __floats2bfloat162_rn(gm.x * bf16r(v[0] * rstd), gm.y * bf16r(v[1] * rstd));
    __nv_bfloat162 u1 = __floats2bfloat162_rn(gm.z * bf16r(v[2] * rstd), gm.w * bf16r(v[3] * rstd));
    *reinterpret_cast<uint2*>(u + (size_t)b * PH + c0) = make_uint2(*reinterpret_cast<uint32_t*>(&u0), *reinterpret_cast<uint32_t*>(&u1));
    __syncthreads();
  }
}

// ---- attention phase: item = (segment, kv head): finish q/k/v from the partials, RoPE, append to the cache, attend ---------
__device__ __forceinline__ void attention_phase(const DecodePersistArgs& a, const DecLayerDev& L, uint8_t* smem) {
  uint8_t* sK = smem;                                          // AKEYS * kAKRow
  bf16* sV = reinterpret_cast<bf16*>(smem + AKEYS * kAKRow);   // AKEYS * 128
  float* sQ = reinterpret_cast<float*>(smem + AKEYS * kAKRow + AKEYS * PHD * 2);   // [4][128]
  float* sP = sQ + PG * PHD;                                   // [4][AKEYS]
  float* sKV = sP + PG * AKEYS;                                // new k (128) | new v (128)
  float* sRed = sKV + 2 * PHD;                                 // [4 heads][4 key groups] max, then sums
  float* sState = sRed + 32;                                   // m[4], l[4], corr[4]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t pstride = (size_t)a.Bpad * PQKV;
  for (int item = blockIdx.x; item < a.B * PKVH; item += gridDim.x) {
    const int seg = item / PKVH, kvh = item - seg * PKVH;
    const int pos = a.gs.ctx_len[seg], kv_len = pos + 1;
    bf16* kc = L.kc + ((size_t)seg * PKVH + kvh) * a.max_ctx * PHD;
    bf16* vc = L.vc + ((size_t)seg * PKVH + kvh) * a.max_ctx * PHD;
    // q (4 heads), k, v of the new token: sum the split-K partials, round to bf16 like the unfused path, rotate
    for (int i = tid; i < (PG + 2) * (PHD / 2); i += kPThreads) {
      const int hh = i / (PHD / 2), j = i - hh * (PHD / 2);     // hh < 4: query head; 4: key; 5: value (pair j, j+64)
      const int col = (hh < PG) ? (kvh * PG + hh) * PHD : (hh == PG ? (16 + kvh) * PHD : (16 + PKVH + kvh) * PHD);
      const float x = bf16r(sum_partials<kSplitQkv>(a.part, pstride, (size_t)seg * PQKV + col + j));
      const float y = bf16r(sum_partials<kSplitQkv>(a.part, pstride, (size_t)seg * PQKV + col + j + PHD / 2));
      if (hh <= PG) {
        const float c = bf16r(a.cos_t[(size_t)pos * (PHD / 2) + j]), s = bf16r(a.sin_t[(size_t)pos * (PHD / 2) + j]);
        const float rx = bf16r(x * c - y * s), ry = bf16r(y * c + x * s);
        if (hh < PG) { sQ[hh * PHD + j] = rx * a.scale; sQ[hh * PHD + j + PHD / 2] = ry * a.scale; }
        else { sKV[j] = rx; sKV[j + PHD / 2] = ry; }
      } else { sKV[PHD + j] = x; sKV[PHD + j + PHD / 2] = y; }
    }
    if (tid < PG) { sState[tid] = -INFINITY; sState[4 + tid] = 0.f; }
    __syncthreads();
    if (tid < PHD) kc[(size_t)pos * PHD + tid] = __float2bfloat16_rn(sKV[tid]);
    else if (tid < 2 * PHD) vc[(size_t)pos * PHD + tid - PHD] = __float2bfloat16_rn(sKV[tid]);
    const int head = warp & 3, kgrp = warp >> 2;                 // scores: 4 heads x 4 groups of 32 keys
    float o_acc = 0.f;                                           // PV: thread = (head = tid / 128, dim = tid % 128)
    for (int k0 = 0; k0 < kv_len; k0 += AKEYS) {
      const int nk = min(AKEYS, kv_len - k0);
      __syncthreads();                                           // previous chunk fully consumed; new k/v row written
      for (int i = tid; i < AKEYS * (PHD / 8); i += kPThreads) {
        const int r = i / (PHD / 8), c8 = i - r * (PHD / 8);
        uint4 kk = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
        if (r < nk) {
          if (k0 + r == pos) {                                   // the row appended above by this CTA: take it from smem
            uint32_t wk[4], wv[4];
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
              __nv_bfloat162 pk = __floats2bfloat162_rn(sKV[c8 * 8 + 2 * e2], sKV[c8 * 8 + 2 * e2 + 1]);
              __nv_bfloat162 pv = __floats2bfloat162_rn(sKV[PHD + c8 * 8 + 2 * e2], sKV[PHD + c8 * 8 + 2 * e2 + 1]);
              wk[e2] = *reinterpret_cast<uint32_t*>(&pk); wv[e2] = *reinterpret_cast<uint32_t*>(&pv);
            }
            kk = make_uint4(wk[0], wk[1], wk[2], wk[3]); vv = make_uint4(wv[0], wv[1], wv[2], wv[3]);
          } else {
            kk = __ldcg(reinterpret_cast<const uint4*>(kc + (size_t)(k0 + r) * PHD + c8 * 8));
            vv = __ldcg(reinterpret_cast<const uint4*>(vc + (size_t)(k0 + r) * PHD + c8 * 8));
          }
        }
        *reinterpret_cast<uint4*>(sK + r * kAKRow + c8 * 16) = kk;
        *reinterpret_cast<uint4*>(sV + r * PHD + c8 * 8) = vv;
      }
      __syncthreads();
      // scores for (head, key = kgrp*32 + lane)
      const int r = kgrp * 32 + lane;
      float sc = -INFINITY;
      if (r < nk) {
        float acc = 0.f;
#pragma unroll 4
        for (int c8 = 0; c8 < PHD / 8; ++c8) {
          const uint4 kv = *reinterpret_cast<const uint4*>(sK + r * kAKRow + c8 * 16);
          const float4 q0 = *reinterpret_cast<const float4*>(sQ + head * PHD + c8 * 8);
          const float4 q1 = *reinterpret_cast<const float4*>(sQ + head * PHD + c8 * 8 + 4);
          acc = fmaf(__uint_as_float(kv.x << 16), q0.x, acc); acc = fmaf(__uint_as_float(kv.x & 0xffff0000u), q0.y, acc);
          acc = fmaf(__uint_as_float(kv.y << 16), q0.z, acc); acc = fmaf(__uint_as_float(kv.y & 0xffff0000u), q0.w, acc);
          acc = fmaf(__uint_as_float(kv.z << 16), q1.x, acc); acc = fmaf(__uint_as_float(kv.z & 0xffff0000u), q1.y, acc);
          acc = fmaf(__uint_as_float(kv.w << 16), q1.z, acc); acc = fmaf(__uint_as_float(kv.w & 0xffff0000u), q1.w, acc);
        }
        sc = acc;
      }
      const float wmax = warp_max(sc);
      if (lane == 0) sRed[head * 4 + kgrp] = wmax;
      __syncthreads();
      const float cmax = fmaxf(fmaxf(sRed[head * 4], sRed[head * 4 + 1]), fmaxf(sRed[head * 4 + 2], sRed[head * 4 + 3]));
      const float m_old = sState[head];
      const float m_new = fmaxf(m_old, cmax);
      const float p = (sc == -INFINITY) ? 0.f : expf(sc - m_new);
      sP[head * AKEYS + r] = p;
      const float wsum = warp_sum(p);
      __syncthreads();                                           // every thread has read the maxima and the running state
      if (lane == 0) sRed[head * 4 + kgrp] = wsum;
      __syncthreads();
      if (kgrp == 0 && lane == 0) {                              // one thread per head advances the online-softmax state
        const float corr = (m_old == -INFINITY) ? 0.f : expf(m_old - m_new);
        const float lsum = sRed[head * 4] + sRed[head * 4 + 1] + sRed[head * 4 + 2] + sRed[head * 4 + 3];
        sState[8 + head] = corr;
        sState[head] = m_new;
        sState[4 + head] = sState[4 + head] * corr + lsum;
      }
      __syncthreads();
      {
        const int ph = tid >> 7, d = tid & 127;
        float acc = o_acc * sState[8 + ph];
        const float* pp = sP + ph * AKEYS;
        for (int j = 0; j < nk; ++j) acc = fmaf(pp[j], __bfloat162float(sV[j * PHD + d]), acc);
        o_acc = acc;
      }
    }
    __syncthreads();
    {
      const int ph = tid >> 7, d = tid & 127;
      a.attn[(size_t)seg * PH + (size_t)(kvh * PG + ph) * PHD + d] = __float2bfloat16_rn(o_acc / sState[4 + ph]);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kPThreads, 1) decode_persist_kernel(DecodePersistArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ float red[32];
  unsigned epoch = 0;
  const int tid = threadIdx.x;
  const int B = a.B, Bpad = a.Bpad;

  // ---- phase 0: x = E[cur_tok]; u = rmsnorm(x) * g(layer 0 input norm)
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const int tok = a.gs.cur_tok[b];
    const int c0 = tid * 4;
    const uint2 e = *reinterpret_cast<const uint2*>(a.embed + (size_t)tok * PH + c0);
    *reinterpret_cast<uint2*>(a.x + (size_t)b * PH + c0) = e;
    float v[4] = {__uint_as_float(e.x << 16), __uint_as_float(e.x & 0xffff0000u), __uint_as_float(e.y << 16), __uint_as_float(e.y & 0xffff0000u)};
    const float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
    const float rstd = rsqrtf(block_sum(ss, red) / PH + a.eps);
    const float4 gm = *reinterpret_cast<const float4*>(a.layers[0].rms1 + c0);
    __nv_bfloat162 u0 = __floats2bfloat162_rn(gm.x * bf16r(v[0] * rstd), gm.y * bf16r(v[1] * rstd));
    __nv_bfloat162 u1 = __floats2bfloat162_rn(gm.z * bf16r(v[2] * rstd), gm.w * bf16r(v[3] * rstd));
    *reinterpret_cast<uint2*>(a.u + (size_t)b * PH + c0) = make_uint2(*reinterpret_cast<uint32_t*>(&u0), *reinterpret_cast<uint32_t*>(&u1));
    __syncthreads();
  }
  grid_barrier(a.bar, epoch);

  for (int l = 0; l < a.n_layers; ++l) {
    const DecLayerDev L = a.layers[l];
    gemm_dispatch(L.wqkv, PQKV, PH, kSplitQkv, a.u, B, Bpad, a.part);
    grid_barrier(a.bar, epoch);
    attention_phase(a, L, smem);
    grid_barrier(a.bar, epoch);
    gemm_dispatch(L.wo, PH, PH, kSplitO, a.attn, B, Bpad, a.part);
    grid_barrier(a.bar, epoch);
    residual_norm_phase<kSplitO>(a.part, B, Bpad, a.x, a.u, L.rms2, a.eps, red);
    grid_barrier(a.bar, epoch);
    gemm_dispatch(L.wgu, 2 * PI, PH, kSplitGu, a.u, B, Bpad, a.part);
    grid_barrier(a.bar, epoch);
    {  // SwiGLU over the interleaved (gate, up) columns
      const size_t stride = (size_t)Bpad * 2 * PI;
      const int total = B * PI;
      for (int i = blockIdx.x * kPThreads + tid; i < total; i += gridDim.x * kPThreads) {
        const int b = i / PI, j = i - b * PI;
        const float gte = sum_partials<kSplitGu>(a.part, stride, (size_t)b * 2 * PI + 2 * j);
        const float up = sum_partials<kSplitGu>(a.part, stride, (size_t)b * 2 * PI + 2 * j + 1);
        a.act[(size_t)b * PI + j] = __float2bfloat16_rn(silu(gte) * up);
      }
    }
    grid_barrier(a.bar, epoch);
    gemm_dispatch(L.wdown, PH, PI, kSplitDown, a.act, B, Bpad, a.part);
    grid_barrier(a.bar, epoch);
    residual_norm_phase<kSplitDown>(a.part, B, Bpad, a.x, a.u, (l + 1 < a.n_layers) ? a.layers[l + 1].rms1 : a.final_norm, a.eps, red);
    grid_barrier(a.bar, epoch);
  }

  // ---- lm_head + greedy pick
  gemm_dispatch(a.lm_head, PV_, PH, kSplitHead, a.u, B, Bpad, a.part);
  grid_barrier(a.bar, epoch);
  {
    __shared__ float s_best[kPWarps], s_second[kPWarps];
    __shared__ int s_idx[kPWarps];
    const size_t stride = (size_t)Bpad * PV_;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
      float best = -INFINITY, second = -INFINITY;
      int bi = 0x7fffffff;
      for (int i = tid; i < PV_; i += kPThreads) {
        const float v = __ldcg(a.part + (size_t)b * PV_ + i) + __ldcg(a.part + stride + (size_t)b * PV_ + i);
        if (a.logits_out) a.logits_out[(size_t)b * PV_ + i] = v;
        if (v > best) { second = best; best = v; bi = i; }
        else if (v > second) second = v;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { second = fmaxf(fmaxf(second, os), best); best = ob; bi = oi; }
        else { second = fmaxf(second, ob); }
      }
      if ((tid & 31) == 0) { s_best[tid >> 5] = best; s_second[tid >> 5] = second; s_idx[tid >> 5] = bi; }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < kPWarps; ++w) {
          const float ob = s_best[w], os = s_second[w];
          const int oi = s_idx[w];
          if (ob > best || (ob == best && oi < bi)) { second = fmaxf(fmaxf(second, os), best); best = ob; bi = oi; }
          else { second = fmaxf(second, ob); }
        }
        const int step = *a.gs.step;
        a.gs.ctx_len[b] += 1;
        if (!a.gs.finished[b]) {
          a.gs.out_ids[(size_t)b * a.gs.max_new + step] = bi;
          if (a.gs.margins) a.gs.margins[(size_t)b * a.gs.max_new + step] = best - second;
          a.gs.n_out[b] = step + 1;
          bool eos = false;
          for (int e = 0; e < a.gs.n_eos; ++e) eos |= (bi == a.gs.eos[e]);
          if (eos || step + 1 >= a.gs.max_new) { a.gs.finished[b] = 1; atomicSub(a.gs.n_unfinished, 1); }
        }
        a.gs.cur_tok[b] = bi;
      }
      __syncthreads();
    }
  }
  grid_barrier(a.bar, epoch);
  if (blockIdx.x == 0 && tid == 0) *a.gs.step += 1;
}

size_t decode_persist_smem_bytes() { return (size_t)AKEYS * kAKRow + (size_t)AKEYS * PHD * 2 + (PG * PHD + PG * AKEYS + 2 * PHD + 32 + 16) * 4; }

size_t decode_persist_part_floats(int Bpad) {
  size_t m = (size_t)kSplitQkv * PQKV;
  if ((size_t)kSplitO * PH > m) m = (size_t)kSplitO * PH;
  if ((size_t)kSplitGu * 2 * PI > m) m = (size_t)kSplitGu * 2 * PI;
  if ((size_t)kSplitDown * PH > m) m = (size_t)kSplitDown * PH;
  if ((size_t)kSplitHead * PV_ > m) m = (size_t)kSplitHead * PV_;
  return m * Bpad;
}

cudaError_t decode_persist_configure() {
  return cudaFuncSetAttribute(decode_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)decode_persist_smem_bytes());
}

cudaError_t launch_decode_persist(const DecodePersistArgs& a, int num_sms, cudaStream_t st) {
  SONIC_CUDA_TRY(cudaMemsetAsync(a.bar, 0, sizeof(unsigned), st));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(num_sms); cfg.blockDim = dim3(kPThreads); cfg.dynamicSmemBytes = decode_persist_smem_bytes(); cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, decode_persist_kernel, a);
}

}  // namespace sonic
