// bf16 GEMM on the 5th-generation tensor cores (tcgen05) for sm_100a:
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> shared-memory ring -> tcgen05.mma (single issuing thread,
//   fp32 accumulators in TMEM) -> tcgen05.ld epilogue (bias / GELU / SwiGLU / residual / dtype) -> global.
//
//   normal mode: D[m,n] = A[m,:].W[n,:]   A = activations [M,K] (tile 128 rows), W = weights [N,K] (tile BN rows)
//   swap mode  : D[f,t] = W[f,:].X[t,:]   the 128-row MMA operand is the WEIGHT tile, the N operand the (few) token
//                rows; the epilogue stores D transposed so the output is still [tokens, features].  This is the
//                weight-streaming shape of the greedy decode step (HBM-bound, tokens <= 64).
//
// Replaces the cuBLAS bf16 GEMMs behind every nn.Linear / nn.Conv1d of GlmAsrEncoder, the projector and the Llama
// decoder (transformers/models/glmasr/modeling_glmasr.py:175-349, transformers/models/llama/modeling_llama.py:171-289).
// Warp roles (256 threads): w0 TMA producer, w1 MMA issuer, w2 TMEM allocator, w4-7 epilogue (TMEM lane quadrants).
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "kernels.h"
#include "gemm_tc.h"
#include "tc_ptx.cuh"

namespace sonic {

static constexpr int BM = 128;          // MMA M (rows of the A-side operand per CTA)
static constexpr int BK = 64;           // 64 bf16 = 128 B = one swizzle atom row
static constexpr int UMMA_K = 16;

template <int BN, bool SWAP = false, bool W8 = false> struct TcCfg {
  // bytes per ring stage of the two TMA-loaded operands.  With W8 the weight operand (A when SWAP, else B; always 128
  // rows) arrives as int8 (64 B rows, no swizzle) and is expanded to bf16 into one of kConv SW128 buffers by warps 4-7.
  static constexpr int kStageA = (W8 && SWAP) ? BM * BK : BM * BK * 2;
  static constexpr int kStageB = (W8 && !SWAP) ? BN * BK : BN * BK * 2;
  static constexpr int kConv = W8 ? 2 : 0;
  static constexpr int kConvBytes = 128 * BK * 2;             // 16 KB
  // stage count is a launch-time choice: "two CTAs per SM" (<= ~110 KB each) when the grid has more CTAs than SMs, otherwise
  // the whole 227 KB of one SM so a lone CTA keeps as many TMA loads in flight as possible (HBM latency ~2 us per box)
  static constexpr int kStagesDual = (BN >= 128 ? 3 : (BN <= 32 ? 6 : 4));
  static constexpr int kMaxStages = 12;
  static constexpr int kStagesSolo = ((227 * 1024 - 1024 - 512) / (kStageA + kStageB)) > kMaxStages ? kMaxStages : ((227 * 1024 - 1024 - 512) / (kStageA + kStageB));
  static constexpr int smem_bytes(int stages) { return stages * (kStageA + kStageB) + kConv * kConvBytes + 1024 /*align slack*/ + 512 /*barriers*/; }
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
};

struct TcEpi {
  void* C; long long ldc; long long c_bstride; long long c_row0;
  const float* bias;
  const void* resid; long long ldr; long long r_bstride;
  int M, N, K;            // logical problem: M tokens, N features
  int act;
  // implicit conv1d (k=3, pad 1): K = 3*cin is walked tap by tap; tap t reads A-map coordinates
  // {tap_col[t] + kc*64, a0 + tap_row[t]} (see launch_gemm_tc).  kb_per_tap = K/64 for a plain GEMM.
  int kb_per_tap;
  int tap_col[3];
  int tap_row[3];
  // split-K (swap mode): gridDim.z CTAs share one output tile; fp32 partials go to `ws`, the CTA that arrives last at
  // `counters[tile]` sums them in split order (deterministic) and runs the epilogue.
  int splits;
  float* ws;
  int* counters;
  int stages;             // shared-memory ring depth chosen at launch
  const float* wscale;    // W8: per-output-feature dequantisation scale (absmax / 127)
  // fused partial RoPE of the encoder QKV projection (normal mode): for columns < rope_ncols, the first 32 dims of every
  // 64-dim head are rotated (pairs j, j+16) with the angle table [rope_T][16] indexed by (row % rope_T)
  const float* rope_cos; const float* rope_sin; int rope_T; int rope_ncols;
};

template <typename TC>
__device__ __forceinline__ void store_chunk32(TC* dst, const float (&x)[32]);
template <>
__device__ __forceinline__ void store_chunk32<bf16>(bf16* dst, const float (&x)[32]) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 o;
    __nv_bfloat162 p0 = __floats2bfloat162_rn(x[8 * i + 0], x[8 * i + 1]);
    __nv_bfloat162 p1 = __floats2bfloat162_rn(x[8 * i + 2], x[8 * i + 3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(x[8 * i + 4], x[8 * i + 5]);
    __nv_bfloat162 p3 = __floats2bfloat162_rn(x[8 * i + 6], x[8 * i + 7]);
    o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
    o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
    d[i] = o;
  }
}
template <>
__device__ __forceinline__ void store_chunk32<float>(float* dst, const float (&x)[32]) {
  float4* d = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
}

template <int BN, bool SWAP, typename TC, bool W8>
__global__ void __launch_bounds__(256, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcEpi e) {
  using Cfg = TcCfg<BN, SWAP, W8>;
  static_assert(!W8 || SWAP || BN == 128, "int8 weight tiles are 128 rows");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* sgen = smem_raw + (base - raw);
  const int kStages = e.stages;
  const uint32_t sA = base, sB = base + kStages * Cfg::kStageA;
  const uint32_t sConv = sB + kStages * Cfg::kStageB;         // W8: bf16 images of the weight tile
  const uint32_t bars = sConv + Cfg::kConv * Cfg::kConvBytes; // full[kStages], empty[kStages], tmem_full, conv_full[2], conv_empty[2], tmem slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + (bars - base) + 8 * (2 * kStages + 5));
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * kStages);
  auto conv_full_bar = [&](int c) { return bars + 8u * (2 * kStages + 1 + c); };
  auto conv_empty_bar = [&](int c) { return bars + 8u * (2 * kStages + 3 + c); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a0 = blockIdx.y * BM;            // first row of the 128-row operand (tokens, or features when SWAP)
  const int b0 = blockIdx.x * BN;            // first row of the BN-row operand (features, or tokens when SWAP)
  const int batch = SWAP ? 0 : blockIdx.z;
  const int split = SWAP ? blockIdx.z : 0;
  const int total_kb = e.K / BK;
  const int kb_begin = SWAP ? (split * total_kb) / e.splits : 0;
  const int kb_end = SWAP ? ((split + 1) * total_kb) / e.splits : total_kb;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    if constexpr (W8) {
      for (int c = 0; c < 2; ++c) { mbar_init(conv_full_bar(c), 128); mbar_init(conv_empty_bar(c), 1); }
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  pdl_launch_dependents();
  if (warp == 0) {
    if (elect_one_sync()) {
      int s = 0; uint32_t ph = 0;
      int kb = kb_begin;
      if constexpr (SWAP) {
        // the weight operand does not depend on the predecessor kernel: stream the first ring of weight tiles before the
        // dependency wait, then add the activation tiles of those stages
        const int pre = min(kb_end - kb_begin, kStages);
        for (int i = 0; i < pre; ++i) {
          mbar_expect_tx(full_bar(i), Cfg::kStageA + Cfg::kStageB);
          tma_load_3d(sA + i * Cfg::kStageA, &tmA, full_bar(i), (kb_begin + i) * BK, a0, 0);
        }
        pdl_wait();
        for (int i = 0; i < pre; ++i) tma_load_3d(sB + i * Cfg::kStageB, &tmB, full_bar(i), (kb_begin + i) * BK, b0, 0);
        kb = kb_begin + pre;
        s = pre % kStages;
        ph = (pre == kStages) ? 1u : 0u;
      } else {
        pdl_wait();
      }
      for (; kb < kb_end; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), Cfg::kStageA + Cfg::kStageB);
        const int tap = kb / e.kb_per_tap, kc = kb - tap * e.kb_per_tap;
        tma_load_3d(sA + s * Cfg::kStageA, &tmA, full_bar(s), (SWAP ? kb : kc) * BK + (SWAP ? 0 : e.tap_col[tap]),
                    a0 + (SWAP ? 0 : e.tap_row[tap]), SWAP ? 0 : batch);
        tma_load_3d(sB + s * Cfg::kStageB, &tmB, full_bar(s), kb * BK, b0, SWAP ? batch : 0);
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      int s = 0; uint32_t ph = 0;
      int c = 0; uint32_t cph = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(full_bar(s), ph);
        if constexpr (W8) mbar_wait(conv_full_bar(c), cph);
        tc_fence_after();
        const uint64_t da = make_sw128_desc((W8 && SWAP) ? sConv + c * Cfg::kConvBytes : sA + s * Cfg::kStageA);
        const uint64_t db = make_sw128_desc((W8 && !SWAP) ? sConv + c * Cfg::kConvBytes : sB + s * Cfg::kStageB);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advancing 16 bf16 = 32 B along K inside the 128 B swizzle atom: +2 in the (addr>>4) field
          tc_mma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > kb_begin || k != 0) ? 1u : 0u);
        }
        tc_commit(empty_bar(s));                 // frees the smem stage once these MMAs retire
        if constexpr (W8) {
          tc_commit(conv_empty_bar(c));
          if (++c == 2) { c = 0; cph ^= 1u; }
        }
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
      tc_commit(tmem_full_bar);                  // accumulator complete
    }
  } else if (warp >= 4) {
    const int q = warp - 4;                      // TMEM lane quadrant == warp % 4
    if constexpr (W8) {
      // int8 -> bf16 expansion of the weight tile: thread r owns weight row r (64 int8 = 64 B in, 128 B out in the
      // K-major SWIZZLE_128B operand layout: 16 B chunk c8 of row r lives at chunk (c8 ^ (r & 7)))
      const int r = threadIdx.x - 128;
      const uint32_t raw0 = SWAP ? sA : sB;
      const int raw_stride = SWAP ? Cfg::kStageA : Cfg::kStageB;
      int s = 0; uint32_t ph = 0;
      int c = 0; uint32_t cph = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(full_bar(s), ph);
        mbar_wait(conv_empty_bar(c), cph ^ 1u);
        const uint8_t* src = sgen + (raw0 - base) + s * raw_stride + r * 64;
        uint8_t* dst = sgen + (sConv - base) + c * Cfg::kConvBytes + r * 128;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 w = *reinterpret_cast<const uint4*>(src + 16 * i);
          const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
          uint32_t o[8];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int b0i = (int)(int8_t)(ww[t] & 0xff), b1i = (int)(int8_t)((ww[t] >> 8) & 0xff);
            const int b2i = (int)(int8_t)((ww[t] >> 16) & 0xff), b3i = (int)(int8_t)(ww[t] >> 24);
            __nv_bfloat162 lo = __floats2bfloat162_rn((float)b0i, (float)b1i);
            __nv_bfloat162 hi = __floats2bfloat162_rn((float)b2i, (float)b3i);
            o[2 * t] = *reinterpret_cast<uint32_t*>(&lo);
            o[2 * t + 1] = *reinterpret_cast<uint32_t*>(&hi);
          }
          *reinterpret_cast<uint4*>(dst + (((2 * i) ^ (r & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(dst + (((2 * i + 1) ^ (r & 7)) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
        }
        fence_proxy_async_smem();
        mbar_arrive(conv_full_bar(c));
        if (++c == 2) { c = 0; cph ^= 1u; }
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
    pdl_wait();                                  // residual / output buffers belong to the predecessor until it completes
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    TC* Cb = reinterpret_cast<TC*>(e.C) + (size_t)batch * e.c_bstride + (size_t)e.c_row0 * e.ldc;
    const TC* Rb = e.resid ? reinterpret_cast<const TC*>(e.resid) + (size_t)batch * e.r_bstride : nullptr;
    if constexpr (!SWAP) {
      const int m = a0 + q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(trow + c * 32, v);
        tmem_ld_wait();
        const int n = b0 + c * 32;
        if (m < e.M && n < e.N) {
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
          if constexpr (W8) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 sv = __ldg(reinterpret_cast<const float4*>(e.wscale + n + j));
              x[j] *= sv.x; x[j + 1] *= sv.y; x[j + 2] *= sv.z; x[j + 3] *= sv.w;
            }
          }
          if (e.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(e.bias + n + j));
              x[j] += bv.x; x[j + 1] += bv.y; x[j + 2] += bv.z; x[j + 3] += bv.w;
            }
          }
          if (e.act == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = gelu_erf(x[j]);
          }
          if (e.rope_cos != nullptr && n < e.rope_ncols && (n & 63) == 0) {
            const int pos = m % e.rope_T;
            const float4* cp = reinterpret_cast<const float4*>(e.rope_cos + (size_t)pos * 16);
            const float4* sp = reinterpret_cast<const float4*>(e.rope_sin + (size_t)pos * 16);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const float4 c4 = __ldg(cp + q4), s4 = __ldg(sp + q4);
              const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const int j = 4 * q4 + t;
                // the reference casts cos/sin to the model dtype (modeling_glmasr.py:109)
                const float c = __bfloat162float(__float2bfloat16_rn(cc[t])), s = __bfloat162float(__float2bfloat16_rn(ss[t]));
                const float a = x[j], b = x[j + 16];
                x[j] = a * c - b * s;
                x[j + 16] = b * c + a * s;
              }
            }
          }
          if (e.act == ACT_SWIGLU) {
            float y[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) y[j] = silu(x[2 * j]) * x[2 * j + 1];
            TC* dst = Cb + (size_t)m * e.ldc + (n >> 1);
            if constexpr (sizeof(TC) == 2) {
              uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(y[8 * i + 0], y[8 * i + 1]);
                __nv_bfloat162 p1 = __floats2bfloat162_rn(y[8 * i + 2], y[8 * i + 3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(y[8 * i + 4], y[8 * i + 5]);
                __nv_bfloat162 p3 = __floats2bfloat162_rn(y[8 * i + 6], y[8 * i + 7]);
                uint4 o;
                o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
                o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
                d[i] = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) dst[j] = from_f32<TC>(y[j]);
            }
          } else {
            if (Rb) {
              const TC* r = Rb + (size_t)m * e.ldr + n;
              if constexpr (sizeof(TC) == 2) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const uint4 rv = *reinterpret_cast<const uint4*>(r + 8 * i);
                  const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                  for (int t = 0; t < 4; ++t) {
                    x[8 * i + 2 * t] += __uint_as_float(w[t] << 16);
                    x[8 * i + 2 * t + 1] += __uint_as_float(w[t] & 0xffff0000u);
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] += to_f32(r[j]);
              }
            }
            store_chunk32<TC>(Cb + (size_t)m * e.ldc + n, x);
          }
        }
      }
    } else {
      // rows of D are features, columns are tokens; store transposed
      const int f = a0 + q * 32 + lane;
      const int ntok = min(BN, e.M - b0);
      const float bias = (e.bias && f < e.N) ? e.bias[f] : 0.f;
      const float wsc = (W8 && f < e.N) ? e.wscale[f] : 1.f;
      const float* wsum = nullptr;                 // != null: this CTA reduces the split-K partials
      if (e.splits > 1) {
        __shared__ int s_last;
        const int tile = blockIdx.y * gridDim.x + blockIdx.x;
        float* wsp = e.ws + ((size_t)(tile * e.splits + split) * BN) * BM + q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN / 16; ++c) {
          uint32_t v[16];
          tmem_ld16(trow + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c * 16 + j < ntok) wsp[(size_t)(c * 16 + j) * BM] = __uint_as_float(v[j]);
        }
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 128) {
          const int prev = atomicAdd(e.counters + tile, 1);
          s_last = (prev == e.splits - 1) ? 1 : 0;
          if (s_last) e.counters[tile] = 0;        // ready for the next launch
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (s_last) {
          __threadfence();
          wsum = e.ws + ((size_t)(tile * e.splits) * BN) * BM + q * 32 + lane;
        }
      }
      if (e.splits == 1 || wsum != nullptr) {
#pragma unroll 1
      for (int c = 0; c < BN / 16; ++c) {
        uint32_t v[16];
        if (wsum == nullptr) {
          tmem_ld16(trow + c * 16, v);
          tmem_ld_wait();
        } else {
          // sum the partials in split order; 64 independent L2 loads are in flight per thread (latency-bound otherwise)
          float acc[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll 1
          for (int z0 = 0; z0 < e.splits; z0 += 4) {
            float tmp[4][16];
#pragma unroll
            for (int zz = 0; zz < 4; ++zz)
#pragma unroll
              for (int j = 0; j < 16; ++j)
                tmp[zz][j] = (z0 + zz < e.splits && c * 16 + j < ntok) ? __ldcg(wsum + ((size_t)(z0 + zz) * BN + c * 16 + j) * BM) : 0.f;
#pragma unroll
            for (int zz = 0; zz < 4; ++zz)
#pragma unroll
              for (int j = 0; j < 16; ++j) acc[j] += tmp[zz][j];
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int t = c * 16 + j;
          float x = __uint_as_float(v[j]) * wsc + bias;
          if (e.act == ACT_GELU) x = gelu_erf(x);
          if (e.act == ACT_SWIGLU) {
            const float other = __shfl_xor_sync(0xffffffffu, x, 1);
            if (t < ntok && f < e.N && (lane & 1) == 0) Cb[(size_t)(b0 + t) * e.ldc + (f >> 1)] = from_f32<TC>(silu(x) * other);
          } else if (t < ntok && f < e.N) {
            if (Rb) x += to_f32(Rb[(size_t)(b0 + t) * e.ldr + f]);
            Cb[(size_t)(b0 + t) * e.ldc + f] = from_f32<TC>(x);
          }
        }
      }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols) : "memory");
  }
}

// ---- persistent 128 x 256 tile kernel for the token-major (non-swap) bf16 GEMMs of the encoder and the prefill ----------------
// A 128 x 128 tile makes tcgen05.mma read 8 KB of shared memory per 64 clocks, which is the SM's whole shared-memory
// bandwidth; with N = 256 it is 12 KB per 128 clocks.  One CTA per SM walks tiles (n fastest, so the A rows of a tile row
// stay in L2), the 4-stage TMA ring runs across tile boundaries, and the two 256-column TMEM accumulators alternate so the
// 8 epilogue warps drain tile i while the tensor core is already working on tile i + 1.
// Warp roles (384 threads): w0 TMA producer, w1 MMA issuer, w2 TMEM owner, w4..11 epilogue (lane quadrant = warp % 4, column
// half = (warp - 4) / 4).
static constexpr int PBN = 256;
static constexpr int kPStages = 4;
static constexpr int kPStageA = BM * BK * 2, kPStageB = PBN * BK * 2;
static constexpr int kPersistGemmSmem = kPStages * (kPStageA + kPStageB) + 1024 + 256;
static constexpr int kPersistGemmThreads = 384;

template <typename TC>
__device__ __forceinline__ void epilogue_chunk32(const TcEpi& e, float (&x)[32], int m, int n, TC* Cb, const TC* Rb) {
  if (e.wscale) {                                                       // weights were int8 rows expanded to bf16: per-feature dequantisation scale
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 sv = __ldg(reinterpret_cast<const float4*>(e.wscale + n + j));
      x[j] *= sv.x; x[j + 1] *= sv.y; x[j + 2] *= sv.z; x[j + 3] *= sv.w;
    }
  }
  if (e.bias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(e.bias + n + j));
      x[j] += bv.x; x[j + 1] += bv.y; x[j + 2] += bv.z; x[j + 3] += bv.w;
    }
  }
  if (e.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = gelu_erf(x[j]);
  }
  if (e.rope_cos != nullptr && n < e.rope_ncols && (n & 63) == 0) {
    const int pos = m % e.rope_T;
    const float4* cp = reinterpret_cast<const float4*>(e.rope_cos + (size_t)pos * 16);
    const float4* sp = reinterpret_cast<const float4*>(e.rope_sin + (size_t)pos * 16);
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const float4 c4 = __ldg(cp + q4), s4 = __ldg(sp + q4);
      const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j = 4 * q4 + t;
        const float c = __bfloat162float(__float2bfloat16_rn(cc[t])), sn = __bfloat162float(__float2bfloat16_rn(ss[t]));
        const float av = x[j], bv = x[j + 16];
        x[j] = av * c - bv * sn;
        x[j + 16] = bv * c + av * sn;
      }
    }
  }
  if (e.act == ACT_SWIGLU) {
    float y[32];
#pragma unroll
    for (int j = 0; j < 16; ++j) y[j] = silu(x[2 * j]) * x[2 * j + 1];
    TC* dst = Cb + (size_t)m * e.ldc + (n >> 1);
    if constexpr (sizeof(TC) == 2) {
      uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(y[8 * i + 0], y[8 * i + 1]);
        __nv_bfloat162 p1 = __floats2bfloat162_rn(y[8 * i + 2], y[8 * i + 3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(y[8 * i + 4], y[8 * i + 5]);
        __nv_bfloat162 p3 = __floats2bfloat162_rn(y[8 * i + 6], y[8 * i + 7]);
        uint4 o;
        o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
        o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
        d[i] = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) dst[j] = from_f32<TC>(y[j]);
    }
  } else {
    if (Rb) {
      const TC* r = Rb + (size_t)m * e.ldr + n;
      if constexpr (sizeof(TC) == 2) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 rv = *reinterpret_cast<const uint4*>(r + 8 * i);
          const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            x[8 * i + 2 * t] += __uint_as_float(w[t] << 16);
            x[8 * i + 2 * t + 1] += __uint_as_float(w[t] & 0xffff0000u);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] += to_f32(r[j]);
      }
    }
    store_chunk32<TC>(Cb + (size_t)m * e.ldc + n, x);
  }
}

template <typename TC>
__global__ void __launch_bounds__(kPersistGemmThreads, 1)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcEpi e, int tiles_m, int tiles_n,
                       int batches) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + kPStages * kPStageA;
  const uint32_t bars = sB + kPStages * kPStageB;            // full[4], empty[4], acc_full[2], acc_empty[2], tmem slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + (bars - base) + 8 * (2 * kPStages + 4));
  auto full_bar = [&](uint32_t s) { return bars + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bars + 8u * (kPStages + s); };
  auto acc_full = [&](uint32_t a) { return bars + 8u * (2 * kPStages + a); };
  auto acc_empty = [&](uint32_t a) { return bars + 8u * (2 * kPStages + 2 + a); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_kb = e.K / BK;
  const int n_tiles = tiles_m * tiles_n * batches;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kPStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(acc_full(a), 1); mbar_init(acc_empty(a), 8); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one_sync()) {
      uint32_t cnt = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int batch = tile / (tiles_m * tiles_n), r = tile - batch * tiles_m * tiles_n;
        const int a0 = (r / tiles_n) * BM, b0 = (r % tiles_n) * PBN;
        for (int kb = 0; kb < total_kb; ++kb, ++cnt) {
          const uint32_t s = cnt % kPStages, ph = (cnt / kPStages) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), kPStageA + kPStageB);
          const int tap = kb / e.kb_per_tap, kc = kb - tap * e.kb_per_tap;
          tma_load_3d(sA + s * kPStageA, &tmA, full_bar(s), kc * BK + e.tap_col[tap], a0 + e.tap_row[tap], batch);
          tma_load_3d(sB + s * kPStageB, &tmB, full_bar(s), kb * BK, b0, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, PBN);
      uint32_t cnt = 0, it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1u, aph = (it >> 1) & 1u;
        mbar_wait(acc_empty(acc), aph ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < total_kb; ++kb, ++cnt) {
          const uint32_t s = cnt % kPStages, ph = (cnt / kPStages) & 1u;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint64_t da = make_sw128_desc(sA + s * kPStageA), db = make_sw128_desc(sB + s * kPStageB);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            tc_mma_bf16(tmem_base + acc * PBN, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb != 0 || k != 0) ? 1u : 0u);
          tc_commit(empty_bar(s));
        }
        tc_commit(acc_full(acc));
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3, half = (warp - 4) >> 2;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int batch = tile / (tiles_m * tiles_n), r = tile - batch * tiles_m * tiles_n;
      const int a0 = (r / tiles_n) * BM, b0 = (r % tiles_n) * PBN;
      const uint32_t acc = it & 1u, aph = (it >> 1) & 1u;
      mbar_wait(acc_full(acc), aph);
      tc_fence_after();
      const int m = a0 + q * 32 + lane;
      TC* Cb = reinterpret_cast<TC*>(e.C) + (size_t)batch * e.c_bstride + (size_t)e.c_row0 * e.ldc;
      const TC* Rb = e.resid ? reinterpret_cast<const TC*>(e.resid) + (size_t)batch * e.r_bstride : nullptr;
#pragma unroll 1
      for (int c = 0; c < PBN / 64; ++c) {
        const int col = half * (PBN / 2) + c * 32;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * PBN + col, v);
        tmem_ld_wait();
        const int n = b0 + col;
        if (m < e.M && n < e.N) {
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
          epilogue_chunk32<TC>(e, x, m, n, Cb, Rb);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(acc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// ---- host side --------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

cudaError_t gemm_tc_init() {
  if (g_encode) return cudaSuccess;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  SONIC_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
  g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  return cudaSuccess;
}

// 3-D bf16 tensor map {K, rows, batch} with a {64, box_rows, 1} box and 128B swizzle; OOB reads return zeros.
static cudaError_t make_map(CUtensorMap* map, const void* ptr, long long K, long long rows, long long ld, long long batch,
                            long long bstride, int box_rows) {
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)(batch > 0 ? batch : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(bstride > 0 ? bstride : ld * rows) * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// 3-D int8 tensor map {K, rows, 1} with a {64, box_rows, 1} box, no swizzle (the converter warps read it row by row)
static cudaError_t make_map_i8(CUtensorMap* map, const void* ptr, long long K, long long rows, long long ld, int box_rows) {
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, 1};
  cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)ld * rows};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// 2-D bf16 tensor map {cols, rows} (row stride ld elements) with a {box_cols, box_rows} box and 128B swizzle
cudaError_t make_tensor_map_2d(CUtensorMap* map, const void* ptr, long long cols, long long rows, long long ld, int box_cols, int box_rows) {
  SONIC_CUDA_TRY(gemm_tc_init());
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t make_tensor_map_2d_u8(CUtensorMap* map, const void* ptr, long long cols, long long rows, long long ld, int box_cols, int box_rows) {
  SONIC_CUDA_TRY(gemm_tc_init());
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int BN, bool SWAP, typename TC, bool W8 = false>
static cudaError_t launch_one(const CUtensorMap& ma, const CUtensorMap& mb, const TcEpi& e, dim3 grid, cudaStream_t st, bool pdl = false) {
  using Cfg = TcCfg<BN, SWAP, W8>;
  TcEpi ee = e;
  const long long ctas = (long long)grid.x * grid.y * grid.z;
  // measured (scripts/bench_gemm.py): a deeper ring does not speed up the weight stream — one SM sustains ~40 GB/s of
  // DRAM-missing TMA traffic whatever the ring depth — so the two-CTAs-per-SM depth is used throughout
  (void)ctas;
  ee.stages = Cfg::kStagesDual;
  return launch_ex(gemm_tc_kernel<BN, SWAP, TC, W8>, grid, dim3(256), (size_t)Cfg::smem_bytes(ee.stages), st, pdl, ma, mb, ee);
}

template <int BN, bool SWAP, typename TC, bool W8 = false>
static cudaError_t configure_one() {
  using Cfg = TcCfg<BN, SWAP, W8>;
  return cudaFuncSetAttribute(gemm_tc_kernel<BN, SWAP, TC, W8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              Cfg::smem_bytes(SWAP ? Cfg::kStagesSolo : Cfg::kStagesDual));
}
// opt every instantiation into its dynamic shared memory up front (must not happen lazily inside a stream capture)
cudaError_t gemm_tc_configure() {
  SONIC_CUDA_TRY(cudaFuncSetAttribute(gemm_tc_persist_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPersistGemmSmem));
  SONIC_CUDA_TRY(cudaFuncSetAttribute(gemm_tc_persist_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPersistGemmSmem));
  SONIC_CUDA_TRY((configure_one<128, false, bf16>()));
  SONIC_CUDA_TRY((configure_one<128, false, float>()));
  SONIC_CUDA_TRY((configure_one<16, true, bf16>()));
  SONIC_CUDA_TRY((configure_one<16, true, float>()));
  SONIC_CUDA_TRY((configure_one<32, true, bf16>()));
  SONIC_CUDA_TRY((configure_one<32, true, float>()));
  SONIC_CUDA_TRY((configure_one<64, true, bf16>()));
  SONIC_CUDA_TRY((configure_one<64, true, float>()));
  SONIC_CUDA_TRY((configure_one<128, false, bf16, true>()));
  SONIC_CUDA_TRY((configure_one<16, true, bf16, true>()));
  SONIC_CUDA_TRY((configure_one<32, true, bf16, true>()));
  SONIC_CUDA_TRY((configure_one<64, true, bf16, true>()));
  return cudaSuccess;
}

static constexpr int kMaxTiles = 2048;

static bool persist_gemm_enabled() {
  static const bool on = [] { const char* v = getenv("SONIC_GEMM_PERSIST"); return !(v && v[0] == '0'); }();
  return on;
}
static int persist_gemm_sms() {
  static const int n = [] {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms > 0 ? sms : 148;
  }();
  return n;
}

int gemm_tc_pick_bn(const GemmArgs& g, bool swap) {
  if (swap) return g.M <= 16 ? 16 : (g.M <= 32 ? 32 : 64);
  return 128;
}

cudaError_t launch_gemm_tc(const GemmArgs& g, bool swap, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return cudaSuccess;
  SONIC_CUDA_TRY(gemm_tc_init());
  if (g.K % BK != 0 || g.lda % 8 != 0 || g.ldw % 16 != 0 || g.N % 32 != 0) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.W) & 15)) return cudaErrorInvalidValue;
  TcEpi e;
  e.C = g.C; e.ldc = g.ldc; e.c_bstride = g.c_bstride; e.c_row0 = g.c_row0;
  e.bias = g.bias; e.resid = g.resid; e.ldr = g.ldr; e.r_bstride = g.r_bstride;
  e.M = g.M; e.N = g.N; e.K = g.K; e.act = g.act;
  e.kb_per_tap = g.K / BK;
  e.splits = 1; e.ws = nullptr; e.counters = nullptr; e.stages = 0;
  e.wscale = g.wscale;
  e.rope_cos = swap ? nullptr : g.rope_cos; e.rope_sin = g.rope_sin; e.rope_T = g.rope_T; e.rope_ncols = g.rope_ncols;
  const bool w8 = g.w_int8 != 0;
  if (w8 && (g.out_f32 || g.conv_cin > 0 || !g.wscale)) return cudaErrorInvalidValue;
  for (int t = 0; t < 3; ++t) { e.tap_col[t] = 0; e.tap_row[t] = 0; }
  CUtensorMap ma, mb;
  const int bn = gemm_tc_pick_bn(g, swap);
  if (!swap && g.conv_cin > 0) {
    // A = padded time-major input [batch][rows_pad, cin]; output row m reads input rows stride*m + {0,1,2}
    if (g.conv_cin % BK != 0 || g.K != 3 * g.conv_cin) return cudaErrorInvalidValue;
    e.kb_per_tap = g.conv_cin / BK;
    if (g.conv_stride == 1) {
      SONIC_CUDA_TRY(make_map(&ma, g.A, g.conv_cin, g.conv_rows_pad, g.conv_cin, g.batch, g.a_bstride, BM));
      for (int t = 0; t < 3; ++t) e.tap_row[t] = t;
    } else if (g.conv_stride == 2) {
      if (g.conv_rows_pad % 2) return cudaErrorInvalidValue;
      SONIC_CUDA_TRY(make_map(&ma, g.A, 2 * g.conv_cin, g.conv_rows_pad / 2, 2 * g.conv_cin, g.batch, g.a_bstride, BM));
      e.tap_col[1] = g.conv_cin;
      e.tap_row[2] = 1;
    } else return cudaErrorInvalidValue;
    SONIC_CUDA_TRY(make_map(&mb, g.W, g.K, g.N, g.ldw, 1, 0, bn));
    dim3 grid(cdiv(g.N, bn), cdiv(g.M, BM), g.batch);
    return g.out_f32 ? launch_one<128, false, float>(ma, mb, e, grid, st) : launch_one<128, false, bf16>(ma, mb, e, grid, st);
  }
  if (!swap && !w8 && g.N % 128 == 0 && persist_gemm_enabled()) {
    // token-major bf16 GEMM: persistent 128 x 256 tiles, one CTA per SM
    SONIC_CUDA_TRY(make_map(&ma, g.A, g.K, g.M, g.lda, g.batch, g.a_bstride, BM));
    SONIC_CUDA_TRY(make_map(&mb, g.W, g.K, g.N, g.ldw, 1, 0, PBN));
    const int tiles_m = cdiv(g.M, BM), tiles_n = cdiv(g.N, PBN);
    const long long total = (long long)tiles_m * tiles_n * g.batch;
    const int grid = (int)(total < persist_gemm_sms() ? total : persist_gemm_sms());
    if (g.out_f32) gemm_tc_persist_kernel<float><<<grid, kPersistGemmThreads, kPersistGemmSmem, st>>>(ma, mb, e, tiles_m, tiles_n, g.batch);
    else gemm_tc_persist_kernel<bf16><<<grid, kPersistGemmThreads, kPersistGemmSmem, st>>>(ma, mb, e, tiles_m, tiles_n, g.batch);
    return cudaGetLastError();
  }
  if (!swap) {
    SONIC_CUDA_TRY(make_map(&ma, g.A, g.K, g.M, g.lda, g.batch, g.a_bstride, BM));
    if (w8) SONIC_CUDA_TRY(make_map_i8(&mb, g.W, g.K, g.N, g.ldw, bn));
    else SONIC_CUDA_TRY(make_map(&mb, g.W, g.K, g.N, g.ldw, 1, 0, bn));
    dim3 grid(cdiv(g.N, bn), cdiv(g.M, BM), g.batch);
    if (w8) return launch_one<128, false, bf16, true>(ma, mb, e, grid, st);
    return g.out_f32 ? launch_one<128, false, float>(ma, mb, e, grid, st) : launch_one<128, false, bf16>(ma, mb, e, grid, st);
  }
  // swap: the 128-row operand is the weight matrix
  if (w8) SONIC_CUDA_TRY(make_map_i8(&ma, g.W, g.K, g.N, g.ldw, BM));
  else SONIC_CUDA_TRY(make_map(&ma, g.W, g.K, g.N, g.ldw, 1, 0, BM));
  SONIC_CUDA_TRY(make_map(&mb, g.A, g.K, g.M, g.lda, g.batch, g.a_bstride, bn));
  if (g.batch != 1) return cudaErrorInvalidValue;
  const int tiles = cdiv(g.M, bn) * cdiv(g.N, BM);
  const int num_kb = g.K / BK;
  int splits = 128 / tiles;                                     // measured optimum: ~100-150 CTAs streaming (bench_gemm.py)
  if (splits > num_kb / 4) splits = num_kb / 4;
  if (splits > g.K / (8 * g.M)) splits = g.K / (8 * g.M);      // keep the fp32 partial traffic well below the weight bytes
  if (splits < 1) splits = 1;
  if (const char* fs = getenv("SONIC_SPLITS")) { const int v = atoi(fs); if (v >= 1 && v <= num_kb) splits = v; }   // tuning override
  if (!g.splitk_ws || !g.splitk_counters || tiles > kMaxTiles) splits = 1;
  while (splits > 1 && (size_t)tiles * splits * bn * BM * 4 > g.splitk_ws_bytes) --splits;
  e.splits = splits;
  e.ws = g.splitk_ws;
  e.counters = g.splitk_counters;
  dim3 grid(cdiv(g.M, bn), cdiv(g.N, BM), splits);
#define SWAP_CASE(BN_)                                                                                         \
  case BN_:                                                                                                    \
    if (w8) return launch_one<BN_, true, bf16, true>(ma, mb, e, grid, st, g.pdl != 0);                         \
    return g.out_f32 ? launch_one<BN_, true, float>(ma, mb, e, grid, st, g.pdl != 0) : launch_one<BN_, true, bf16>(ma, mb, e, grid, st, g.pdl != 0);
  switch (bn) {
    SWAP_CASE(16)
    SWAP_CASE(32)
    SWAP_CASE(64)
  }
#undef SWAP_CASE
  return cudaErrorInvalidValue;
}

}  // namespace sonic
