// Causal GQA attention of the decoder PREFILL on tcgen05 (sm_100a): 16 query heads / 4 KV heads, head_dim 128, variable
// prompt lengths packed row-wise (LlamaAttention, transformers/models/llama/modeling_llama.py:225-289 via
// sdpa_attention.py:40-104 with is_causal=True).
//
// One CTA = one (segment, query head, 128-query tile).  Same pipeline as attention_tc.cu with head_dim 128:
//   Q / K tiles are two 64-wide SWIZZLE_128B atoms; S = Q.K^T is 8 k-steps; O_t = P.V is issued as two N=64 halves
//   (V atoms used as MN-major operands) into TMEM columns [128,192) and [192,256).  Q comes from the rotated fused QKV
//   rows, K and V straight from the KV cache [segment][kv head][position][128] that rope_dec_kv has just filled.
// Warp roles (192 threads): w0 TMA producer, w1 MMA issuer + TMEM owner, w2..5 softmax (TMEM lane quadrant = warp % 4).
#include "common.cuh"
#include "kernels.h"
#include "gemm_tc.h"
#include "tc_ptx.cuh"

namespace sonic {

static constexpr int PQ = 128, PK = 128, PD = 128;
static constexpr int kAtom = 128 * 64 * 2;                   // 16 KB: [128 rows x 64 cols] bf16
static constexpr int kTile = 2 * kAtom;                      // 32 KB: [128 x 128]
static constexpr int kKStages = 2;
static constexpr int kPfSmem = kTile * (1 + kKStages + 1 + 1) + 1024 + 256;     // Q, K x2, V, P  (~161 KB, one CTA per SM)
static constexpr int kPfTmemCols = 256;

struct PrefillAttnArgs {
  bf16* out; long long out_stride;          // out[(tok_off[seg] + q) * out_stride + h*128 + d]
  const int* tok_off;                       // [B+1] first packed row of each segment
  int q_col0;                               // column of query head 0 in the fused QKV row
  int kv_heads, group, max_ctx;             // 4, 4 (query heads per kv head), rows per (segment, kv head) in the cache
  float scale_log2;
};

__global__ void __launch_bounds__(192, 1)
attention_prefill_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                            const __grid_constant__ CUtensorMap tmV, PrefillAttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (base - raw);
  const uint32_t sQ = base, sK = sQ + kTile, sV = sK + kKStages * kTile, sP = sV + kTile;
  const uint32_t bars = sP + kTile;
  // barriers: 0 q_full | 1,2 k_full | 3,4 k_empty | 5 v_full | 7 v_empty | 9 s_full | 10 p_full | 11 o_full
  auto bar = [&](int i) { return bars + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + (bars - base) + 8 * 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * PQ, h = blockIdx.y, seg = blockIdx.z;
  const int row0 = a.tok_off[seg];
  const int S = a.tok_off[seg + 1] - row0;                   // prompt length == keys in the cache
  if (q0 >= S) return;                                       // CTA-uniform
  const int kvh = h / a.group;
  const int kv_row0 = (seg * a.kv_heads + kvh) * a.max_ctx;
  const int k_hi = min(S, q0 + PQ);                          // causal: no query of this tile sees keys >= k_hi
  const int n_kt = (k_hi + PK - 1) / PK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    for (int i = 0; i < 12; ++i) mbar_init(bar(i), i == 10 ? 128 : 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kPfTmemCols>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(bar(0), kTile);
      tma_load_2d(sQ, &tmQ, bar(0), a.q_col0 + h * PD, row0 + q0);
      tma_load_2d(sQ + kAtom, &tmQ, bar(0), a.q_col0 + h * PD + 64, row0 + q0);
      for (int j = 0; j < n_kt; ++j) {
        const int s = j & 1;
        mbar_wait(bar(3 + s), (uint32_t)((j >> 1) & 1) ^ 1u);
        mbar_expect_tx(bar(1 + s), kTile);
        tma_load_2d(sK + s * kTile, &tmK, bar(1 + s), 0, kv_row0 + j * PK);
        tma_load_2d(sK + s * kTile + kAtom, &tmK, bar(1 + s), 64, kv_row0 + j * PK);
        mbar_wait(bar(7), (uint32_t)(j & 1) ^ 1u);
        mbar_expect_tx(bar(5), kTile);
        tma_load_2d(sV, &tmV, bar(5), 0, kv_row0 + j * PK);
        tma_load_2d(sV + kAtom, &tmV, bar(5), 64, kv_row0 + j * PK);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc_s = make_idesc_bf16_major(PQ, PK, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16_major(PQ, 64, 0, 1);       // B = V atom, MN-major
      auto issue_s = [&](int j) {
        const int s = j & 1;
        mbar_wait(bar(1 + s), (uint32_t)((j >> 1) & 1));
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < PD / 16; ++k) {
          const uint64_t dq = make_sw128_desc(sQ + (k >> 2) * kAtom + (k & 3) * 32);
          const uint64_t dk = make_sw128_desc(sK + s * kTile + (k >> 2) * kAtom + (k & 3) * 32);
          tc_mma_bf16(tS, dq, dk, idesc_s, k != 0 ? 1u : 0u);
        }
        tc_commit(bar(3 + s));
        tc_commit(bar(9));
      };
      mbar_wait(bar(0), 0);
      issue_s(0);
      for (int j = 0; j < n_kt; ++j) {
        mbar_wait(bar(10), (uint32_t)(j & 1));
        mbar_wait(bar(5), (uint32_t)(j & 1));
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < PK / 16; ++k) {
          const uint64_t dp = make_sw128_desc(sP + (k >> 2) * kAtom + (k & 3) * 32);
#pragma unroll
          for (int dh = 0; dh < 2; ++dh) {
            const uint64_t dv = make_sw128_mn_desc(sV + dh * kAtom + k * 16 * 128, 16);
            tc_mma_bf16(tO + dh * 64, dp, dv, idesc_o, k != 0 ? 1u : 0u);
          }
        }
        tc_commit(bar(7));
        tc_commit(bar(11));
        if (j + 1 < n_kt) issue_s(j + 1);
      }
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    float o[PD];
#pragma unroll
    for (int d = 0; d < PD; ++d) o[d] = 0.f;
    float m = -INFINITY, l = 0.f;
    uint8_t* pP = sgen + (sP - base);
    const int qi = q0 + r;                                      // query position inside the segment
    for (int j = 0; j < n_kt; ++j) {
      mbar_wait(bar(9), (uint32_t)(j & 1));
      tc_fence_after();
      const int vis = min(qi, S - 1) - j * PK + 1;              // keys [0, vis) of this tile are visible to this query
      float tmax = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < PK / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tS + lane_off + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, (c * 32 + i < vis) ? __uint_as_float(v[i]) : -INFINITY);
      }
      const float m_new = fmaxf(m, tmax);
      // rows that see nothing in this tile (only possible for padding rows q >= S, never for the first tile of a real
      // row) keep m = -inf: guard the arithmetic so no NaN is produced
      const float corr = (m_new == -INFINITY) ? 1.f : exp2f((m - m_new) * a.scale_log2);
      const float mb = (m_new == -INFINITY) ? 0.f : m_new * a.scale_log2;
      float psum = 0.f;
#pragma unroll 1
      for (int c = 0; c < PK / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tS + lane_off + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = (c * 32 + i < vis) ? exp2f(fmaf(__uint_as_float(v[i]), a.scale_log2, -mb)) : 0.f;
          const float p1 = (c * 32 + i + 1 < vis) ? exp2f(fmaf(__uint_as_float(v[i + 1]), a.scale_log2, -mb)) : 0.f;
          __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
          psum += __low2float(pb) + __high2float(pb);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&pb);
        }
        const int atom = c >> 1;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (c & 1) * 4 + q;
          *reinterpret_cast<uint4*>(pP + atom * kAtom + r * 128 + ((chunk ^ (r & 7)) << 4)) =
              make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      l = l * corr + psum;
      m = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar(10));
      mbar_wait(bar(11), (uint32_t)(j & 1));
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < PD / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tO + lane_off + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], corr, __uint_as_float(v[i]));
      }
      tc_fence_before();
    }
    if (qi < S) {
      const float inv = 1.0f / l;
      bf16* dst = a.out + (size_t)(row0 + qi) * a.out_stride + h * PD;
#pragma unroll
      for (int c = 0; c < PD / 8; ++c) {
        uint4 val;
        __nv_bfloat162 p0 = __floats2bfloat162_rn(o[8 * c] * inv, o[8 * c + 1] * inv);
        __nv_bfloat162 p1 = __floats2bfloat162_rn(o[8 * c + 2] * inv, o[8 * c + 3] * inv);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(o[8 * c + 4] * inv, o[8 * c + 5] * inv);
        __nv_bfloat162 p3 = __floats2bfloat162_rn(o[8 * c + 6] * inv, o[8 * c + 7] * inv);
        val.x = *reinterpret_cast<uint32_t*>(&p0); val.y = *reinterpret_cast<uint32_t*>(&p1);
        val.z = *reinterpret_cast<uint32_t*>(&p2); val.w = *reinterpret_cast<uint32_t*>(&p3);
        reinterpret_cast<uint4*>(dst)[c] = val;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kPfTmemCols>(tmem_base);
}

cudaError_t attention_prefill_tc_configure() {
  return cudaFuncSetAttribute(attention_prefill_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPfSmem);
}

// qkv: rotated fused rows [total_rows, row_width] (q heads at q_col0 + h*128); kcache / vcache: this layer's
// [max_batch][kv_heads][max_ctx][128]; out: [total_rows, out_stride]
cudaError_t launch_attention_prefill_tc(const bf16* qkv, int row_width, int total_rows, int q_col0, const bf16* kcache, const bf16* vcache,
                                        int max_batch, int kv_heads, int heads, int max_ctx, const int* tok_off, int batch, int max_q,
                                        bf16* out, int out_stride, float scale, cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  SONIC_CUDA_TRY(gemm_tc_init());
  CUtensorMap tq, tk, tv;
  SONIC_CUDA_TRY(make_tensor_map_2d(&tq, qkv, row_width, total_rows, row_width, 64, PQ));
  const long long kv_rows = (long long)max_batch * kv_heads * max_ctx;
  SONIC_CUDA_TRY(make_tensor_map_2d(&tk, kcache, PD, kv_rows, PD, 64, PK));
  SONIC_CUDA_TRY(make_tensor_map_2d(&tv, vcache, PD, kv_rows, PD, 64, PK));
  PrefillAttnArgs a;
  a.out = out; a.out_stride = out_stride; a.tok_off = tok_off; a.q_col0 = q_col0; a.kv_heads = kv_heads; a.group = heads / kv_heads;
  a.max_ctx = max_ctx; a.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid(cdiv(max_q, PQ), heads, batch);
  attention_prefill_tc_kernel<<<grid, 192, kPfSmem, st>>>(tq, tk, tv, a);
  return cudaGetLastError();
}

}  // namespace sonic
