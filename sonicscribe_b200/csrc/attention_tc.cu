// Non-causal multi-head attention for the audio encoder on tcgen05 (sm_100a): head_dim 64, T = 1500 keys, no mask
// (GlmAsrAttention, transformers/models/glmasr/modeling_glmasr.py:175-225 via sdpa_attention.py:40-104).
//
// One CTA = one (segment, head, PAIR of 128-query tiles A / B); one CTA per SM.  Per 128-key tile and query tile:
//   S = Q.K^T   tcgen05.mma 128x128x64 (Q, K tiles K-major, TMA 128B swizzle)            -> TMEM cols [0,128) (A) / [128,256) (B)
//   softmax     8 warps per query tile, TWO threads per query row (64 keys each): tcgen05.ld S, row max exchanged between the
//               two halves through shared memory, running max / partial sum in registers (ex2.approx), P written as bf16 into
//               shared memory in the K-major SW128 operand layout
//   O_t = P.V   tcgen05.mma 128x64x128, V tile used as an MN-major SW128 B operand (no transpose)  -> TMEM cols [256,320) / [320,384)
//   O = O*corr + O_t accumulated in registers (exact online softmax, fp32), 32 of the 64 output columns per thread
// The exponentials (MUFU: 16 per clock per SM) and the MMA issue rate (>= 80 clk per tcgen05.mma whatever N: 12 MMAs per tile
// pair and key tile) bound the kernel about equally, so the two must overlap.  Two independent CTAs per SM (the round-1 design)
// fell into lock-step — both in their softmax, then both in their MMAs: 4400 clk per pair of tiles, the SUM of the two bounds.
// Here the two softmax warpgroups of one CTA take turns (named barriers): while A's 8 warps own the MUFU pipe, the tensor pipe
// runs P_B.V and the next S_B, and vice versa; K / V tiles are loaded once for both query tiles.
// Warp roles (576 threads): w0..7 softmax of tile A, w8..15 softmax of tile B (TMEM lane quadrant = warp % 4, key /
// output-column half = (warp % 8) / 4), w16 TMA producer, w17 MMA issuer + TMEM owner.
#include <type_traits>
#include "common.cuh"
#include "kernels.h"
#include "gemm_tc.h"
#include "tc_ptx.cuh"

namespace sonic {

static constexpr int AQ = 128, AK = 128, AD = 64;
static constexpr int kTileBytes = AQ * AD * 2;          // 16 KB: a [128 x 64] bf16 tile
static constexpr int kKvStages = 2;                     // K and V tiles are double-buffered (each is used by both query tiles)
static constexpr int kPBytes = AQ * AK * 2;             // 32 KB: P as two K-major SW128 atoms of 64 keys
static constexpr int kXchgBytes = 2 * 2 * 2 * AQ * 4;   // row-max exchange: [query tile][tile parity][half][row]
static constexpr int kAttnSmem = kTileBytes * (2 + 2 * kKvStages) + 2 * kPBytes + 1024 + 256 + kXchgBytes;   // ~165 KB -> 1 CTA / SM
static constexpr int kAttnTmemCols = 512;
static constexpr int kAttnThreads = 576;
// The warp scheduler favours the highest warp id among ready warps: the two single-thread roles get the highest ids, so that an
// MMA or TMA issue never queues behind sixteen issue-bound softmax warps
static constexpr int kTmaWarp = 16, kMmaWarp = 17;

struct AttnTcArgs {
  bf16* out; long long out_stride;    // out[(seg*T + q) * out_stride + h*64 + d]
  int T;                              // queries == keys per segment
  int q_col, k_col, v_col;            // column offsets of head 0 inside the fused QKV row
  float scale_log2;                   // softmax scale * log2(e)
  int dbg;                            // timing experiments only (SONIC_ATTN_DBG): 1 skip the P.V MMAs, 2 skip the S MMAs, 4 skip the exponential pass
  int turns;                          // 0: the two softmax warpgroups run freely; 1: they take turns for the exponential pass; 2: for the whole softmax
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// The softmax warps are issue-bound (ncu: issue-active 67 %, ALU 50 %, XU 41 %, tensor pipe 20 %), so the per-element
// arithmetic uses Blackwell's packed fp32 pairs and the three-input max: half the FMA / ADD / MNMX instructions.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm, AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (base - raw);
  const uint32_t sQ = base, sK = sQ + 2 * kTileBytes, sV = sK + kKvStages * kTileBytes, sP = sV + kKvStages * kTileBytes;
  const uint32_t bars = sP + 2 * kPBytes;
  // barriers: 0 q_full | 1,2 k_full | 3,4 k_empty | 5,6 v_full | 7,8 v_empty | 9,10 s_full[A,B] | 11,12 p_full[A,B] | 13,14 o_full[A,B]
  auto bar = [&](int i) { return bars + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + (bars - base) + 8 * 16);
  float* sXall = reinterpret_cast<float*>(sgen + (bars - base) + 256);      // [tile][parity][half][row]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * AQ, h = blockIdx.y, seg = blockIdx.z;
  const int row0 = seg * a.T;
  const int n_kt = (a.T + AK - 1) / AK;

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tm);
    for (int i = 0; i < 15; ++i) mbar_init(bar(i), (i == 11 || i == 12) ? 256 : 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<kAttnTmemCols>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kTmaWarp) {
    if (elect_one_sync()) {
      mbar_expect_tx(bar(0), 2 * kTileBytes);
      tma_load_2d(sQ, &tm, bar(0), a.q_col + h * AD, row0 + q0);
      tma_load_2d(sQ + kTileBytes, &tm, bar(0), a.q_col + h * AD, row0 + q0 + AQ);
      for (int j = 0; j < n_kt; ++j) {
        const int s = j & 1;
        const uint32_t ph = (uint32_t)((j >> 1) & 1);
        mbar_wait(bar(3 + s), ph ^ 1u);
        mbar_expect_tx(bar(1 + s), kTileBytes);
        tma_load_2d(sK + s * kTileBytes, &tm, bar(1 + s), a.k_col + h * AD, row0 + j * AK);
        mbar_wait(bar(7 + s), ph ^ 1u);
        mbar_expect_tx(bar(5 + s), kTileBytes);
        tma_load_2d(sV + s * kTileBytes, &tm, bar(5 + s), a.v_col + h * AD, row0 + j * AK);
      }
    }
  } else if (warp == kMmaWarp) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc_s = make_idesc_bf16_major(AQ, AK, 0, 0);     // S: A = Q (K-major), B = K (K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16_major(AQ, AD, 0, 1);     // O: A = P (K-major), B = V (MN-major)
      auto issue_s = [&](int j, int t) {                                    // S_t(j) = Q_t . K_j^T
        const int s = j & 1;
        const uint64_t dq = make_sw128_desc(sQ + t * kTileBytes), dk = make_sw128_desc(sK + s * kTileBytes);
        if (!(a.dbg & 2))
#pragma unroll
        for (int k = 0; k < AD / 16; ++k) tc_mma_bf16(tmem_base + t * 128, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k != 0 ? 1u : 0u);
        tc_commit(bar(9 + t));                                              // S_t ready
      };
      auto issue_pv = [&](int j, int t) {                                   // O_t = P_t(j) . V_j
        const int s = j & 1;
        if (!(a.dbg & 1))
#pragma unroll
        for (int k = 0; k < AK / 16; ++k) {
          const uint64_t dp = make_sw128_desc(sP + t * kPBytes + (k >> 2) * (kPBytes / 2) + (k & 3) * 32);
          const uint64_t dv = make_sw128_mn_desc(sV + s * kTileBytes + k * 16 * 128, 16);
          tc_mma_bf16(tmem_base + 256 + t * 64, dp, dv, idesc_o, k != 0 ? 1u : 0u);
        }
        tc_commit(bar(13 + t));                                             // O_t ready (also: P_t buffer free)
      };
      mbar_wait(bar(0), 0);
      mbar_wait(bar(1), 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(0, 1);
      tc_commit(bar(3));                                                    // K stage 0 free once both S MMAs retire
      for (int j = 0; j < n_kt; ++j) {
        const uint32_t par = (uint32_t)(j & 1), kvpar = (uint32_t)((j >> 1) & 1);
        // query tile A: its P_j is ready while tile B's softmax runs
        mbar_wait(bar(11), par);                                            // P_A(j) in smem, S_A(j) and O_A(j-1) drained
        mbar_wait(bar(5 + (j & 1)), kvpar);                                 // V_j landed
        tc_fence_after();
        issue_pv(j, 0);
        if (j + 1 < n_kt) {
          mbar_wait(bar(1 + ((j + 1) & 1)), (uint32_t)(((j + 1) >> 1) & 1)); // K_{j+1} landed
          tc_fence_after();
          issue_s(j + 1, 0);
        }
        // query tile B
        mbar_wait(bar(12), par);
        tc_fence_after();
        issue_pv(j, 1);
        tc_commit(bar(7 + (j & 1)));                                        // V stage free
        if (j + 1 < n_kt) {
          issue_s(j + 1, 1);
          tc_commit(bar(3 + ((j + 1) & 1)));                                // K stage free
        }
      }
    }
  } else {
    const int t = warp >> 3;                         // query tile of this warpgroup: 0 = A, 1 = B
    const int wl = warp & 7;                         // warp inside the warpgroup
    const int quad = warp & 3;                       // TMEM lanes [32*quad, 32*quad+32)  (= warp % 4)
    const int half = wl >> 2;                        // keys [64*half, +64) of the tile, output columns [32*half, +32)
    const int r = quad * 32 + lane;                  // query row inside the tile
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t tS = tmem_base + t * 128, tO = tmem_base + 256 + t * 64;
    float* sX = sXall + t * 4 * AQ;
    const int pair_bar = 1 + t * 4 + quad;           // named barrier of the two warps that share this lane quadrant
    const int turn_mine = 9 + t, turn_other = 9 + (t ^ 1);   // ping-pong: a warpgroup computes its exponentials only on its turn
    float o[AD / 2];
#pragma unroll
    for (int d = 0; d < AD / 2; ++d) o[d] = 0.f;
    float m = -INFINITY, ref = -INFINITY, l = 0.f;   // m: running row max; ref: the max P, l and o are currently scaled by; l: this thread's keys only
    uint8_t* pP = sgen + (sP - base) + t * kPBytes + half * (kPBytes / 2);           // this half's 64-key atom
    const int turns = a.turns;
    if (turns && t == 1) asm volatile("bar.arrive %0, 512;" ::"r"(9) : "memory");    // tile A goes first
    // TMEM reads run at 64 B per clock per SM (measured: the two-pass softmax of round 1 — every S tile read twice plus the O tile —
    // took 4400 clk per pair of tiles whatever the overlap).  So S is read ONCE: the exponentials of key tile j are taken
    // relative to the running maximum of the tiles before it (softmax is shift-invariant; P, l and o just carry the factor
    // 2^(scale*(max_j - ref)) until the next tile rescales them), and the row maximum of tile j is collected in the same pass.
    // Tile 0 has no predecessor and keeps the max pass.  If a row's maximum jumps by more than 2^64 over the reference (never
    // with LayerNorm'ed inputs) the tile is redone relative to its own maximum, so nothing can overflow.
    float psum = 0.f, tmax = -INFINITY;
    // FULL / WANT_MAX are compile-time: with a run-time mask test inside the loop ptxas predicates the three mask instructions
    // of every element instead of branching, and the predicated-off instructions still take issue slots — 640 instead of
    // ~230 instructions per thread and key tile, which made the kernel issue-bound (ncu: 10 k warp instructions per tile pair)
    auto exp_pass = [&](float refn, int kv_left, auto full_c, auto want_max_c) {
      constexpr bool FULL = decltype(full_c)::value, WANT_MAX = decltype(want_max_c)::value;
      const float mb = refn * a.scale_log2;
      float2 psum2 = make_float2(0.f, 0.f);
      const float2 sc2 = make_float2(a.scale_log2, a.scale_log2), nmb2 = make_float2(-mb, -mb);
      // p = exp2(s*scale - ref*scale) -> bf16 -> shared memory (K-major SW128: 16 B chunk c8 of row r at (c8 ^ (r & 7)))
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tS + lane_off + half * 64 + c * 32, v);
        tmem_ld_wait();
        if (WANT_MAX) {
          if (FULL) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) tmax = max3(tmax, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, (c * 32 + i < kv_left) ? __uint_as_float(v[i]) : -INFINITY);
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 e = fma2(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nmb2);
          float p0 = ex2_approx(e.x), p1 = ex2_approx(e.y);
          if (!FULL) {
            if (c * 32 + i >= kv_left) p0 = 0.f;
            if (c * 32 + i + 1 >= kv_left) p1 = 0.f;
          }
          // the denominator is the fp32 sum, the numerator uses P cast to bf16: the reference's softmax (fp32) then cast
          psum2 = add2(psum2, make_float2(p0, p1));
          __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&pb);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = c * 4 + q;                                 // 16 B chunk index inside the 128 B row
          uint4 val = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          *reinterpret_cast<uint4*>(pP + r * 128 + ((chunk ^ (r & 7)) << 4)) = val;
        }
      }
      psum = psum2.x + psum2.y;
    };
    auto exp_pass_rt = [&](float refn, bool full, int kv_left, bool want_max) {
      if (full) {
        if (want_max) exp_pass(refn, kv_left, std::true_type{}, std::true_type{});
        else exp_pass(refn, kv_left, std::true_type{}, std::false_type{});
      } else {
        if (want_max) exp_pass(refn, kv_left, std::false_type{}, std::true_type{});
        else exp_pass(refn, kv_left, std::false_type{}, std::false_type{});
      }
    };
    for (int j = 0; j < n_kt; ++j) {
      mbar_wait(bar(9 + t), (uint32_t)(j & 1));
      tc_fence_after();
      if (turns == 2) asm volatile("bar.sync %0, 512;" ::"r"(turn_mine) : "memory");  // my turn on the MUFU pipe
      const int kv_left = a.T - j * AK - half * 64;  // keys >= kv_left (in this half's numbering) are outside the segment
      const bool full = kv_left >= 64;
      float* xb = sX + (j & 1) * 2 * AQ;
      tmax = -INFINITY;
      if (j == 0) {
        // first tile: row max over this thread's 64 keys, then over both halves
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(tS + lane_off + half * 64 + c * 32, v);
          tmem_ld_wait();
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) tmax = max3(tmax, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, (c * 32 + i < kv_left) ? __uint_as_float(v[i]) : -INFINITY);
          }
        }
        xb[half * AQ + r] = tmax;
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");      // the two warps that share this lane quadrant
        m = fmaxf(tmax, xb[(half ^ 1) * AQ + r]);
      }
      float refn = m;                                                  // reference of this tile: the maximum seen so far
      if (turns == 1) asm volatile("bar.sync %0, 512;" ::"r"(turn_mine) : "memory");  // my turn on the MUFU pipe
      if (!(a.dbg & 4)) exp_pass_rt(refn, full, kv_left, j != 0);
      if (turns == 1 && !(t == 1 && j == n_kt - 1)) asm volatile("bar.arrive %0, 512;" ::"r"(turn_other) : "memory");
      if (j != 0) {
        xb[half * AQ + r] = tmax;
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        const float m_new = fmaxf(m, fmaxf(tmax, xb[(half ^ 1) * AQ + r]));
        // same rows and same (m_new, refn) in both warps of the pair, so both take the same decision
        if (__any_sync(0xffffffffu, (m_new - refn) * a.scale_log2 > 64.f)) {
          refn = m_new;
          exp_pass_rt(refn, full, kv_left, false);
        }
        m = m_new;
      }
      const float corr = ex2_approx((ref - refn) * a.scale_log2);      // ref = -inf on the first tile -> 0
      l = l * corr + psum;
      ref = refn;
      fence_proxy_async_smem();          // generic-proxy writes of P -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(bar(11 + t));
      if (turns == 2 && !(t == 1 && j == n_kt - 1)) asm volatile("bar.arrive %0, 512;" ::"r"(turn_other) : "memory");
      // O accumulate (this thread's 32 output columns)
      mbar_wait(bar(13 + t), (uint32_t)(j & 1));
      tc_fence_after();
      {
        uint32_t v[32];
        tmem_ld32(tO + lane_off + half * 32, v);
        tmem_ld_wait();
        const float2 corr2 = make_float2(corr, corr);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 r2 = fma2(make_float2(o[i], o[i + 1]), corr2, make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
          o[i] = r2.x; o[i + 1] = r2.y;
        }
      }
      tc_fence_before();
    }
    // denominator of the row = both halves' partial sums (same reference, so they add directly)
    float* xb = sX + (n_kt & 1) * 2 * AQ;
    xb[half * AQ + r] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    const float l_row = (half == 0) ? l + xb[AQ + r] : xb[r] + l;     // same order in both threads
    const int q = q0 + t * AQ + r;
    if (q < a.T) {
      const float inv = 1.0f / l_row;
      bf16* dst = a.out + (size_t)(row0 + q) * a.out_stride + h * AD + half * 32;
#pragma unroll
      for (int c = 0; c < AD / 16; ++c) {
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
  if (warp == kMmaWarp) tmem_dealloc<kAttnTmemCols>(tmem_base);
}

cudaError_t attention_tc_configure() {
  return cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
}

// qkv: fused [segments*T, row_width] bf16 (q | k | v column blocks), out: [segments*T, out_stride]
cudaError_t launch_attention_tc(const bf16* qkv, int row_width, int q_col, int k_col, int v_col, bf16* out, int out_stride, int segments,
                                int T, int heads, float scale, cudaStream_t st) {
  if (segments <= 0) return cudaSuccess;
  SONIC_CUDA_TRY(gemm_tc_init());
  CUtensorMap tm;
  SONIC_CUDA_TRY(make_tensor_map_2d(&tm, qkv, row_width, (long long)segments * T, row_width, AD, AQ));
  AttnTcArgs a;
  a.out = out; a.out_stride = out_stride; a.T = T; a.q_col = q_col; a.k_col = k_col; a.v_col = v_col;
  a.scale_log2 = scale * 1.4426950408889634f;
  { static const int turns = [] { const char* e = getenv("SONIC_ATTN_TURNS"); return e ? atoi(e) : 1; }(); a.turns = turns; }
  { static const int dbg = [] { const char* e = getenv("SONIC_ATTN_DBG"); return e ? atoi(e) : 0; }(); a.dbg = dbg; }
  dim3 grid(cdiv(T, 2 * AQ), heads, segments);
  attention_tc_kernel<<<grid, kAttnThreads, kAttnSmem, st>>>(tm, a);
  return cudaGetLastError();
}

}  // namespace sonic
