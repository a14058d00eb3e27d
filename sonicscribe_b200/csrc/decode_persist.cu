// Persistent greedy-decode step for the Llama decoder (bf16 weights/activations, fp32 accumulation): ONE cooperative
// kernel per generated token runs all 28 layers, the lm_head and the greedy pick, with grid-wide barriers between
// dependent phases instead of ~200 separate launches.
//
// Why: a decode step moves 2.9 GB of weights (0.45 ms at HBM speed) through ~200 tiny dependent operations; as separate
// kernels each one costs ~9-20 us of launch/setup/drain latency (measured).  Here every SM keeps 16 warps
// resident for the whole step, one kernel instantiation per batch class:
//   * 33..64 tokens, bf16: every GEMM phase is TMA (6-stage ring) -> tcgen05.mma (128 weight rows x 64 tokens, two TMEM
//     accumulators) -> tcgen05.ld epilogue; split-K partials are summed by the consumer phase; the next phase's weight tiles
//     are put in flight before the grid barrier (gemm_phase_tc).
//   * <= 32 tokens, and int8 weights: each warp streams 16 weight rows x a K-slice straight from global memory into
//     mma.sync fragments (16 B loads, 8 in flight per lane), the 16 warps of a CTA split K and reduce through shared memory
//     (gemm_phase).
//   * attention: a team of warps per (segment, kv head); K/V chunks by TMA tile loads into two alternating buffers,
//     QK^T / online softmax / PV on mma.sync (attention_phase_mma); with few segments the keys are split over CTAs with a
//     last-arriver merge (attention_phase).
// Every reduction has a fixed order (deterministic); residual / RMSNorm / RoPE / SwiGLU / KV append / argmax are phases or
// epilogues of the same kernel and never exist as separate passes.
//
// Replaces, for one new token per segment: LlamaDecoderLayer x28 + final norm + lm_head + argmax/EOS bookkeeping
// (transformers/models/llama/modeling_llama.py:53-499, transformers/generation/utils.py:2743-2809).
#include <cooperative_groups.h>
#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace sonic {

static constexpr int PH = 2048, PQKV = 3072, PI = 6144, PV_ = 59264, PHD = 128, PKVH = 4, PG = 4;
static constexpr int kPThreads = 512, kPWarps = 16;

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
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)::"memory");     // "memory": stays on its side of barriers
  return t;
}

// grid-wide barrier: monotonically increasing arrival counter (zeroed by the host before the launch).  The gpu-scope fences
// order every thread's global writes before the arrival and invalidate L1 after the wait, so plain loads see fresh data.
template <bool PROXY = false>
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch, int flags = 0, unsigned long long* dbg = nullptr) {
  // PROXY: the next phase may read this phase's global writes, or overwrite shared memory it used, through the async proxy
  // (TMA); order the generic-proxy accesses of every thread before that
  const bool wdbg = dbg && (int)blockIdx.x == (flags >> 8) && (threadIdx.x & 31) == 0;
  if (wdbg) dbg[960 + (threadIdx.x >> 5)] = gtimer();
  if (PROXY && !(flags & 1)) asm volatile("fence.proxy.async.global;" ::: "memory");
  if (wdbg) dbg[980 + (threadIdx.x >> 5)] = gtimer();
  __syncthreads();
  if (wdbg) dbg[1000 + (threadIdx.x >> 5)] = gtimer();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    // release-arrive without waiting for the atomic's return value (one L2 round trip less than fence + atomicAdd), relaxed
    // polling, one acquire fence at the end (it also invalidates L1 for the whole CTA)
    unsigned v;
    if (dbg) {                                             // debug: arrival with a return value, so that its completion can be stamped
      dbg[480 + blockIdx.x] = gtimer();
      asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(v) : "l"(counter) : "memory");
      dbg[640 + blockIdx.x] = gtimer() + (v & 0u);
    } else {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    }
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < epoch);
    if (dbg) dbg[800 + blockIdx.x] = gtimer();
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  __syncthreads();
  if (PROXY && !(flags & 2)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- GEMM phase: part[s][tok][n] = sum_{k in slice s} W[n][k] * X[tok][k] ------------------------------------------------
// item = (K-slice s, 16-row block rb); the warps of one CTA take consecutive row blocks of the same slice so that the
// activation slice they all read stays in L1.
// ---- GEMM phase ---------------------------------------------------------------------------------------------------------
// item = (16-row block rb, 2048-wide K section gs).  The 16 warps of the CTA split the section (128 k each: one group of
// 8 weight loads per lane), exchange their fp32 fragments through shared memory and sum them in warp order (deterministic),
// so a finished 16 x tokens tile leaves the CTA and the epilogue can be fused: fp32 store (per section), residual add into
// the bf16 stream, or SwiGLU of interleaved (gate, up) rows.  The next item's weight loads are issued before the exchange.
enum { EPI_F32 = 0, EPI_RESID = 1, EPI_SWIGLU = 2 };

static constexpr int kXRowBytes = 2048 * 2 + 64;      // staged activation row: 2048 bf16 + 64 B pad (conflict-free 16 B reads)
static constexpr int kXStageTok = 32;                 // tokens staged per pass
static constexpr int kXStageBytes = kXStageTok * kXRowBytes;

// NT: 8-token tiles per pass; EPI: fused epilogue; STAGE: copy the pass's activation rows for the current 2048-wide K
// section into shared memory once and feed every item from there (at >= 9 tokens the per-item activation reads from L2
// would otherwise exceed the weight bytes); tok0: first token of the pass.
// W8: the weight matrix is int8 [N][K] with a per-row fp32 scale (weight-only quantisation): each lane
// loads 16 int8 (16 B) per row and 64-k chunk, expands them to bf16 pairs in registers and the row scale is applied to the
// summed tile in the epilogue.
__device__ __forceinline__ void i8x4_to_bf16x2(uint32_t w, uint32_t& lo, uint32_t& hi) {
  const float f0 = (float)(int)(int8_t)(w & 0xff), f1 = (float)(int)(int8_t)((w >> 8) & 0xff);
  const float f2 = (float)(int)(int8_t)((w >> 16) & 0xff), f3 = (float)(int)(int8_t)(w >> 24);
  __nv_bfloat162 a = __floats2bfloat162_rn(f0, f1), b = __floats2bfloat162_rn(f2, f3);
  lo = *reinterpret_cast<uint32_t*>(&a);
  hi = *reinterpret_cast<uint32_t*>(&b);
}

template <int NT, int EPI, bool STAGE, bool W8>
__device__ __forceinline__ void gemm_phase(const void* __restrict__ Wv, const float* __restrict__ wscale, int N, int K, const bf16* X, int B,
                                           int Bpad, int tok0, float* __restrict__ out32, bf16* xres, bf16* act, uint8_t* smem) {
  // int8 rows are half as long in bytes: an int8 item covers two 16-row blocks per warp so that every lane still has eight
  // 16 B weight loads in flight (and the activation fragments are shared by both blocks)
  constexpr int RBW = (W8 && NT <= 4) ? 2 : 1;
  const bf16* W = reinterpret_cast<const bf16*>(Wv);
  const int8_t* W8p = reinterpret_cast<const int8_t*>(Wv);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int nib = N / (16 * RBW), gsplit = K >> 11, n_items = nib * gsplit;
  uint8_t* sX = smem;
  float* sR = reinterpret_cast<float*>(smem + (STAGE ? kXStageBytes : 0));
  int item = blockIdx.x;
  int staged_gs = -1;
  uint4 wa[4], wb[4];
  auto load_item = [&](int it) {
    const int gs = it / nib, ib = it - gs * nib;
    if (W8) {                                                    // 2 x 16 B per row: k = 64 v + 16 t + [0, 16)
#pragma unroll
      for (int rbl = 0; rbl < RBW; ++rbl) {
        const int8_t* q0 = W8p + (size_t)((ib * RBW + rbl) * 16 + g) * K + (size_t)gs * 2048 + warp * 128 + 16 * t;
        const int8_t* q1 = q0 + (size_t)8 * K;
#pragma unroll
        for (int v = 0; v < 2; ++v) { wa[rbl * 2 + v] = ldg_stream(q0 + 64 * v); wb[rbl * 2 + v] = ldg_stream(q1 + 64 * v); }
      }
    } else {
      const bf16* w0 = W + (size_t)(ib * 16 + g) * K + (size_t)gs * 2048 + warp * 128 + 8 * t;
      const bf16* w1 = w0 + (size_t)8 * K;
#pragma unroll
      for (int u = 0; u < 4; ++u) { wa[u] = ldg_stream(w0 + 32 * u); wb[u] = ldg_stream(w1 + 32 * u); }
    }
  };
  if (item < n_items) load_item(item);
  for (; item < n_items; item += gridDim.x) {
    const int gs = item / nib, ib = item - gs * nib;
    if (STAGE && gs != staged_gs) {                              // CTA-uniform
      __syncthreads();
      for (int i = threadIdx.x; i < NT * 8 * 256; i += kPThreads) {
        const int row = i >> 8, c = i & 255;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (tok0 + row < B) v = __ldcg(reinterpret_cast<const uint4*>(X + (size_t)(tok0 + row) * K + (size_t)gs * 2048 + c * 8));
        *reinterpret_cast<uint4*>(sX + (size_t)row * kXRowBytes + c * 16) = v;
      }
      __syncthreads();
      staged_gs = gs;
    }
    float acc[RBW][NT][4];
#pragma unroll
    for (int rbl = 0; rbl < RBW; ++rbl)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) { acc[rbl][nt][0] = acc[rbl][nt][1] = acc[rbl][nt][2] = acc[rbl][nt][3] = 0.f; }
    if (W8) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        uint32_t a_lo[RBW][4], a_hi[RBW][4], b_lo[RBW][4], b_hi[RBW][4];
#pragma unroll
        for (int rbl = 0; rbl < RBW; ++rbl) {
          const uint4 qa = wa[rbl * 2 + v], qb = wb[rbl * 2 + v];
          const uint32_t ra[4] = {qa.x, qa.y, qa.z, qa.w}, rb[4] = {qb.x, qb.y, qb.z, qb.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) { i8x4_to_bf16x2(ra[q], a_lo[rbl][q], a_hi[rbl][q]); i8x4_to_bf16x2(rb[q], b_lo[rbl][q], b_hi[rbl][q]); }
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          uint4 x0, x1;
          if (STAGE) {
            const uint8_t* px = sX + (size_t)(nt * 8 + g) * kXRowBytes + (warp * 128 + 64 * v + 16 * t) * 2;
            x0 = *reinterpret_cast<const uint4*>(px);
            x1 = *reinterpret_cast<const uint4*>(px + 16);
          } else {
            // token rows >= B are not read (their accumulator columns are never stored): at one segment the activation
            // footprint in L1 is one row instead of eight
            const bf16* px = X + (size_t)(tok0 + nt * 8 + g) * K + (size_t)gs * 2048 + warp * 128 + 64 * v + 16 * t;
            x0 = x1 = make_uint4(0, 0, 0, 0);
            if (tok0 + nt * 8 + g < B) { x0 = *reinterpret_cast<const uint4*>(px); x1 = *reinterpret_cast<const uint4*>(px + 8); }
          }
#pragma unroll
          for (int rbl = 0; rbl < RBW; ++rbl) {
            mma16816(acc[rbl][nt], a_lo[rbl][0], b_lo[rbl][0], a_hi[rbl][0], b_hi[rbl][0], x0.x, x0.y);
            mma16816(acc[rbl][nt], a_lo[rbl][1], b_lo[rbl][1], a_hi[rbl][1], b_hi[rbl][1], x0.z, x0.w);
            mma16816(acc[rbl][nt], a_lo[rbl][2], b_lo[rbl][2], a_hi[rbl][2], b_hi[rbl][2], x1.x, x1.y);
            mma16816(acc[rbl][nt], a_lo[rbl][3], b_lo[rbl][3], a_hi[rbl][3], b_hi[rbl][3], x1.z, x1.w);
          }
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          uint4 xv;
          if (STAGE) {
            xv = *reinterpret_cast<const uint4*>(sX + (size_t)(nt * 8 + g) * kXRowBytes + (warp * 128 + 32 * u + 8 * t) * 2);
          } else {
            xv = make_uint4(0, 0, 0, 0);
            if (tok0 + nt * 8 + g < B) xv = *reinterpret_cast<const uint4*>(X + (size_t)(tok0 + nt * 8 + g) * K + (size_t)gs * 2048 + warp * 128 + 32 * u + 8 * t);
          }
          mma16816(acc[0][nt], wa[u].x, wb[u].x, wa[u].y, wb[u].y, xv.x, xv.y);
          mma16816(acc[0][nt], wa[u].z, wb[u].z, wa[u].w, wb[u].w, xv.z, xv.w);
        }
      }
    }
    if (item + (int)gridDim.x < n_items) load_item(item + gridDim.x);     // in flight during the exchange below
#pragma unroll
    for (int rbl = 0; rbl < RBW; ++rbl)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        *reinterpret_cast<float4*>(sR + ((size_t)((warp * RBW + rbl) * NT + nt) * 32 + lane) * 4) =
            make_float4(acc[rbl][nt][0], acc[rbl][nt][1], acc[rbl][nt][2], acc[rbl][nt][3]);
    __syncthreads();
    for (int e = threadIdx.x; e < RBW * NT * 128; e += kPThreads) {
      const int rbl = e / (NT * 128), e2 = e - rbl * (NT * 128);
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kPWarps; ++w) v += sR[(size_t)(w * RBW + rbl) * NT * 128 + e2];
      const int c = e2 & 3, ml = (e2 >> 2) & 31, nt = e2 >> 7;
      const int row = (ib * RBW + rbl) * 16 + (ml >> 2) + 8 * (c >> 1), tok = tok0 + nt * 8 + 2 * (ml & 3) + (c & 1);
      if (W8) v *= __ldg(wscale + row);
      if (EPI == EPI_SWIGLU) {
        const float up = __shfl_down_sync(0xffffffffu, v, 16);            // row + 1 of the same token sits 16 threads up
        if (((ml >> 2) & 1) == 0 && tok < B) act[(size_t)tok * (N >> 1) + (row >> 1)] = __float2bfloat16_rn(silu(v) * up);
      } else if (tok < B) {
        if (EPI == EPI_F32) out32[((size_t)gs * Bpad + tok) * N + row] = v;
        else {
          bf16* px = xres + (size_t)tok * N + row;
          *px = __float2bfloat16_rn(__bfloat162float(*px) + v);
        }
      }
    }
    __syncthreads();
  }
}

// NT (8-token tiles) is a template parameter of the whole kernel: one instantiation per batch class keeps a single copy of
// the GEMM phase per epilogue in the kernel, which matters for register allocation (measured)
template <int EPI, bool W8, int NT>
__device__ __forceinline__ void gemm_dispatch(const void* W, const float* wscale, int N, int K, const bf16* X, int B, int Bpad, float* out32,
                                              bf16* xres, bf16* act, uint8_t* smem) {
  // activations staged in shared memory for 9..32 tokens; <= 8 read them directly (tiny), > 32 too (two staged 32-token
  // passes measured slower)
  gemm_phase<NT, EPI, (NT == 2 || NT == 4), W8>(W, wscale, N, K, X, B, Bpad, 0, out32, xres, act, smem);
}

// ---- GEMM phase on tcgen05 (batch class 33..64, bf16 weights) ---------------------------------------------------------------
// D[128 weight rows x 64 tokens] (fp32, TMEM) += W tile . X tile^T per 64-wide k block; both operands arrive by TMA
// (SWIZZLE_128B) in a 6-stage shared-memory ring, so the weight stream does not pass through registers and each weight byte
// meets only half a byte of activation traffic (the mma.sync phase re-reads 4 B of activations per weight byte at 64 tokens).
// item = (128-row tile, K split); `splits` is chosen per matrix so that items ~ grid.  Warp 0 lane 0 produces, warp 1 lane 0
// issues the MMAs, warps 4..11 drain the two alternating TMEM accumulators (lane quadrant = warp % 4, 32 tokens each).
// The ring / accumulator barriers live for the whole kernel; their phase is derived from running counters that every thread
// advances identically.
static constexpr int kTcStages = 6;                   // 8 measured no faster
static constexpr int kTcStageA = 128 * 64 * 2, kTcStageB = 128 * 64 * 2;
static constexpr int kTcRingBytes = kTcStages * (kTcStageA + kTcStageB) + 1024;     // + alignment slack
static constexpr int kTcEpiBytes = 8 * 16 * 32 * 4;   // epilogue transpose buffers (TcCtx::epi): [16 tokens][32 features] fp32 per warp
// int8 weights (W8): the 96 KB weight area of the ring holds six raw int8 stages (128 rows x 64 B = 8 KB, no swizzle) followed by
// three bf16 slots (16 KB, SWIZZLE_128B) that warps 2..15 fill from the raw stages; the MMA reads the slots
static constexpr int kTcRawA = 128 * 64, kTcConvSlots = 3, kTcConvWarps = 14;
static constexpr int kTcConvBase = kTcStages * kTcRawA;
struct TcCtx {
  uint32_t ringA, ringB, bars, tmem;
  uint32_t kb_count, item_count;
  uint32_t pre;                  // k blocks of the coming phase whose weight tile is already in flight (tc_prefetch_weights)
  uint32_t ntok;                 // token rows of the activation tiles / MMA N of this launch: 16, 32, 64, 128 or 256 (>= live batch)
  uint32_t stages, bstride;      // ring depth and activation-stage stride: 6 x (16 KB + 16 KB), or 4 x (16 KB + 32 KB) at 256 tokens
  uint64_t wpolicy;              // L2 policy of the weight / KV streams (evict_first), or 0: default
  uint32_t pre_depth;            // how many stages tc_prefetch_weights may fill before a grid barrier (<= kTcStages)
  float* epi;                    // 8 epilogue warps x [32 tokens][32 features] fp32 transpose buffers
  uint8_t* ring_gen;             // generic-address view of ringA (int8 converter warps)
};
__device__ __forceinline__ uint32_t tc_full(const TcCtx& c, uint32_t s) { return c.bars + 8u * s; }
__device__ __forceinline__ uint32_t tc_empty(const TcCtx& c, uint32_t s) { return c.bars + 8u * (kTcStages + s); }
__device__ __forceinline__ uint32_t tc_acc_full(const TcCtx& c, uint32_t a) { return c.bars + 8u * (2 * kTcStages + a); }
__device__ __forceinline__ uint32_t tc_acc_empty(const TcCtx& c, uint32_t a) { return c.bars + 8u * (2 * kTcStages + 2 + a); }
__device__ __forceinline__ uint32_t tc_conv_full(const TcCtx& c, uint32_t s) { return c.bars + 8u * (2 * kTcStages + 4 + s); }
__device__ __forceinline__ uint32_t tc_conv_empty(const TcCtx& c, uint32_t s) { return c.bars + 8u * (2 * kTcStages + 4 + kTcConvSlots + s); }
// (An fp16 expansion — one PRMT per two values — would halve the ALU work of the converters, but tcgen05 kind::f16 rejects an fp16 A
// operand next to a bf16 B operand: illegal instruction, measured.)
// four int8 of a word (already XORed with 0x80808080: bytes are value + 128) -> two bf16 pairs, exactly: the byte is placed in the
// low mantissa bits of 2^23 and (2^23 + 128) is subtracted — one PRMT + one FADD per value instead of an I2F
__device__ __forceinline__ void u8x4_to_bf16x2(uint32_t w, uint32_t& lo, uint32_t& hi) {
  const float f0 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540)) - 8388736.0f;
  const float f1 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7541)) - 8388736.0f;
  const float f2 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7542)) - 8388736.0f;
  const float f3 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7543)) - 8388736.0f;
  __nv_bfloat162 a = __floats2bfloat162_rn(f0, f1), b = __floats2bfloat162_rn(f2, f3);
  lo = *reinterpret_cast<uint32_t*>(&a);
  hi = *reinterpret_cast<uint32_t*>(&b);
}

// Weights do not depend on the previous phase: thread 0 (the producer) puts the weight tiles of this CTA's first k blocks of
// the NEXT tcgen05 phase in flight (arming the stage for weight + activation bytes) before the grid barrier / while a
// non-GEMM phase runs; the producer of that phase then only adds the activation tiles.  Every thread computes `pre`.
template <bool W8 = false>
__device__ __forceinline__ void tc_prefetch_weights(const CUtensorMap* wmap, int N, int K, int splits, TcCtx& tc, int tile_rows = 128) {
  constexpr int kRowB = W8 ? 64 : 128, kStA = W8 ? kTcRawA : kTcStageA;
  const int tiles = (N + tile_rows - 1) / tile_rows, nkb = K >> 6, n_items = tiles * splits;
  uint32_t n = 0;
  for (int item = blockIdx.x; item < n_items && n < tc.pre_depth; item += gridDim.x) {
    const int tile = item / splits, ks = item - tile * splits;
    const int kb0 = (ks * nkb) / splits, kb1 = ((ks + 1) * nkb) / splits;
    for (int kb = kb0; kb < kb1 && n < tc.pre_depth; ++kb, ++n) {
      if (threadIdx.x < 32 && elect_one_sync()) {
        const uint32_t cnt = tc.kb_count + n, st = cnt % tc.stages, par = (cnt / tc.stages) & 1u;
        mbar_wait(tc_empty(tc, st), par ^ 1u);
        mbar_expect_tx(tc_full(tc, st), tile_rows * kRowB + tc.ntok * 128);
        if (tc.wpolicy) tma_load_2d_hint(tc.ringA + st * kStA, wmap, tc_full(tc, st), kb * 64, tile * tile_rows, tc.wpolicy);
        else tma_load_2d(tc.ringA + st * kStA, wmap, tc_full(tc, st), kb * 64, tile * tile_rows);
      }
    }
  }
  tc.pre = n;
}

template <int EPI, bool W8 = false>
__device__ __forceinline__ void gemm_phase_tc(const CUtensorMap* wmap, const CUtensorMap* xmap, int N, int K, int splits, int B, int Bpad,
                                              float* __restrict__ out32, bf16* __restrict__ act, TcCtx& tc, int tile_rows = 128,
                                              unsigned long long* dbg = nullptr, int dbgf = 0, const float* __restrict__ wscale = nullptr) {
  constexpr int kRowB = W8 ? 64 : 128, kStA = W8 ? kTcRawA : kTcStageA;
  // tile_rows: weight rows per item (the box height of `wmap`).  128 fills the MMA tile; a smaller value (gate/up: 84, so that
  // 147 whole-K items cover the 12288 rows, one per CTA) leaves the remaining rows of the 128-row shared-memory tile stale:
  // they only feed accumulator rows the epilogue never reads.  Rows past N are zero-filled by TMA.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = (N + tile_rows - 1) / tile_rows, nkb = K >> 6, n_items = tiles * splits;
  if (warp == 0) {
    if (elect_one_sync()) {
      if (dbg) dbg[0] = gtimer();
      uint32_t cnt = tc.kb_count, idx = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tile = item / splits, ks = item - tile * splits;
        const int kb0 = (ks * nkb) / splits, kb1 = ((ks + 1) * nkb) / splits;
        for (int kb = kb0; kb < kb1; ++kb, ++cnt, ++idx) {
          const uint32_t st = cnt % tc.stages, par = (cnt / tc.stages) & 1u;
          if (idx >= tc.pre) {
            mbar_wait(tc_empty(tc, st), par ^ 1u);
            mbar_expect_tx(tc_full(tc, st), tile_rows * kRowB + tc.ntok * 128);
            if (tc.wpolicy) tma_load_2d_hint(tc.ringA + st * kStA, wmap, tc_full(tc, st), kb * 64, tile * tile_rows, tc.wpolicy);
            else tma_load_2d(tc.ringA + st * kStA, wmap, tc_full(tc, st), kb * 64, tile * tile_rows);
          }
          tma_load_2d(tc.ringB + st * tc.bstride, xmap, tc_full(tc, st), kb * 64, 0);
        }
        if (dbg) dbg[1 + (item >= (int)gridDim.x)] = gtimer();       // producer: all loads of item 0 / 1 issued
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc_bf16(128, (int)tc.ntok);
      uint32_t cnt = tc.kb_count, ic = tc.item_count;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ic) {
        const int tile = item / splits, ks = item - tile * splits;
        const int kb0 = (ks * nkb) / splits, kb1 = ((ks + 1) * nkb) / splits;
        const uint32_t acc = ic & 1u, apar = (ic >> 1) & 1u;
        mbar_wait(tc_acc_empty(tc, acc), apar ^ 1u);
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb, ++cnt) {
          const uint32_t st = cnt % tc.stages, par = (cnt / tc.stages) & 1u;
          const uint32_t cs = cnt % kTcConvSlots, cpar = (cnt / kTcConvSlots) & 1u;
          if (W8) mbar_wait(tc_conv_full(tc, cs), cpar);              // the converters waited for the raw stage (and its X tile)
          else mbar_wait(tc_full(tc, st), par);
          tc_fence_after();
          if (dbg && kb == kb0) dbg[3 + (item >= (int)gridDim.x)] = gtimer();   // first stage of item 0 / 1 landed
          const uint64_t da = make_sw128_desc(W8 ? tc.ringA + kTcConvBase + cs * kTcStageA : tc.ringA + st * kTcStageA);
          const uint64_t db = make_sw128_desc(tc.ringB + st * tc.bstride);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(tc.tmem + acc * kPersistTcMaxTokens, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > kb0 || k != 0) ? 1u : 0u);
          tc_commit(tc_empty(tc, st));
          if (W8) tc_commit(tc_conv_empty(tc, cs));
        }
        tc_commit(tc_acc_full(tc, acc));
        if (dbg) dbg[5 + (item >= (int)gridDim.x)] = gtimer();       // all MMAs of item 0 / 1 issued
      }
    }
  } else if (warp >= 2 && (W8 || (warp >= 4 && warp < 12))) {
    // Warps 4..11: epilogue — lane = weight row (TMEM lane), registers = 32 tokens.  The outputs are token-major, so the 32 x 32
    // block is transposed through a per-warp shared-memory buffer and leaves as 16 B vectors: lane (tr, fc) handles token rows
    // tr + 4 i (i = 0..7), features fc..fc+3 — 8 vector stores per lane instead of 32 scalar ones (measured: the scalar
    // version spent 1.3 us issuing the fp32 stores and 3.8 us in the SwiGLU stores of one item).
    // int8 weights: warps 2..15 (the epilogue warps included — every int8 phase has at most one item per CTA, so they are idle
    // until its accumulator completes) first expand the item's raw int8 stages to bf16: thread piece = (row r, 16 values j) of
    // a stage, 16 B in, two 16 B chunks out at their SWIZZLE_128B places (chunk c of row r lives at c ^ (r & 7)); rows past the
    // tile are left alone.
    const bool is_epi = warp >= 4 && warp < 12;
    const int q = warp & 3, half = (warp - 4) >> 2;
    float* eb = tc.epi + (warp - 4) * 512;
    const int tr = lane >> 3, fc = (lane & 7) * 4;
    uint32_t ic = tc.item_count, ccnt = tc.kb_count;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ic) {
      const int tile = item / splits, ks = item - tile * splits;
      if (W8) {
        const int kb0 = (ks * nkb) / splits, kb1 = ((ks + 1) * nkb) / splits;
        const int n_pieces = min(tile_rows, N - tile * tile_rows) * 4;
        // this thread's (at most two) pieces: offsets are the same for every k block
        const int i0 = (warp - 2) * 32 + lane, i1 = i0 + kTcConvWarps * 32;
        const bool v0 = i0 < n_pieces, v1 = i1 < n_pieces;
        const int r0 = i0 >> 2, j0 = i0 & 3, r1 = i1 >> 2, j1 = i1 & 3;
        const int s0 = i0 * 16, s1 = i1 * 16;
        const int d00 = r0 * 128 + (((2 * j0) ^ (r0 & 7)) << 4), d01 = r0 * 128 + (((2 * j0 + 1) ^ (r0 & 7)) << 4);
        const int d10 = r1 * 128 + (((2 * j1) ^ (r1 & 7)) << 4), d11 = r1 * 128 + (((2 * j1 + 1) ^ (r1 & 7)) << 4);
        uint32_t st = ccnt % tc.stages, par = (ccnt / tc.stages) & 1u;
        uint32_t cs = ccnt % kTcConvSlots, cpar = (ccnt / kTcConvSlots) & 1u;
        for (int kb = kb0; kb < kb1; ++kb, ++ccnt) {
          unsigned long long* dc = (dbg && warp == 2 && lane == 0 && kb - kb0 < 10) ? dbg + 100 + 3 * (kb - kb0) : nullptr;
          mbar_wait(tc_full(tc, st), par);
          if (dc) dc[0] = gtimer();
          mbar_wait(tc_conv_empty(tc, cs), cpar ^ 1u);
          if (dc) dc[1] = gtimer();
          const uint8_t* raw = tc.ring_gen + st * kTcRawA;
          uint8_t* dst = tc.ring_gen + kTcConvBase + cs * kTcStageA;
          uint4 w0 = make_uint4(0, 0, 0, 0), w1 = make_uint4(0, 0, 0, 0);
          if (v0) w0 = *reinterpret_cast<const uint4*>(raw + s0);
          if (v1) w1 = *reinterpret_cast<const uint4*>(raw + s1);
          if (v0) {
            uint32_t o[8];
            u8x4_to_bf16x2(w0.x ^ 0x80808080u, o[0], o[1]); u8x4_to_bf16x2(w0.y ^ 0x80808080u, o[2], o[3]);
            u8x4_to_bf16x2(w0.z ^ 0x80808080u, o[4], o[5]); u8x4_to_bf16x2(w0.w ^ 0x80808080u, o[6], o[7]);
            *reinterpret_cast<uint4*>(dst + d00) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(dst + d01) = make_uint4(o[4], o[5], o[6], o[7]);
          }
          if (v1) {
            uint32_t o[8];
            u8x4_to_bf16x2(w1.x ^ 0x80808080u, o[0], o[1]); u8x4_to_bf16x2(w1.y ^ 0x80808080u, o[2], o[3]);
            u8x4_to_bf16x2(w1.z ^ 0x80808080u, o[4], o[5]); u8x4_to_bf16x2(w1.w ^ 0x80808080u, o[6], o[7]);
            *reinterpret_cast<uint4*>(dst + d10) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(dst + d11) = make_uint4(o[4], o[5], o[6], o[7]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(tc_conv_full(tc, cs));
          if (dc) dc[2] = gtimer();
          if (++st == tc.stages) { st = 0; par ^= 1u; }
          if (++cs == (uint32_t)kTcConvSlots) { cs = 0; cpar ^= 1u; }
        }
      }
      if (!is_epi) continue;
      const uint32_t acc = ic & 1u, apar = (ic >> 1) & 1u;
      mbar_wait(tc_acc_full(tc, acc), apar);
      tc_fence_after();
      unsigned long long* dw = (dbg && warp == 4 && lane == 0) ? dbg + 8 + 8 * (item >= (int)gridDim.x) : nullptr;
      if (dw) dw[0] = gtimer();                                     // accumulator complete
      float rs = 1.0f;                                             // int8 weights: dequantisation scale of this lane's weight row
      if (W8) { const int lr = q * 32 + lane, row = tile * tile_rows + lr; rs = (lr < tile_rows && row < N) ? __ldg(wscale + row) : 0.f; }
      const int feat = tile * tile_rows + q * 32 + fc;              // first of this lane's 4 features
      const bool fvalid = (q * 32 + fc < tile_rows) && (feat < N);  // tile_rows and N are multiples of 4
      // this warp drains the 32-token column blocks `half` and `half + 2` of the accumulator (those below the tile width)
      const int n_blk = max(0, ((int)tc.ntok / 32 + (tc.ntok == 16 ? 1 : 0) - half + 1) / 2);    // blocks half, half + 2, ... below the tile width
      if (n_blk == 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tc_acc_empty(tc, acc));
      }
      for (int bi = 0; bi < n_blk; ++bi) {
        const int blk = half + 2 * bi;
        uint32_t v[32];
        tmem_ld32(tc.tmem + ((uint32_t)(q * 32) << 16) + acc * kPersistTcMaxTokens + blk * 32, v);
        tmem_ld_wait();
        if (bi == n_blk - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tc_acc_empty(tc, acc));       // the accumulator is in registers: the next item may start
        }
#pragma unroll
        for (int p = 0; p < 2; ++p) {                               // two passes of 16 tokens through the [16][32] transpose buffer
#pragma unroll
          for (int j = 0; j < 16; ++j) eb[j * 32 + lane] = W8 ? __uint_as_float(v[16 * p + j]) * rs : __uint_as_float(v[16 * p + j]);
          __syncwarp();
          if (EPI == EPI_SWIGLU) {
            // rows are interleaved (gate, up): features (fc, fc+1) and (fc+2, fc+3) are two (gate, up) pairs
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int tl = i * 4 + tr, tok = blk * 32 + 16 * p + tl;
              // accumulator rows past the tile come from stale shared memory (possibly NaN / denormal bit patterns, for which
              // the division takes its slow path: measured 5 us on the CTAs whose ring never held a 128-row tile): no
              // arithmetic on them
              if (fvalid && tok < B) {
                const float4 x = *reinterpret_cast<const float4*>(eb + tl * 32 + fc);
                __nv_bfloat162 ob = __floats2bfloat162_rn(silu(x.x) * x.y, silu(x.z) * x.w);
                *reinterpret_cast<uint32_t*>(act + (size_t)tok * (N >> 1) + (feat >> 1)) = *reinterpret_cast<uint32_t*>(&ob);
              }
            }
          } else {
            float* o = out32 + (size_t)ks * Bpad * N + feat;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int tl = i * 4 + tr, tok = blk * 32 + 16 * p + tl;
              if (fvalid && tok < B) *reinterpret_cast<float4*>(o + (size_t)tok * N) = *reinterpret_cast<const float4*>(eb + tl * 32 + fc);
            }
          }
          __syncwarp();                                             // the buffer is rewritten by the next pass / item
        }
      }
      if (dw) dw[5] = gtimer();
    }
  }
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int tile = item / splits, ks = item - tile * splits;
    tc.kb_count += (uint32_t)(((ks + 1) * nkb) / splits - (ks * nkb) / splits);
    tc.item_count += 1u;
  }
  tc.pre = 0;
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
// ---- row phase: x[b] += sum of KS fp32 sections (KS = 0: x is already final); u[b] = rmsnorm(x[b]) * gamma ------------------
template <int KS>
__device__ __forceinline__ void residual_norm_phase(const float* part, int B, int Bpad, bf16* x, bf16* u,
                                                    const float* __restrict__ gamma, float eps, float* red) {
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const int c0 = threadIdx.x * 4;
    float v[4];
    const uint2 xr = __ldcg(reinterpret_cast<const uint2*>(x + (size_t)b * PH + c0));
    v[0] = __uint_as_float(xr.x << 16); v[1] = __uint_as_float(xr.x & 0xffff0000u);
    v[2] = __uint_as_float(xr.y << 16); v[3] = __uint_as_float(xr.y & 0xffff0000u);
    float ss = 0.f;
    if (KS > 0) {
      float add[4] = {0.f, 0.f, 0.f, 0.f};
      {
        float4 pv[KS > 0 ? KS : 1];                                  // all split-K partials in flight, summed in split order
#pragma unroll
        for (int s2 = 0; s2 < KS; ++s2) pv[s2] = __ldcg(reinterpret_cast<const float4*>(part + (size_t)s2 * Bpad * PH + (size_t)b * PH + c0));
#pragma unroll
        for (int s2 = 0; s2 < KS; ++s2) { add[0] += pv[s2].x; add[1] += pv[s2].y; add[2] += pv[s2].z; add[3] += pv[s2].w; }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = bf16r(v[i] + add[i]);
      __nv_bfloat162 o0 = __floats2bfloat162_rn(v[0], v[1]), o1 = __floats2bfloat162_rn(v[2], v[3]);
      *reinterpret_cast<uint2*>(x + (size_t)b * PH + c0) = make_uint2(*reinterpret_cast<uint32_t*>(&o0), *reinterpret_cast<uint32_t*>(&o1));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) ss += v[i] * v[i];
    const float rstd = rsqrtf(block_sum(ss, red) / PH + eps);
    const float4 gm = *reinterpret_cast<const float4*>(gamma + c0);
    __nv_bfloat162 u0 = __floats2bfloat162_rn(gm.x * bf16r(v[0] * rstd), gm.y * bf16r(v[1] * rstd));
    __nv_bfloat162 u1 = __floats2bfloat162_rn(gm.z * bf16r(v[2] * rstd), gm.w * bf16r(v[3] * rstd));
    *reinterpret_cast<uint2*>(u + (size_t)b * PH + c0) = make_uint2(*reinterpret_cast<uint32_t*>(&u0), *reinterpret_cast<uint32_t*>(&u1));
    __syncthreads();
  }
}

// ---- attention phase: item = (segment, kv head, chunk of 128 keys).  Every item finishes q (4 heads) from the fp32 qkv
// section and rotates it; the item owning the newest position also rotates k and appends k, v to the cache.  Each item writes
// an (m, l, o) partial; the item arriving last at the (segment, kv head) counter merges the partials in chunk order.
template <int KSQ>
__device__ __forceinline__ void attention_phase(const DecodePersistArgs& a, const DecLayerDev& L, uint8_t* smem, const int* s_ctx,
                                                unsigned long long* dbg = nullptr) {
  int n_dbg = 0;
#define ASTAMP() do { if (dbg && threadIdx.x == 0 && n_dbg < 16) dbg[n_dbg++] = gtimer(); } while (0)
  ASTAMP();
  uint8_t* sK = smem;                                          // AKEYS * kAKRow
  bf16* sV = reinterpret_cast<bf16*>(smem + AKEYS * kAKRow);   // AKEYS * 128
  float* sQ = reinterpret_cast<float*>(smem + AKEYS * kAKRow + AKEYS * PHD * 2);   // [4][128]
  float* sP = sQ + PG * PHD;                                   // [4][AKEYS]
  float* sKV = sP + PG * AKEYS;                                // new k (128) | new v (128)
  float* sMax = sKV + 2 * PHD;                                 // [4 heads][4 key groups]
  float* sSum = sMax + 16;                                     // [4 heads][4 key groups]
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = a.attn_chunks;
  const int n_items = a.B * PKVH * C;
  const int CKEYS = a.attn_chunk_keys;                          // 64 or 128 keys per item (<= AKEYS)
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int chunk = item % C, grp = item / C;                  // grp = seg * 4 + kvh
    const int seg = grp / PKVH, kvh = grp - seg * PKVH;
    const int pos = s_ctx[seg], kv_len = pos + 1;
    const int n_chunks = (kv_len + CKEYS - 1) / CKEYS;
    float* ws = a.attn_ws + ((size_t)grp * C + chunk) * PG * (PHD + 2);
    if (chunk < n_chunks) {
      bf16* kc = L.kc + ((size_t)seg * PKVH + kvh) * a.max_ctx * PHD;
      bf16* vc = L.vc + ((size_t)seg * PKVH + kvh) * a.max_ctx * PHD;
      const bool owner = (chunk == pos / CKEYS);
      const int k0 = chunk * CKEYS;
      const int nk = min(CKEYS, kv_len - k0);
      // the chunk's cached K / V rows do not depend on this step's q, k, v: their loads are issued first and land while the
      // partial sums / RoPE below wait for theirs
      uint4 kreg[AKEYS * (PHD / 8) / kPThreads], vreg[AKEYS * (PHD / 8) / kPThreads];
#pragma unroll
      for (int it = 0; it < AKEYS * (PHD / 8) / kPThreads; ++it) {
        const int i = tid + it * kPThreads, r = i / (PHD / 8), c8 = i - r * (PHD / 8);
        kreg[it] = make_uint4(0, 0, 0, 0); vreg[it] = make_uint4(0, 0, 0, 0);
        if (r < nk && k0 + r != pos) {
          kreg[it] = __ldcg(reinterpret_cast<const uint4*>(kc + (size_t)(k0 + r) * PHD + c8 * 8));
          vreg[it] = __ldcg(reinterpret_cast<const uint4*>(vc + (size_t)(k0 + r) * PHD + c8 * 8));
        }
      }
      ASTAMP();                                                    // K / V loads issued
      const int n_pairs = (owner ? PG + 2 : PG) * (PHD / 2);
      for (int i = tid; i < n_pairs; i += kPThreads) {
        const int hh = i / (PHD / 2), j = i - hh * (PHD / 2);     // hh < 4: query head; 4: key; 5: value (pair j, j+64)
        const int col = (hh < PG) ? (kvh * PG + hh) * PHD : (hh == PG ? (16 + kvh) * PHD : (16 + PKVH + kvh) * PHD);
        const float x = bf16r(sum_partials<KSQ>(a.part, (size_t)a.Bpad * PQKV, (size_t)seg * PQKV + col + j));
        const float y = bf16r(sum_partials<KSQ>(a.part, (size_t)a.Bpad * PQKV, (size_t)seg * PQKV + col + j + PHD / 2));
        if (hh <= PG) {
          const float c = bf16r(a.cos_t[(size_t)pos * (PHD / 2) + j]), sn = bf16r(a.sin_t[(size_t)pos * (PHD / 2) + j]);
          const float rx = bf16r(x * c - y * sn), ry = bf16r(y * c + x * sn);
          if (hh < PG) { sQ[hh * PHD + j] = rx * a.scale; sQ[hh * PHD + j + PHD / 2] = ry * a.scale; }
          else { sKV[j] = rx; sKV[j + PHD / 2] = ry; }
        } else { sKV[PHD + j] = x; sKV[PHD + j + PHD / 2] = y; }
      }
      __syncthreads(); ASTAMP();
      if (owner) {
        if (tid < PHD) kc[(size_t)pos * PHD + tid] = __float2bfloat16_rn(sKV[tid]);
        else if (tid < 2 * PHD) vc[(size_t)pos * PHD + tid - PHD] = __float2bfloat16_rn(sKV[tid]);
      }
#pragma unroll
      for (int it = 0; it < AKEYS * (PHD / 8) / kPThreads; ++it) {
        const int i = tid + it * kPThreads, r = i / (PHD / 8), c8 = i - r * (PHD / 8);
        uint4 kk = kreg[it], vv = vreg[it];
        if (r < nk && k0 + r == pos) {                           // the row this CTA appends: take it from shared memory
          uint32_t wk[4], wv[4];
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            __nv_bfloat162 pk = __floats2bfloat162_rn(sKV[c8 * 8 + 2 * e2], sKV[c8 * 8 + 2 * e2 + 1]);
            __nv_bfloat162 pv = __floats2bfloat162_rn(sKV[PHD + c8 * 8 + 2 * e2], sKV[PHD + c8 * 8 + 2 * e2 + 1]);
            wk[e2] = *reinterpret_cast<uint32_t*>(&pk); wv[e2] = *reinterpret_cast<uint32_t*>(&pv);
          }
          kk = make_uint4(wk[0], wk[1], wk[2], wk[3]); vv = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
        *reinterpret_cast<uint4*>(sK + r * kAKRow + c8 * 16) = kk;
        *reinterpret_cast<uint4*>(sV + r * PHD + c8 * 8) = vv;
      }
      __syncthreads(); ASTAMP();
      const int head = warp & 3, kgrp = warp >> 2;               // scores: 4 heads x 4 groups of 32 keys
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
      if (lane == 0) sMax[head * 4 + kgrp] = wmax;
      __syncthreads(); ASTAMP();
      const float cmax = fmaxf(fmaxf(sMax[head * 4], sMax[head * 4 + 1]), fmaxf(sMax[head * 4 + 2], sMax[head * 4 + 3]));
      const float p = (sc == -INFINITY) ? 0.f : expf(sc - cmax);
      sP[head * AKEYS + r] = p;
      const float wsum = warp_sum(p);
      if (lane == 0) sSum[head * 4 + kgrp] = wsum;
      __syncthreads(); ASTAMP();
      {
        const int ph = tid >> 7, d = tid & 127;                  // PV: thread = (head, dim)
        float acc = 0.f;
        const float* pp = sP + ph * AKEYS;
        // rows >= nk hold p = 0 and zero V rows: the loop runs over the whole chunk, 8 independent loads per step
        for (int j0 = 0; j0 < CKEYS; j0 += 8) {
          float pv[8], vv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) { pv[u] = pp[j0 + u]; vv[u] = __bfloat162float(sV[(j0 + u) * PHD + d]); }
#pragma unroll
          for (int u = 0; u < 8; ++u) acc = fmaf(pv[u], vv[u], acc);
        }
        ws[(size_t)ph * (PHD + 2) + d] = acc;
        if (d == 0) {
          ws[(size_t)ph * (PHD + 2) + PHD] = fmaxf(fmaxf(sMax[ph * 4], sMax[ph * 4 + 1]), fmaxf(sMax[ph * 4 + 2], sMax[ph * 4 + 3]));
          ws[(size_t)ph * (PHD + 2) + PHD + 1] = sSum[ph * 4] + sSum[ph * 4 + 1] + sSum[ph * 4 + 2] + sSum[ph * 4 + 3];
        }
      }
    }
    // arrival + last-arriver merge (chunk order => deterministic)
    __threadfence();
    __syncthreads(); ASTAMP();
    if (tid == 0) {
      const int prev = atomicAdd(a.attn_counters + grp, 1);
      s_last = (prev == C - 1) ? 1 : 0;
      if (s_last) a.attn_counters[grp] = 0;
    }
    __syncthreads(); ASTAMP();
    if (s_last) {
      __threadfence();
      const int ph = tid >> 7, d = tid & 127;
      const float* wg = a.attn_ws + (size_t)grp * C * PG * (PHD + 2);
      // four chunks' partials in flight per step; accumulation stays in chunk order
      float M = -INFINITY, Ls = 0.f, acc = 0.f;
      for (int c0 = 0; c0 < n_chunks; c0 += 4) {
        float mv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) mv[u] = (c0 + u < n_chunks) ? __ldcg(wg + ((size_t)(c0 + u) * PG + ph) * (PHD + 2) + PHD) : -INFINITY;
#pragma unroll
        for (int u = 0; u < 4; ++u) M = fmaxf(M, mv[u]);
      }
      for (int c0 = 0; c0 < n_chunks; c0 += 4) {
        float mv[4], lv[4], av[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float* pc = wg + ((size_t)(c0 + u) * PG + ph) * (PHD + 2);
          const bool ok = c0 + u < n_chunks;
          mv[u] = ok ? __ldcg(pc + PHD) : -INFINITY; lv[u] = ok ? __ldcg(pc + PHD + 1) : 0.f; av[u] = ok ? __ldcg(pc + d) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float w = (c0 + u < n_chunks) ? expf(mv[u] - M) : 0.f;
          Ls += lv[u] * w;
          acc += av[u] * w;
        }
      }
      a.attn[(size_t)seg * PH + (size_t)(kvh * PG + ph) * PHD + d] = __float2bfloat16_rn(acc / Ls);
    }
    __syncthreads(); ASTAMP();
  }
}
#undef ASTAMP

// ---- attention phase on mma.sync (used when every CTA has at least one (segment, kv head) group to itself) -----------------
// Per 128-key chunk: S[4 heads(16) x 128 keys] = Q.K^T with warp w owning keys 8w..8w+7 (8 k-steps over head_dim), online
// softmax state per head kept by the lanes that own that head's fragment rows, P (bf16) through shared memory, then
// O[4(16) x 128 dims] += P.V with warp w owning dims 8w..8w+7 (V fragments via ldmatrix.trans).  Rows 4..15 of the 16-row MMA
// tile are zero padding.
static constexpr int kARow = PHD * 2 + 16;               // padded bf16 row (272 B): conflict-free 32-bit fragment loads

__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

// TW warps form a team that owns one (segment, kv head) group at a time: 16 (one team, 128-key chunks) or 8 (two teams per
// CTA, 64-key chunks; used by the 33..64-token class, where there are more groups than CTAs and the per-chunk latency chain,
// not bandwidth, sets the time).  Teams synchronise on their own named barrier.
// K and V chunks arrive by bulk-async row copies (cp.async.bulk, 256 B per key row into 272 B padded rows) into two
// alternating buffers per team, completion on an mbarrier: the next chunk (of this group, or the first chunk of the team's
// next group) is in flight while the current one is multiplied, without passing through registers or the LSU queue
// (register-staged 16 B loads stalled ~1 us per chunk at issue: 64 KB per SM of outstanding LDG is the limit, measured).
template <int TW>
__device__ __forceinline__ void team_sync(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(TW * 32) : "memory");
}
template <int TW>
static constexpr size_t attn_kv_smem() { return (size_t)(kPWarps / TW) * 2 * (8 * TW) * 512; }   // teams x 2 buffers x (K lo|hi, V lo|hi) x 128 B rows
template <int TW>
static constexpr size_t attn_small_smem() { return 2 * 4 * (size_t)kARow + (2 * PHD + 8 * TW) * 4; }       // per team
// byte offset of the 16 B piece `c16` (0..15 over the 256 B of a key row) of row r inside a chunk part laid out by TMA as
// two [rows x 128 B] SWIZZLE_128B halves (dims 0..63 | 64..127), `half_bytes` = rows * 128
__device__ __forceinline__ uint32_t kv_sw(int r, int c16, int half_bytes) {
  return (uint32_t)((c16 >> 3) * half_bytes + r * 128 + ((((c16 & 7) ^ (r & 7))) << 4));
}

template <int KSQ, int TW>
__device__ __forceinline__ void attention_phase_mma(const DecodePersistArgs& a, const DecLayerDev& L, uint8_t* smem_kv, uint8_t* smem_small,
                                                    uint32_t bars, uint32_t& n_chunk, const int* s_ctx) {
  constexpr int CK = 8 * TW;                              // keys per chunk
  constexpr int TT = 32 * TW;                             // threads per team
  constexpr int NTEAM = kPWarps / TW;
  constexpr int NTD = 16 / TW;                            // 8-dim output tiles per warp
  constexpr int kHalfB = CK * 128;                        // one [CK rows x 128 B] swizzled half of K or V
  constexpr int kBuf = 4 * kHalfB;                        // K lo | K hi | V lo | V hi
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int team = warp / TW, wt = warp - team * TW, tt = tid - team * TT;
  uint8_t* kv = smem_kv + (size_t)team * 2 * kBuf;         // two buffers
  uint8_t* small = smem_small + (size_t)team * attn_small_smem<TW>();
  uint8_t* sQ = small;                                    // [4][272 B]   rotated query heads (bf16); MMA rows 4..15 are zero registers
  uint8_t* sP = sQ + 4 * kARow;                           // [4][272 B]   probabilities of the chunk (bf16)
  float* sKV = reinterpret_cast<float*>(sP + 4 * kARow);  // new k (128) | new v (128)
  float* sMax = sKV + 2 * PHD;                            // [4 heads][TW warps]
  float* sSum = sMax + 4 * TW;                            // [4 heads][TW warps]
  const CUtensorMap* kmap = reinterpret_cast<const CUtensorMap*>(a.kv_maps);
  const CUtensorMap* vmap = kmap + 1;
  const int n_items = a.B * PKVH;
  const int item_stride = gridDim.x * NTEAM;
  auto bar_of = [&](uint32_t b) { return bars + 8u * (team * 2 + b); };
  const uint64_t kvpol = l2_policy_evict_first();          // every cached key / value is read once per step
  // one thread asks TMA for the whole chunk: [64 keys x 64 dims] boxes of the K and V caches (one 2-D map per cache over all
  // layers, segments and kv heads; row = key), landing as SWIZZLE_128B halves.  Rows past the context are finite cache
  // contents (masked below); the row appended this step is overwritten from sKV after the chunk has landed.
  auto issue_chunk = [&](int row0, int k0, uint32_t b) {
    if (wt == 0 && elect_one_sync()) {
      mbar_expect_tx(bar_of(b), (uint32_t)kBuf);
      const uint32_t dst = smem_u32(kv + (size_t)b * kBuf);
#pragma unroll
      for (int s2 = 0; s2 < CK / 64; ++s2) {
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          tma_load_2d_hint(dst + h2 * kHalfB + s2 * 8192, kmap, bar_of(b), 64 * h2, row0 + k0 + 64 * s2, kvpol);
          tma_load_2d_hint(dst + 2 * kHalfB + h2 * kHalfB + s2 * 8192, vmap, bar_of(b), 64 * h2, row0 + k0 + 64 * s2, kvpol);
        }
      }
    }
  };
  auto group_of = [&](int item, int& seg, int& kvh, int& pos, int& row0) {
    seg = item / PKVH; kvh = item - seg * PKVH;
    pos = s_ctx[seg];
    row0 = (int)((L.kc - a.kc_base) / PHD) + (seg * PKVH + kvh) * a.max_ctx;
  };
  int item = blockIdx.x * NTEAM + team;
  if (item < n_items) {
    int seg, kvh, pos, row0;
    group_of(item, seg, kvh, pos, row0);
    issue_chunk(row0, 0, n_chunk & 1u);
  }
  for (; item < n_items; item += item_stride) {
    int seg, kvh, pos, row0;
    group_of(item, seg, kvh, pos, row0);
    const int kv_len = pos + 1;
    bf16* kcw = L.kc + ((size_t)seg * PKVH + kvh) * a.max_ctx * PHD;
    bf16* vcw = L.vc + ((size_t)seg * PKVH + kvh) * a.max_ctx * PHD;
    team_sync<TW>(team);                                   // previous group's fragments are consumed
    for (int i = tt; i < (PG + 2) * (PHD / 2); i += TT) {
      const int hh = i / (PHD / 2), j = i - hh * (PHD / 2);     // hh < 4: query head; 4: key; 5: value (pair j, j+64)
      const int col = (hh < PG) ? (kvh * PG + hh) * PHD : (hh == PG ? (16 + kvh) * PHD : (16 + PKVH + kvh) * PHD);
      const float x = bf16r(sum_partials<KSQ>(a.part, (size_t)a.Bpad * PQKV, (size_t)seg * PQKV + col + j));
      const float y = bf16r(sum_partials<KSQ>(a.part, (size_t)a.Bpad * PQKV, (size_t)seg * PQKV + col + j + PHD / 2));
      if (hh <= PG) {
        const float c = bf16r(a.cos_t[(size_t)pos * (PHD / 2) + j]), sn = bf16r(a.sin_t[(size_t)pos * (PHD / 2) + j]);
        const float rx = bf16r(x * c - y * sn), ry = bf16r(y * c + x * sn);
        if (hh < PG) {
          reinterpret_cast<bf16*>(sQ + hh * kARow)[j] = __float2bfloat16_rn(rx);
          reinterpret_cast<bf16*>(sQ + hh * kARow)[j + PHD / 2] = __float2bfloat16_rn(ry);
        } else { sKV[j] = rx; sKV[j + PHD / 2] = ry; }
      } else { sKV[PHD + j] = x; sKV[PHD + j + PHD / 2] = y; }
    }
    team_sync<TW>(team);
    if (tt < PHD) kcw[(size_t)pos * PHD + tt] = __float2bfloat16_rn(sKV[tt]);
    else if (tt < 2 * PHD) vcw[(size_t)pos * PHD + tt - PHD] = __float2bfloat16_rn(sKV[tt]);
    float m_run = -INFINITY, l_part = 0.f;                 // online-softmax state of head g (lanes with g < 4); l_part: this lane's keys only
    float o[NTD][2];                                       // O[head g][dims 8*(NTD*wt + nt) + 2t, +1]
#pragma unroll
    for (int nt = 0; nt < NTD; ++nt) { o[nt][0] = 0.f; o[nt][1] = 0.f; }
    for (int k0 = 0; k0 < kv_len; k0 += CK, ++n_chunk) {
      const int nk = min(CK, kv_len - k0);
      const uint32_t b = n_chunk & 1u;
      uint8_t* sK = kv + (size_t)b * kBuf;
      uint8_t* sV = sK + 2 * kHalfB;
      mbar_wait(bar_of(b), (n_chunk >> 1) & 1u);           // this chunk has landed
      if (pos >= k0 && pos < k0 + CK && tt < 32) {         // the appended row comes from sKV, not from the cache
        const int r = pos - k0, c16 = tt & 15, isv = tt >> 4;
        uint32_t w[4];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          __nv_bfloat162 q2 = __floats2bfloat162_rn(sKV[isv * PHD + c16 * 8 + 2 * e2], sKV[isv * PHD + c16 * 8 + 2 * e2 + 1]);
          w[e2] = *reinterpret_cast<uint32_t*>(&q2);
        }
        *reinterpret_cast<uint4*>((isv ? sV : sK) + kv_sw(r, c16, kHalfB)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      team_sync<TW>(team);                                 // fix-up row visible; the other buffer is fully consumed
      fence_proxy_async_smem();                            // ... by generic-proxy reads, before the async proxy overwrites it
      if (k0 + CK < kv_len) {
        issue_chunk(row0, k0 + CK, b ^ 1u);
      } else if (item + item_stride < n_items) {
        int nseg, nkvh, npos, nrow0;
        group_of(item + item_stride, nseg, nkvh, npos, nrow0);
        issue_chunk(nrow0, 0, b ^ 1u);
      }
      // S tile of this warp: keys 8*wt + {2t, 2t+1} for head g
      float sc[4] = {0.f, 0.f, 0.f, 0.f};
      const int kr = 8 * wt + g;                           // key row of this lane's B fragment
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        uint32_t qa0 = 0u, qa2 = 0u;                       // Q fragment: row g of the 16-row tile, zero for g >= 4
        if (g < PG) {
          qa0 = *reinterpret_cast<const uint32_t*>(sQ + g * kARow + (16 * ks + 2 * t) * 2);
          qa2 = *reinterpret_cast<const uint32_t*>(sQ + g * kARow + (16 * ks + 8 + 2 * t) * 2);
        }
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(sK + kv_sw(kr, 2 * ks, kHalfB) + 4 * t);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(sK + kv_sw(kr, 2 * ks + 1, kHalfB) + 4 * t);
        mma16816(sc, qa0, 0u, qa2, 0u, b0, b1);
      }
      const int key0 = 8 * wt + 2 * t;
      const float s0 = (key0 < nk) ? sc[0] * a.scale : -INFINITY;
      const float s1 = (key0 + 1 < nk) ? sc[1] * a.scale : -INFINITY;
      float wmax = fmaxf(s0, s1);
      wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, 1));
      wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, 2));
      if (g < PG && t == 0) sMax[g * TW + wt] = wmax;
      team_sync<TW>(team);
      float corr = 1.f;
      if (g < PG) {
        float cmax = sMax[g * TW];
#pragma unroll
        for (int w = 1; w < TW; ++w) cmax = fmaxf(cmax, sMax[g * TW + w]);
        const float m_new = fmaxf(m_run, cmax);
        corr = (m_run == -INFINITY) ? 0.f : expf(m_run - m_new);
        m_run = m_new;
        const float p0 = (s0 == -INFINITY) ? 0.f : expf(s0 - m_new);
        const float p1 = (s1 == -INFINITY) ? 0.f : expf(s1 - m_new);
        __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
        *reinterpret_cast<uint32_t*>(sP + g * kARow + key0 * 2) = *reinterpret_cast<uint32_t*>(&pb);
        l_part = l_part * corr + (__low2float(pb) + __high2float(pb));   // the denominator sums what the tensor core multiplies
      }
      team_sync<TW>(team);
      // O tiles of this warp: dims 8*(NTD*wt + nt) + {2t, 2t+1} for head g
      float oc[NTD][4];
#pragma unroll
      for (int nt = 0; nt < NTD; ++nt) { oc[nt][0] = oc[nt][1] = oc[nt][2] = oc[nt][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < CK / 16; ++ks) {
        uint32_t pa0 = 0u, pa2 = 0u;
        if (g < PG) {
          pa0 = *reinterpret_cast<const uint32_t*>(sP + g * kARow + (16 * ks + 2 * t) * 2);
          pa2 = *reinterpret_cast<const uint32_t*>(sP + g * kARow + (16 * ks + 8 + 2 * t) * 2);
        }
#pragma unroll
        for (int nt = 0; nt < NTD; ++nt) {
          uint32_t vb0, vb1;
          ldmatrix_x2_trans(vb0, vb1, sV + kv_sw(16 * ks + (lane & 15), NTD * wt + nt, kHalfB));
          mma16816(oc[nt], pa0, 0u, pa2, 0u, vb0, vb1);
        }
      }
#pragma unroll
      for (int nt = 0; nt < NTD; ++nt) { o[nt][0] = o[nt][0] * corr + oc[nt][0]; o[nt][1] = o[nt][1] * corr + oc[nt][1]; }
    }
    // denominator: this lane's keys -> the 4 lanes of the head row -> the team's warps (fixed order)
    if (g < PG) {
      float ws = l_part;
      ws += __shfl_xor_sync(0x0000ffffu, ws, 1);
      ws += __shfl_xor_sync(0x0000ffffu, ws, 2);
      if (t == 0) sSum[g * TW + wt] = ws;
    }
    team_sync<TW>(team);
    if (g < PG) {
      float ls = 0.f;
#pragma unroll
      for (int w = 0; w < TW; ++w) ls += sSum[g * TW + w];
      const float inv = 1.f / ls;
#pragma unroll
      for (int nt = 0; nt < NTD; ++nt) {
        __nv_bfloat162 ob = __floats2bfloat162_rn(o[nt][0] * inv, o[nt][1] * inv);
        *reinterpret_cast<uint32_t*>(a.attn + (size_t)seg * PH + (size_t)(kvh * PG + g) * PHD + 8 * (NTD * wt + nt) + 2 * t) = *reinterpret_cast<uint32_t*>(&ob);
      }
    }
  }
  __syncthreads();
}

#define STAMP()                                                                      \
  do {                                                                               \
    if (a.timestamps && blockIdx.x == 0 && threadIdx.x == 0) a.timestamps[n_stamp++] = gtimer(); \
  } while (0)

// K splits of the tcgen05 phases (items = tiles x splits ~ one per CTA of a 148-SM grid): qkv 24 tiles x 6, o 16 x 9,
// gate/up 147 tiles of kPersistGuTileRows = 84 rows x 1 (whole K: the SwiGLU epilogue needs whole sums), down 16 x 9, lm_head 570 tiles of kPersistLmTileRows = 104 rows x 1 (four waves of 104 rows instead of four of 128)
// (measured: 84 / 56 / 72-row tiles with 4 / 4 / 5 splits, i.e. 145-148 items and fewer partials, are slower — 50 vs 46.5 us per
// layer at 64 segments: a CTA then issues 2-3 x as many MMAs (>= 80 clk each) behind its last weight bytes)
static constexpr int kTcSplitQkv = 6, kTcSplitO = 9, kTcSplitDown = 9;

template <bool W8, int NT, bool TC>
__global__ void __launch_bounds__(kPThreads, 1) decode_persist_kernel(DecodePersistArgs a) {
  static_assert(!TC || NT == 8, "tcgen05 phases: one instantiation (64-token tiles) per weight type");
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ float red[32];
  __shared__ __align__(8) uint64_t tc_bars[2 * kTcStages + 4 + 2 * kTcConvSlots];
  __shared__ uint32_t tc_tmem_slot;
  __shared__ __align__(8) uint64_t attn_bars[4];                  // [team][buffer] K/V chunk arrival
  // context length of every segment (constant until the pick phase).  The register-streaming classes (<= 64 segments) keep the
  // array at 64 entries: their 194 KB of dynamic shared memory plus the static arrays must stay within the 196 KB carve-out —
  // 512 B more selects the 228 KB one, leaves 28 KB instead of 60 KB of L1 for the activation rows they re-read, and cost
  // 18 % of the int8 batch-1 step (1.29 -> 1.53 ms per token, measured)
  constexpr int kCtxSlots = TC ? kPersistTcMaxTokens : 64;
  __shared__ int s_ctx[kCtxSlots];
  unsigned epoch = 0;
  int n_stamp = 0;
  STAMP();
  const int tid = threadIdx.x;
  const int B = a.B, Bpad = a.Bpad;
  TcCtx tc;
  uint32_t attn_chunks_done = 0;                                   // this team's running chunk count (buffer / barrier phase)
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&attn_bars[i]), 1);
    fence_barrier_init();
  }
  if (!TC) __syncthreads();
  const CUtensorMap* tmaps = reinterpret_cast<const CUtensorMap*>(a.tmaps);
  // activation maps {u, attn, act} x token-tile widths {64, 32, 16, 128, 256}
  const CUtensorMap* xmaps = tmaps + 4 * a.n_layers + 1 + (a.tc_ntok == 64 ? 0 : (a.tc_ntok == 32 ? 3 : (a.tc_ntok == 16 ? 6 : (a.tc_ntok == 128 ? 9 : 12))));
  if (TC) {
    const uint32_t raw = smem_u32(smem);
    tc.ringA = (raw + 1023u) & ~1023u;
    tc.ntok = (uint32_t)a.tc_ntok;
    tc.stages = tc.ntok > 128 ? 4u : (uint32_t)kTcStages;
    tc.bstride = tc.ntok > 128 ? 2u * kTcStageB : (uint32_t)kTcStageB;
    tc.ringB = tc.ringA + tc.stages * kTcStageA;
    tc.bars = smem_u32(tc_bars);
    tc.ring_gen = smem + (tc.ringA - raw);
    tc.kb_count = 0; tc.item_count = 0; tc.pre = 0;
    tc.pre_depth = min((uint32_t)max(a.tc_pre_depth, 0), tc.stages);
    tc.wpolicy = (a.dbg_flags & 8) ? 0ull : l2_policy_evict_first();
    if (tid == 32) {
      for (int s = 0; s < kTcStages; ++s) { mbar_init(tc_full(tc, s), 1); mbar_init(tc_empty(tc, s), 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(tc_acc_full(tc, i), 1); mbar_init(tc_acc_empty(tc, i), 8); }
      for (int i = 0; i < kTcConvSlots; ++i) { mbar_init(tc_conv_full(tc, i), kTcConvWarps); mbar_init(tc_conv_empty(tc, i), 1); }
      fence_barrier_init();
    }
    if ((tid >> 5) == 2) tmem_alloc<2 * kPersistTcMaxTokens>(smem_u32(&tc_tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tc.tmem = tc_tmem_slot;
    tc_prefetch_weights<W8>(tmaps, PQKV, PH, kTcSplitQkv, tc, kPersistQkvTileRows);
  }
  // attention: K/V chunk buffers overlay the (then idle) TMA ring / GEMM staging area, the small per-team arrays sit after it
  constexpr int ATW = (NT == 8) ? 8 : 16;                 // attention team width of this batch class
  const uint32_t ring_off = ((smem_u32(smem) + 1023u) & ~1023u) - smem_u32(smem);
  uint8_t* smem_kv = smem + ring_off;                     // 1024 B aligned (TMA swizzle atoms)
  uint8_t* smem_small = smem + ring_off + (TC ? (size_t)kTcStages * (kTcStageA + kTcStageB) : attn_kv_smem<ATW>());
  if (TC) tc.epi = reinterpret_cast<float*>(smem + ((ring_off + (size_t)kTcStages * (kTcStageA + kTcStageB) + 2 * attn_small_smem<8>() + 15) & ~(size_t)15));

  if (tid < kCtxSlots) s_ctx[tid] = (tid < B) ? a.gs.ctx_len[tid] : 0;     // visible after the first grid barrier's __syncthreads
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
  grid_barrier<TC>(a.bar, epoch); STAMP();

  for (int l = 0; l < a.n_layers; ++l) {
    const DecLayerDev L = a.layers[l];
    if (TC) {
      gemm_phase_tc<EPI_F32, W8>(tmaps + 4 * l, xmaps, PQKV, PH, kTcSplitQkv, B, Bpad, a.part, nullptr, tc, kPersistQkvTileRows, nullptr, 0, L.sqkv);
    } else gemm_dispatch<EPI_F32, W8, NT>(L.wqkv, L.sqkv, PQKV, PH, a.u, B, Bpad, a.part, nullptr, nullptr, smem);
    grid_barrier<TC>(a.bar, epoch); STAMP();
    if (a.attn_chunks > 1) attention_phase<TC ? kTcSplitQkv : 1>(a, L, TC ? smem_kv : smem, s_ctx, (a.timestamps && l == 1 && (int)blockIdx.x == a.dbg_cta) ? a.timestamps + 1024 + 900 : nullptr);   // few segments: split the keys over CTAs
    else attention_phase_mma<TC ? kTcSplitQkv : 1, ATW>(a, L, smem_kv, smem_small, smem_u32(attn_bars), attn_chunks_done, s_ctx);
    if (TC) {
      // the attention phase read / wrote the ring's shared memory through the generic proxy; TMA overwrites it next
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      tc_prefetch_weights<W8>(tmaps + 4 * l + 1, PH, PH, kTcSplitO, tc, kPersistOTileRows);      // the ring is free again
    }
    grid_barrier<TC>(a.bar, epoch); STAMP();
    if (TC) {
      gemm_phase_tc<EPI_F32, W8>(tmaps + 4 * l + 1, xmaps + 1, PH, PH, kTcSplitO, B, Bpad, a.part, nullptr, tc, kPersistOTileRows, nullptr, 0, L.so);
      tc_prefetch_weights<W8>(tmaps + 4 * l + 2, 2 * PI, PH, 1, tc, kPersistGuTileRows);
      grid_barrier<TC>(a.bar, epoch); STAMP();
      residual_norm_phase<kTcSplitO>(a.part, B, Bpad, a.x, a.u, L.rms2, a.eps, red);
    } else {
      gemm_dispatch<EPI_RESID, W8, NT>(L.wo, L.so, PH, PH, a.attn, B, Bpad, nullptr, a.x, nullptr, smem);
      grid_barrier<TC>(a.bar, epoch); STAMP();
      residual_norm_phase<0>(nullptr, B, Bpad, a.x, a.u, L.rms2, a.eps, red);
    }
    grid_barrier<TC>(a.bar, epoch); STAMP();
    // debug (dbg_cta == -2): per-CTA stamps around the gate/up phase of layer 1: [cta] start, [160 + cta] work done, [320 + cta] barrier passed
    const bool dbg_all = a.timestamps && a.dbg_cta == -2 && l == 1;      // CTA-uniform
    if (dbg_all && tid == 0) a.timestamps[1024 + blockIdx.x] = gtimer();
    if (TC) {
      gemm_phase_tc<EPI_SWIGLU, W8>(tmaps + 4 * l + 2, xmaps, 2 * PI, PH, 1, B, Bpad, nullptr, a.act, tc, kPersistGuTileRows,
                                (a.timestamps && l == 1 && (int)blockIdx.x == a.dbg_cta) ? a.timestamps + 1024 : nullptr, a.dbg_flags, L.sgu);
      const bool wd = a.timestamps && l == 1 && (int)blockIdx.x == a.dbg_cta && (tid & 31) == 0;
      if (wd) a.timestamps[1024 + 40 + (tid >> 5)] = gtimer();
      tc_prefetch_weights<W8>(tmaps + 4 * l + 3, PH, PI, kTcSplitDown, tc, kPersistDownTileRows);
      if (wd) a.timestamps[1024 + 60 + (tid >> 5)] = gtimer();
    } else gemm_dispatch<EPI_SWIGLU, W8, NT>(L.wgu, L.sgu, 2 * PI, PH, a.u, B, Bpad, nullptr, nullptr, a.act, smem);
    if (dbg_all) { __syncthreads(); if (tid == 0) a.timestamps[1024 + 160 + blockIdx.x] = gtimer(); }
    grid_barrier<TC>(a.bar, epoch, a.dbg_flags, dbg_all ? a.timestamps + 1024 : nullptr); STAMP();
    if (dbg_all && tid == 0) a.timestamps[1024 + 320 + blockIdx.x] = gtimer();
    const float* next_gamma = (l + 1 < a.n_layers) ? a.layers[l + 1].rms1 : a.final_norm;
    if (TC) {
      gemm_phase_tc<EPI_F32, W8>(tmaps + 4 * l + 3, xmaps + 2, PH, PI, kTcSplitDown, B, Bpad, a.part, nullptr, tc, kPersistDownTileRows, nullptr, 0, L.sdown);
      if (l + 1 < a.n_layers) tc_prefetch_weights<W8>(tmaps + 4 * (l + 1), PQKV, PH, kTcSplitQkv, tc, kPersistQkvTileRows);
      else {
        // the lm_head stays bf16 in int8 mode: its 16 KB stages overlay the raw stages / converter slots, so every MMA of the
        // int8 layout must have retired (all epilogue warps have drained their accumulators) before the first load lands
        if (W8) __syncthreads();
        tc_prefetch_weights<false>(tmaps + 4 * a.n_layers, PV_, PH, 1, tc, kPersistLmTileRows);
      }
      grid_barrier<TC>(a.bar, epoch); STAMP();
      residual_norm_phase<kTcSplitDown>(a.part, B, Bpad, a.x, a.u, next_gamma, a.eps, red);
    } else {
      gemm_dispatch<EPI_F32, W8, NT>(L.wdown, L.sdown, PH, PI, a.act, B, Bpad, a.part, nullptr, nullptr, smem);
      grid_barrier<TC>(a.bar, epoch); STAMP();
      residual_norm_phase<3>(a.part, B, Bpad, a.x, a.u, next_gamma, a.eps, red);
    }
    grid_barrier<TC>(a.bar, epoch); STAMP();
  }

  // ---- lm_head + greedy pick: every CTA scans its slice of the vocabulary for all tokens, CTA b merges token b
  if (TC) gemm_phase_tc<EPI_F32, false>(tmaps + 4 * a.n_layers, xmaps, PV_, PH, 1, B, Bpad, a.part, nullptr, tc, kPersistLmTileRows);
  else gemm_dispatch<EPI_F32, false, NT>(a.lm_head, nullptr, PV_, PH, a.u, B, Bpad, a.part, nullptr, nullptr, smem);
  grid_barrier<TC>(a.bar, epoch); STAMP();
  {
    const int per = (PV_ + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = min(PV_, lo + per);
    const int lane = tid & 31, warp = tid >> 5;
    for (int b = warp; b < B; b += kPWarps) {                       // one warp per token over this CTA's slice
      float best = -INFINITY, second = -INFINITY;
      int bi = 0x7fffffff;
      for (int i0 = lo + lane; i0 < hi; i0 += 32 * 16) {             // 16 independent L2 loads in flight per lane
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = (i0 + 32 * u < hi) ? __ldcg(a.part + (size_t)b * PV_ + i0 + 32 * u) : -INFINITY;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int i = i0 + 32 * u;
          if (i < hi) {
            if (a.logits_out) a.logits_out[(size_t)b * PV_ + i] = v[u];
            if (v[u] > best) { second = best; best = v[u]; bi = i; }
            else if (v[u] > second) second = v[u];
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { second = fmaxf(fmaxf(second, os), best); best = ob; bi = oi; }
        else { second = fmaxf(second, ob); }
      }
      if (lane == 0) {
        float* pp = a.pick_scratch + ((size_t)b * gridDim.x + blockIdx.x) * 4;
        pp[0] = best; pp[1] = second; pp[2] = __int_as_float(bi);
      }
    }
  }
  grid_barrier<false>(a.bar, epoch); STAMP();
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    if (tid < 32) {
      float best = -INFINITY, second = -INFINITY;
      int bi = 0x7fffffff;
      for (int c = tid; c < (int)gridDim.x; c += 32) {
        const float* pp = a.pick_scratch + ((size_t)b * gridDim.x + c) * 4;
        const float ob = __ldcg(pp), os = __ldcg(pp + 1);
        const int oi = __float_as_int(__ldcg(pp + 2));
        if (ob > best || (ob == best && oi < bi)) { second = fmaxf(fmaxf(second, os), best); best = ob; bi = oi; }
        else { second = fmaxf(second, ob); }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { second = fmaxf(fmaxf(second, os), best); best = ob; bi = oi; }
        else { second = fmaxf(second, ob); }
      }
      if (tid == 0) {
        const int step = *a.gs.step;
        if (bi < 0 || bi >= PV_) bi = a.gs.eos[0];         // all-NaN logits: emit EOS rather than an out-of-range id
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
    }
  }
  grid_barrier<false>(a.bar, epoch); STAMP();
  if (blockIdx.x == 0 && tid == 0) *a.gs.step += 1;
  if (TC) {
    tc_fence_before();
    __syncthreads();
    if ((tid >> 5) == 2) tmem_dealloc<2 * kPersistTcMaxTokens>(tc.tmem);
  }
}

// dynamic shared memory of variant i (persist_variant): the register-streaming classes that do not stage activations (NT = 1: <= 8
// segments, NT = 8: 33..64) need only the attention buffers and the 64 KB fragment exchange — 133 KB, which selects the 164 KB
// carve-out and leaves 92 KB of L1 for the activation rows they re-read from global memory; the staged classes (NT = 2, 4) add the
// 130 KB staging area.
static size_t persist_smem_for(int variant) {
  const bool tc = variant >= 8;
  const int nt = tc ? 8 : (1 << (variant & 3));
  const size_t attn = (size_t)AKEYS * kAKRow + (size_t)AKEYS * PHD * 2 + (PG * PHD + PG * AKEYS + 2 * PHD + 32 + 16) * 4;
  const size_t xchg = (size_t)kPWarps * 2 * 4 * 128 * 4;                            // 16 warps x NT(4) x 128 fp32 x RBW 2 (>= the NT = 8 exchange)
  const size_t exch = ((nt == 2 || nt == 4) ? (size_t)kXStageBytes : 0) + xchg;     // staged activations + fragment exchange
  const size_t attn_mma16 = 1024 + attn_kv_smem<16>() + attn_small_smem<16>(), attn_mma8 = 1024 + attn_kv_smem<8>() + 2 * attn_small_smem<8>();
  size_t m = attn > exch ? attn : exch;
  if (attn_mma16 > m) m = attn_mma16;
  if (attn_mma8 > m) m = attn_mma8;
  // tcgen05 class: TMA ring (the attention K/V buffers overlay it) + the per-team attention arrays after it
  if (tc) m = kTcRingBytes + 2 * attn_small_smem<8>() + 16 + kTcEpiBytes;
  return m;
}
size_t decode_persist_smem_bytes() { return persist_smem_for(8); }

size_t decode_persist_part_floats(int Bpad) { return (size_t)PV_ * Bpad; }     // >= qkv (3072) and 3 down sections (6144)
size_t decode_persist_pick_floats(int max_batch, int num_sms) { return (size_t)max_batch * num_sms * 4; }

typedef void (*PersistKernel)(DecodePersistArgs);
static constexpr int kPersistVariants = 10;
static PersistKernel persist_variant(int i) {
  static const PersistKernel tab[kPersistVariants] = {
      decode_persist_kernel<false, 1, false>, decode_persist_kernel<false, 2, false>, decode_persist_kernel<false, 4, false>,
      decode_persist_kernel<false, 8, false>, decode_persist_kernel<true, 1, false>,  decode_persist_kernel<true, 2, false>,
      decode_persist_kernel<true, 4, false>,  decode_persist_kernel<true, 8, false>,  decode_persist_kernel<false, 8, true>,
      decode_persist_kernel<true, 8, true>};
  return tab[i];
}
static PersistKernel persist_kernel_for(bool w8, int B, bool tc) {
  const int cls = B <= 8 ? 0 : (B <= 16 ? 1 : (B <= 32 ? 2 : 3));
  if (tc) return persist_variant(w8 ? 9 : 8);                     // the caller passed tensor maps: tcgen05 phases
  return persist_variant((w8 ? 4 : 0) + cls);
}

int decode_persist_occupancy() {
  int worst = 1 << 30;
  for (int i = 0; i < kPersistVariants; ++i) {
    int per_sm = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, persist_variant(i), kPThreads, persist_smem_for(i));
    if (per_sm < worst) worst = per_sm;
  }
  return worst;
}
// largest cooperative grid (<= one CTA per SM) the device can hold for this kernel right now; 0 if it cannot be launched
int decode_persist_max_grid(int num_sms) {
  int coop = 0, dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  if (!coop) return 0;
  return decode_persist_occupancy() >= 1 ? num_sms : 0;
}

cudaError_t decode_persist_configure() {
  for (int i = 0; i < kPersistVariants; ++i)
    SONIC_CUDA_TRY(cudaFuncSetAttribute(persist_variant(i), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)persist_smem_for(i)));
  return cudaSuccess;
}

cudaError_t launch_decode_persist(const DecodePersistArgs& a, int grid, cudaStream_t st, int* mode) {
  if (grid < 1 || a.B < 1 || a.B > (a.tmaps ? kPersistTcMaxTokens : 64)) return cudaErrorInvalidConfiguration;
  const int num_sms = grid;
  PersistKernel kern = persist_kernel_for(a.w8 != 0, a.B, a.tmaps != nullptr);
  int variant = 0;
  while (variant < kPersistVariants - 1 && persist_variant(variant) != kern) ++variant;
  const size_t smem_bytes = persist_smem_for(variant);
  {
    cudaError_t me = cudaMemsetAsync(a.bar, 0, sizeof(unsigned), st);
    if (me != cudaSuccess) { fprintf(stderr, "[sonicscribe_b200] barrier memset failed: %s\n", cudaGetErrorName(me)); return me; }
  }
  // *mode (per handle): 0: cudaLaunchKernelEx + cooperative attribute, 1: cudaLaunchCooperativeKernel.  Both guarantee
  // co-residency of the grid (the software grid barrier spins); there is no plain-launch variant — when both are refused
  // the caller falls back to the CUDA-graph decode path.
  cudaError_t e = cudaErrorUnknown;
  for (; *mode < 2; ++*mode) {
    if (*mode == 0) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(num_sms); cfg.blockDim = dim3(kPThreads); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
      cudaLaunchAttribute attr[1];
      memset(attr, 0, sizeof(attr));
      attr[0].id = cudaLaunchAttributeCooperative;
      attr[0].val.cooperative = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      e = cudaLaunchKernelEx(&cfg, kern, a);
    } else {
      DecodePersistArgs copy = a;
      void* args[1] = {&copy};
      e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(num_sms), dim3(kPThreads), args, smem_bytes, st);
    }
    if (e == cudaSuccess) return e;
    fprintf(stderr, "[sonicscribe_b200] decode_persist launch mode %d failed: %s (%s)\n", *mode, cudaGetErrorName(e), cudaGetErrorString(e));
    cudaGetLastError();
  }
  return e;
}

}  // namespace sonic
