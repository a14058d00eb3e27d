// Row-sliced persistent greedy-decode step for up to 32 segments (bf16) / 16 segments (int8 weight-only linears): ONE
// cooperative kernel per generated token; every weight matrix is cut into 148 contiguous ROW slices (one per CTA) that are
// streamed over the FULL K dimension by TMA into a shared-memory ring and multiplied by mma.sync consumer warps that read
// the 128B-swizzled stages with ldmatrix (int8: 16 B loads expanded to bf16 in registers).
//
// Why this shape:
//   * a batch-1..32 decode step is a chain of ~140 dependent phases whose weight bytes (96 MB per layer) take 15 us at HBM
//     speed while the dependency chain took 49 us with the split-K design of decode_persist.cu.  Here no phase produces
//     split-K partials in global memory, so residual add, RoPE + KV append, SwiGLU and the argmax are epilogues of the GEMM
//     that owns the rows, and RMSNorm is recomputed by every CTA while it builds its operand in shared memory: 5 grid
//     barriers per layer instead of 7, no fp32 partial traffic, no separate norm / embed / pick-scan phases;
//   * weights do not depend on the token: a dedicated producer warp walks the CTA's whole weight schedule (all layers) and
//     is throttled only by ring space, so the TMA stream keeps flowing through grid barriers, attention and epilogues (a TMA
//     ring sustains ~2x the per-SM bandwidth of the 16 B register loads of decode_persist.cu's mma.sync classes);
//   * why mma.sync and not tcgen05 for the multiplication: a tcgen05.mma costs 78-82 clocks whatever its shape up to N = 128
//     (profiles/r02_tcgen05_mma_rate.txt), and a full-K row slice of 14-24 rows needs 128-384 of them per phase while using
//     a quarter of the 64-row minimum tile: the first version of this kernel (tcgen05, TMEM accumulators) spent 3.4-10 us
//     per phase just issuing MMAs.  Eight consumer warps with m16n8k16 fragments have no such floor;
//   * accumulation order is fixed and independent of the batch: bit-identical ids for a segment alone or in any batch.
//
// Phases per layer:  P1 [u = rmsnorm(x) g1 -> smem] qkv slice, epilogue RoPE + q store + K/V append | barrier |
//                    P2 attention (split over key chunks and CTAs, last-arriver merge)              | barrier |
//                    P3 o-proj slice (operand: attention output by TMA), epilogue x += o            | barrier |
//                    P4 [u = rmsnorm(x) g2 -> smem] gate/up slice, epilogue SwiGLU -> act           | barrier |
//                    P5 down slice (operand: act by TMA, 1024-k chunks), epilogue x += d            | barrier |
// then lm_head slice with the argmax as epilogue | barrier | per-token merge + greedy bookkeeping.
//
// Warp roles (17 warps): w16 weight producer (TMA); w0 activation-chunk loader (TMA); w8-15 consumers (ldmatrix + mma.sync,
// cross-warp reduction through shared memory, fused epilogues); w0-15 build the normalised operand and run attention.
//
// Replaces, for one new token per segment: LlamaDecoderLayer x28 + final norm + lm_head + argmax/EOS bookkeeping
// (transformers/models/llama/modeling_llama.py:53-499, transformers/generation/utils.py:2743-2809), with the reference's
// bf16 rounding points (linear outputs, RoPE products, SiLU, residual sums are rounded to bf16 where HF materialises a
// bf16 tensor).
#include <cuda.h>
#include "common.cuh"
#include "kernels.h"
#include "gemm_tc.h"
#include "tc_ptx.cuh"

namespace sonic {

namespace {

constexpr int RH = 2048, RQKV = 3072, RI = 6144, RV = 59264, RHD = 128, RKVH = 4, RG = 4;
constexpr int kRsThreads = 544, kRsWork = 512;            // 16 worker warps + the producer warp
constexpr int kStageBytes = 16384;
constexpr int kAttnCK = 64;                               // keys per attention chunk
constexpr int kRopePairs = 16;                            // q/k row pairs per CTA covered by the shared-memory RoPE table
constexpr int kAttnKRow = RHD * 2 + 16;                   // padded K row (bytes): conflict-free 16 B reads across keys

// ---- bounded waits: a protocol bug must surface as a launch failure, not as a hung GPU -------------------------------
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// fine-grained stamps of one layer (a.dbg_layer) on CTA a.dbg_cta: slot = 16 * phase + point (see scripts/rs_phases.py)
#define RS_DBG(phase, point)                                                                                          \
  do {                                                                                                                \
    if (a.dbg && (int)blockIdx.x == a.dbg_cta && c.layer == a.dbg_layer) a.dbg[16 * (phase) + (point)] = gtime();   \
  } while (0)

__device__ __noinline__ void rs_timeout(int tag, unsigned v) {
  printf("[sonicscribe_b200] decode_rs: wait %d timed out (block %d thread %d, value %u)\n", tag, blockIdx.x, threadIdx.x, v);
  __trap();
}
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity, int tag) {
  uint32_t done;
  unsigned long long t0 = 0;
  for (unsigned it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((it & 1023u) == 1023u) {
      const unsigned long long t = gtime();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000ull) rs_timeout(tag, parity);
    }
  }
}
__device__ __forceinline__ void sync_workers() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void sync_consumers() { asm volatile("bar.sync 2, 256;" ::: "memory"); }
// PROXY: the next phase reads this phase's global writes through TMA (async proxy): order the generic writes before it
template <bool PROXY>
__device__ __forceinline__ void grid_barrier_rs(unsigned* counter, unsigned& epoch, unsigned long long* dbg3 = nullptr) {
  if (PROXY) asm volatile("fence.proxy.async;" ::: "memory");
  sync_workers();
  if (threadIdx.x == 0) {
    if (dbg3) dbg3[0] = gtime();
    epoch += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    if (dbg3) dbg3[1] = gtime();
    unsigned v;
    unsigned long long t0 = 0;
    for (unsigned it = 0;; ++it) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v >= epoch) break;
      if ((it & 1023u) == 1023u) {
        const unsigned long long t = gtime();
        if (t0 == 0) t0 = t;
        else if (t - t0 > 4000000000ull) rs_timeout(100, v);
      }
    }
    __threadfence();
    if (dbg3) dbg3[2] = gtime();
  }
  sync_workers();
  if (PROXY) asm volatile("fence.proxy.async;" ::: "memory");
}

__device__ __forceinline__ float bf16r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

// ---- geometry of one weight matrix as seen by one CTA ----------------------------------------------------------------------
enum { MAT_QKV = 0, MAT_O = 1, MAT_GU = 2, MAT_DOWN = 3, MAT_HEAD = 4, MAT_HEAD_TAIL = 5 };
struct Geom {
  const CUtensorMap* map;
  int r0, r1;        // this CTA's rows [r0, r1)
  int box;           // rows per TMA box == rows per sub-tile
  int nkb;           // 64-wide k blocks
  int kps;           // k blocks per ring stage
  int sub;           // bytes between the tiles of consecutive k blocks inside a ring stage (bf16: 1024-aligned; int8: 128-aligned)
  int rtiles;        // 16-row tiles per sub-tile
  int tpw;           // 16-row tiles per consumer warp (1 or 2)
  int rsplit;        // consumer warps along rows (1, 2 or 4); the other 8 / rsplit split the k blocks
  int i8;            // weights arrive as int8
};
// box rows per matrix kind (multiples of 8 that cover the largest slice of a 148-CTA grid; other grids walk sub-tiles)
__host__ __device__ constexpr int rs_box(int kind) { return kind == MAT_QKV ? 24 : kind == MAT_O ? 16 : kind == MAT_GU ? 88 : kind == MAT_DOWN ? 16 : kind == MAT_HEAD ? 128 : 32; }

template <bool W8>
__device__ __forceinline__ Geom make_geom(const CUtensorMap* maps, int n_layers, int layer, int kind) {
  Geom g;
  const int G = gridDim.x, c = blockIdx.x;
  g.box = rs_box(kind);
  g.i8 = (W8 && kind <= MAT_DOWN) ? 1 : 0;
  g.nkb = (kind == MAT_DOWN) ? RI / 64 : RH / 64;
  if (kind == MAT_QKV) { g.r0 = 2 * (int)(((long long)c * (RQKV / 2)) / G); g.r1 = 2 * (int)(((long long)(c + 1) * (RQKV / 2)) / G); }
  else if (kind == MAT_GU) { g.r0 = 2 * (int)(((long long)c * RI) / G); g.r1 = 2 * (int)(((long long)(c + 1) * RI) / G); }
  else if (kind == MAT_O || kind == MAT_DOWN) { g.r0 = (int)(((long long)c * RH) / G); g.r1 = (int)(((long long)(c + 1) * RH) / G); }
  else {
    const int a0 = (int)(((long long)c * RV) / G), a1 = (int)(((long long)(c + 1) * RV) / G);
    const int full = (a1 - a0) / 128 * 128;                           // 128-row tiles first, the remainder in 32-row boxes
    if (kind == MAT_HEAD) { g.r0 = a0; g.r1 = a0 + full; } else { g.r0 = a0 + full; g.r1 = a1; }
  }
  g.map = maps + ((kind <= MAT_DOWN) ? 4 * layer + kind : 4 * n_layers + (kind - MAT_HEAD));
  g.sub = g.i8 ? ((g.box * 64 + 127) & ~127) : ((g.box * 128 + 1023) & ~1023);
  int kps = kStageBytes / g.sub;
  if (kps > 16) kps = 16;
  while (g.nkb % kps || 16 % kps) --kps;                               // divides the k blocks of the matrix and of a 16-block activation chunk
  g.kps = kps;
  g.rtiles = (g.box + 15) / 16;
  g.tpw = g.rtiles >= 5 ? 2 : 1;                                       // 88 / 128 rows: two tiles per warp, k split in two (6 or 8 busy warps)
  g.rsplit = g.rtiles >= 3 ? 4 : g.rtiles;
  return g;
}

struct RsCtx {
  uint32_t ring, region, bars;
  uint8_t* region_g;     // generic pointer to the operand region (also the attention scratch)
  uint8_t* ring_g;
  float* scratch;        // [8 k-slices][rows][NTOK + 1] partial sums of the consumer warps
  int n_stages;
  uint32_t cnt;          // ring stages consumed so far (all matrices)
  uint32_t au[2];        // activation half-region loads so far
  int layer;             // current layer (debug stamps)
  int phase;             // current phase of the layer (debug stamps)
};
// barrier slots: full[NS] empty[NS] act_full[2] act_empty[2]
__device__ __forceinline__ uint32_t b_full(const RsCtx& c, uint32_t s) { return c.bars + 8u * s; }
__device__ __forceinline__ uint32_t b_empty(const RsCtx& c, uint32_t s) { return c.bars + 8u * (c.n_stages + s); }
__device__ __forceinline__ uint32_t b_misc(const RsCtx& c, uint32_t i) { return c.bars + 8u * (2 * c.n_stages + i); }
enum { ACT_FULL = 0, ACT_EMPTY = 2, N_MISC = 4 };
constexpr int kConsumers = 8;                                          // warps 8..15

// ---- the weight producer: one elected thread walks the CTA's whole schedule ---------------------------------------------------
// Cursor over the CTA's weight schedule: every (layer, matrix, sub-tile, stage group) in consumption order.
template <bool W8>
struct WCursor {
  const CUtensorMap* maps; int n_layers;
  int layer, kind, r, kb0;
  Geom g;
  bool done;
  unsigned long long bytes;                 // bytes of the schedule before the cursor
  __device__ __forceinline__ void load_geom() {
    for (;;) {
      if (layer >= n_layers && kind < MAT_HEAD) { kind = MAT_HEAD; }
      if (kind > MAT_HEAD_TAIL) { done = true; return; }
      g = make_geom<W8>(maps, n_layers, layer < n_layers ? layer : 0, kind);
      if (g.r0 < g.r1) { r = g.r0; kb0 = 0; return; }
      advance_matrix();
    }
  }
  __device__ __forceinline__ void advance_matrix() {
    if (kind < MAT_DOWN) ++kind;
    else if (kind == MAT_DOWN) { kind = MAT_QKV; ++layer; }
    else ++kind;                                                         // lm_head, lm_head tail, end
  }
  __device__ __forceinline__ void init(const CUtensorMap* m, int nl) { maps = m; n_layers = nl; layer = 0; kind = MAT_QKV; done = false; bytes = 0; load_geom(); }
  __device__ __forceinline__ uint32_t stage_bytes() const { return (uint32_t)g.kps * g.box * (g.i8 ? 64u : 128u); }
  __device__ __forceinline__ void next() {
    bytes += stage_bytes();
    kb0 += g.kps;
    if (kb0 >= g.nkb) { kb0 = 0; r += g.box; if (r >= g.r1) { advance_matrix(); load_geom(); } }
  }
};
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1) : "memory");
}
// The ring holds ~100-150 KB per SM, a fifth of one layer's slice, so while a CTA sits in attention, epilogues and grid barriers
// its ring is full and its share of HBM idle.  A second cursor runs up to kPrefetchAhead bytes in front of the loads and pulls
// the coming boxes into L2 (126 MB: 148 x 384 KB = 57 MB), so the ring refills at L2 speed once it drains.
constexpr unsigned long long kPrefetchAhead = 384u << 10;

template <bool W8>
__device__ void producer_loop(const DecodeRsArgs& a, const RsCtx& c) {
  const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(a.wmaps);
  WCursor<W8> ld, pf;
  ld.init(maps, a.n_layers);
  pf.init(maps, a.n_layers);
  uint32_t cnt = 0;
  while (!ld.done) {
    while (a.l2_prefetch && !pf.done && pf.bytes < ld.bytes + kPrefetchAhead) {
      if (pf.bytes >= ld.bytes + (unsigned long long)c.n_stages * kStageBytes) {        // what the ring itself will request soon is not prefetched
#pragma unroll 1
        for (int j = 0; j < pf.g.kps; ++j) tma_prefetch_2d(pf.g.map, (pf.kb0 + j) * 64, pf.r);
      }
      pf.next();
    }
    const uint32_t s = cnt % c.n_stages, par = (cnt / c.n_stages) & 1u;
    mbar_wait_wd(b_empty(c, s), par ^ 1u, 1);
    mbar_expect_tx(b_full(c, s), ld.stage_bytes());
    const uint32_t dst = c.ring + s * kStageBytes;
#pragma unroll 1
    for (int j = 0; j < ld.g.kps; ++j) tma_load_2d(dst + j * ld.g.sub, ld.g.map, b_full(c, s), (ld.kb0 + j) * 64, ld.r);
    if (a.dbg && (int)blockIdx.x == a.dbg_cta && ld.layer == a.dbg_layer && ld.kind <= MAT_DOWN && ld.kb0 + ld.g.kps >= ld.g.nkb && ld.r + ld.g.box >= ld.g.r1)
      a.dbg[16 * (ld.kind == MAT_QKV ? 0 : ld.kind + 1) + 10] = gtime();
    ld.next();
    ++cnt;
  }
}

// ---- operand builders ------------------------------------------------------------------------------------------------------
// u[t] = gamma * bf16(x[t] * rsqrt(mean(x[t]^2) + eps)) for every token, written as the K-major SWIZZLE_128B operand the
// consumers read: k block kb is a [NTOK rows x 128 B] slab, 16 B chunk ch of row t sits at chunk (ch ^ (t & 7)).  One warp per token.
// from_embed: x[t] is the embedding row of the token picked by the previous step (and is stored to x by one CTA).
struct GammaRegs { uint4 v[8]; };
// bf16 image of an RMSNorm weight vector, the 8 chunks this lane multiplies with; issued BEFORE the grid barrier that precedes
// the build (the vector is read once per step, so the load is a DRAM miss: ~1.5 us when it sits in the dependency chain)
__device__ __forceinline__ void load_gamma(GammaRegs& gm, const bf16* __restrict__ gamma, int B) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < B) {
#pragma unroll
    for (int i = 0; i < 8; ++i) gm.v[i] = __ldg(reinterpret_cast<const uint4*>(gamma) + lane + 32 * i);
  }
}

template <int NTOK>
__device__ __forceinline__ void build_norm_operand(const DecodeRsArgs& a, const RsCtx& c, const GammaRegs& gm, bool from_embed) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = warp; t < a.B; t += 16) {
    const bf16* row;
    if (from_embed) {
      int tok = a.gs.cur_tok[t];
      tok = tok < 0 ? 0 : (tok >= RV ? RV - 1 : tok);
      row = a.embed + (size_t)tok * RH;
    } else row = a.x + (size_t)t * RH;
    uint4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldcg(reinterpret_cast<const uint4*>(row) + lane + 32 * i);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float lo = bf_lo(w[e]), hi = bf_hi(w[e]); ss = fmaf(lo, lo, ss); ss = fmaf(hi, hi, ss); }
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss * (1.0f / RH) + a.eps);
    if (from_embed && (int)blockIdx.x == t % (int)gridDim.x) {
#pragma unroll
      for (int i = 0; i < 8; ++i) *(reinterpret_cast<uint4*>(a.x + (size_t)t * RH) + lane + 32 * i) = v[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ch = lane + 32 * i;                       // 16 B chunk of the row: k = 8 * ch
      const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w}, gw[4] = {gm.v[i].x, gm.v[i].y, gm.v[i].z, gm.v[i].w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = pack_bf16(bf_lo(gw[e]) * bf16r(bf_lo(w[e]) * rstd), bf_hi(gw[e]) * bf16r(bf_hi(w[e]) * rstd));
      const int kb = ch >> 3, cc = ch & 7;
      *reinterpret_cast<uint4*>(c.region_g + (size_t)kb * (NTOK * 128) + t * 128 + ((cc ^ (t & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  fence_proxy_async_smem();
  sync_workers();
}

// ---- consumers -------------------------------------------------------------------------------------------------------------
enum { EPI_QKV = 0, EPI_RESID = 1, EPI_SWIGLU = 2, EPI_HEAD = 3 };
__device__ __forceinline__ void pick_merge2(float& mb, float& ms, int& mi, float ob, float os, int oi) {
  if (ob > mb || (ob == mb && oi < mi)) { ms = fmaxf(fmaxf(ms, os), mb); mb = ob; mi = oi; }
  else { ms = fmaxf(ms, ob); }
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// two's complement bytes -> bf16 pairs without integer conversion: b = low7 - 128 * bit7; bf16(0x4300 | low7) = 128 + low7 and
// bf16(0x4300 | bit7 << 7) = 128 or 256 are exact, and so is their difference
__device__ __forceinline__ void i8x4_to_bf16x2(uint32_t w, uint32_t& lo, uint32_t& hi) {
  const uint32_t lo2 = __byte_perm(w, 0x43434343u, 0x4140), hi2 = __byte_perm(w, 0x43434343u, 0x4342);
  const uint32_t m1 = 0xbf80bf80u;                                     // -1.0 in both halves
  __nv_bfloat162 alo, blo, ahi, bhi;
  *reinterpret_cast<uint32_t*>(&alo) = lo2 & 0xff7fff7fu; *reinterpret_cast<uint32_t*>(&blo) = lo2 & 0xff80ff80u;
  *reinterpret_cast<uint32_t*>(&ahi) = hi2 & 0xff7fff7fu; *reinterpret_cast<uint32_t*>(&bhi) = hi2 & 0xff80ff80u;
  const __nv_bfloat162 rlo = __hfma2(blo, *reinterpret_cast<const __nv_bfloat162*>(&m1), alo);
  const __nv_bfloat162 rhi = __hfma2(bhi, *reinterpret_cast<const __nv_bfloat162*>(&m1), ahi);
  lo = *reinterpret_cast<const uint32_t*>(&rlo);
  hi = *reinterpret_cast<const uint32_t*>(&rhi);
}

// one 64-wide k block of one 16-row tile: acc[nt] += W[16 rows][64 k] . U[8 tokens of n-tile nt][64 k]^T
// bf16: A fragments by ldmatrix from the SWIZZLE_128B stage tile, B fragments by ldmatrix from the operand slab (same swizzle)
template <int NTOK>
__device__ __forceinline__ void kblock_bf16(float (&acc)[NTOK / 8][4], uint32_t tile, uint32_t slab, int row0, int lane, int nt_used) {
  const int mi = lane >> 3, ri = lane & 7;
  const int arow = row0 + (mi & 1) * 8 + ri;                           // A: matrices (rows 0-7 | 8-15) x (k 0-7 | 8-15)
  const uint32_t abase = tile + arow * 128, asw = (uint32_t)(arow & 7);
  const int brow = (mi >> 1) * 8 + ri;                                 // B: matrices (n-tile 0 | 1) x (k 0-7 | 8-15)
  uint32_t af[4][4], bf[4][NTOK / 16][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {                                     // every fragment of the k block in flight before the first MMA
    ldsm_x4(abase + (((uint32_t)(2 * ks + (mi >> 1)) ^ asw) << 4), af[ks][0], af[ks][1], af[ks][2], af[ks][3]);
#pragma unroll
    for (int np = 0; np < NTOK / 16; ++np) {
      const int n = np * 16 + brow;
      ldsm_x4(slab + n * 128 + (((uint32_t)(2 * ks + (mi & 1)) ^ (uint32_t)(n & 7)) << 4), bf[ks][np][0], bf[ks][np][1], bf[ks][np][2], bf[ks][np][3]);
    }
  }
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < NTOK / 16; ++np) {
      if (2 * np < nt_used) mma16816(acc[2 * np], af[ks][0], af[ks][1], af[ks][2], af[ks][3], bf[ks][np][0], bf[ks][np][1]);           // uniform
      if (2 * np + 1 < nt_used) mma16816(acc[2 * np + 1], af[ks][0], af[ks][1], af[ks][2], af[ks][3], bf[ks][np][2], bf[ks][np][3]);
    }
  }
}
// int8: thread (g, t) owns bytes 16 t .. 16 t + 15 of rows g and g + 8 of the raw [rows][64 B] tile; the k order inside the
// block is permuted identically for weights and activations (a dot product does not care)
template <int NTOK>
__device__ __forceinline__ void kblock_i8(float (&acc)[NTOK / 8][4], const uint8_t* tile, const uint8_t* slab, int row0, int lane, int nt_used) {
  const int g = lane >> 2, t = lane & 3;
  const uint4 wa = *reinterpret_cast<const uint4*>(tile + (row0 + g) * 64 + 16 * t);
  const uint4 wb = *reinterpret_cast<const uint4*>(tile + (row0 + g + 8) * 64 + 16 * t);
  const uint32_t ra[4] = {wa.x, wa.y, wa.z, wa.w}, rb[4] = {wb.x, wb.y, wb.z, wb.w};
  uint32_t alo[4], ahi[4], blo[4], bhi[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) { i8x4_to_bf16x2(ra[s], alo[s], ahi[s]); i8x4_to_bf16x2(rb[s], blo[s], bhi[s]); }
#pragma unroll
  for (int nt = 0; nt < NTOK / 8; ++nt) {
    if (nt >= nt_used) break;                                           // uniform: 8-token tiles that hold no segment are skipped
    const int n = nt * 8 + g;
    const uint8_t* pr = slab + n * 128;
    const uint4 x0 = *reinterpret_cast<const uint4*>(pr + (((2 * t) ^ (n & 7)) << 4));
    const uint4 x1 = *reinterpret_cast<const uint4*>(pr + (((2 * t + 1) ^ (n & 7)) << 4));
    mma16816(acc[nt], alo[0], blo[0], ahi[0], bhi[0], x0.x, x0.y);
    mma16816(acc[nt], alo[1], blo[1], ahi[1], bhi[1], x0.z, x0.w);
    mma16816(acc[nt], alo[2], blo[2], ahi[2], bhi[2], x1.x, x1.y);
    mma16816(acc[nt], alo[3], blo[3], ahi[3], bhi[3], x1.z, x1.w);
  }
}

// epilogue of one sub-tile from the consumers' partial sums: item = (row pair, token); the 256 consumer threads walk the items
template <int NTOK, int EPI>
__device__ __forceinline__ void epilogue_rs(const DecodeRsArgs& a, const RsLayer& L, const RsCtx& c, const Geom& g, int r, const float* __restrict__ wscale,
                                            const int* s_pos, const float2* s_rope, float* s_pick) {
  constexpr int P = NTOK + 1;                                          // scratch row pitch (floats)
  const int tc = threadIdx.x - 256;
  const int nrows = min(g.box, g.r1 - r);
  const int ksw = kConsumers / g.rsplit;
  const int slice = g.rtiles * 16 * P;                                 // floats per k-slice
  if constexpr (EPI == EPI_HEAD) {
    // argmax per token: warp w takes tokens w, w + 8, ...; lanes stride the rows
    const int lane = tc & 31, w = tc >> 5;
    for (int t = w; t < a.B; t += kConsumers) {
      float best = -INFINITY, second = -INFINITY;
      int bi = 0x7fffffff;
      for (int rl = lane; rl < nrows; rl += 32) {
        float x = 0.f;
        for (int ks = 0; ks < ksw; ++ks) x += c.scratch[ks * slice + rl * P + t];
        if (!(x == x)) x = -INFINITY;                                   // NaN never wins
        if (a.logits_out) a.logits_out[(size_t)t * RV + r + rl] = x;
        if (x > best) { second = best; best = x; bi = r + rl; }
        else if (x > second) second = x;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        pick_merge2(best, second, bi, ob, os, oi);
      }
      if (lane == 0) {                                                  // token t is always handled by this warp: private running state
        float* sp = s_pick + t * 3;
        float mb = sp[0], ms = sp[1];
        int mi = __float_as_int(sp[2]);
        pick_merge2(mb, ms, mi, best, second, bi);
        sp[0] = mb; sp[1] = ms; sp[2] = __int_as_float(mi);
      }
    }
    return;
  }
  const int pairs = (nrows + 1) >> 1;
  for (int it = tc; it < pairs * a.B; it += 32 * kConsumers) {
    const int p = it / a.B, t = it - p * a.B;
    const int rl = 2 * p, gr = r + rl;
    float v0 = 0.f, v1 = 0.f;
    for (int ks = 0; ks < ksw; ++ks) {                                  // k-slice order: deterministic
      v0 += c.scratch[ks * slice + rl * P + t];
      v1 += c.scratch[ks * slice + (rl + 1) * P + t];
    }
    const bool has1 = rl + 1 < nrows;
    if (g.i8) { v0 *= __ldg(wscale + gr); if (has1) v1 *= __ldg(wscale + gr + 1); }
    if constexpr (EPI == EPI_QKV) {
      // rows of q and k heads are interleaved (2j, 2j+1) = natural dims (j, j+64); slices are even-aligned, so has1 holds
      const float x = bf16r(v0), y = bf16r(v1);
      const int pos = s_pos[t];
      if (gr < RQKV - RKVH * RHD) {
        const int hb = gr >> 7, j = (gr & 127) >> 1;
        float2 cs;
        const int pair = (gr - g.r0) >> 1;
        if ((g.r1 - g.r0) <= 2 * kRopePairs) cs = s_rope[pair * NTOK + t];
        else cs = make_float2(bf16r(__ldg(a.cos_t + (size_t)pos * (RHD / 2) + j)), bf16r(__ldg(a.sin_t + (size_t)pos * (RHD / 2) + j)));
        const bf16 o0 = __float2bfloat16_rn(bf16r(x * cs.x) - bf16r(y * cs.y));
        const bf16 o1 = __float2bfloat16_rn(bf16r(y * cs.x) + bf16r(x * cs.y));
        bf16* dst = (hb < 16) ? a.q + (size_t)t * RH + hb * RHD : L.kc + ((size_t)(t * RKVH + (hb - 16)) * a.max_ctx + pos) * RHD;
        dst[j] = o0; dst[j + 64] = o1;
      } else {
        const int vr = gr - (RQKV - RKVH * RHD);
        bf16* dst = L.vc + ((size_t)(t * RKVH + (vr >> 7)) * a.max_ctx + pos) * RHD + (vr & 127);
        dst[0] = __float2bfloat16_rn(x); dst[1] = __float2bfloat16_rn(y);
      }
    } else if constexpr (EPI == EPI_RESID) {
      bf16* px = a.x + (size_t)t * RH + gr;                            // this thread is the only writer of these two elements
      const float x0 = __uint_as_float((uint32_t)__ldcg(reinterpret_cast<const unsigned short*>(px)) << 16);
      px[0] = __float2bfloat16_rn(x0 + bf16r(v0));
      if (has1) {
        const float x1 = __uint_as_float((uint32_t)__ldcg(reinterpret_cast<const unsigned short*>(px + 1)) << 16);
        px[1] = __float2bfloat16_rn(x1 + bf16r(v1));
      }
    } else {                                                            // SwiGLU: rows are interleaved (gate, up)
      a.act[(size_t)t * RI + (gr >> 1)] = __float2bfloat16_rn(bf16r(silu_exact(bf16r(v0))) * bf16r(v1));
    }
  }
}

// ---- one GEMM phase: rows [g.r0, g.r1) of one matrix against the operand region ------------------------------------------
// ACT: the operand is a global bf16 activation [tokens][K] loaded by TMA in 1024-k half-region chunks (act_map); otherwise the
// region already holds the normalised operand for the whole K = 2048.
template <int NTOK, bool W8, int EPI, bool ACT>
__device__ __forceinline__ void gemm_phase_rs(const DecodeRsArgs& a, const RsLayer& L, RsCtx& c, const Geom& g, const CUtensorMap* act_map,
                                              const float* __restrict__ wscale, const int* s_pos, const float2* s_rope, float* s_pick) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_sub = (g.r1 - g.r0 + g.box - 1) / g.box;
  const int n_groups = g.nkb / g.kps;
  const int n_chunks = g.nkb / 16;                                    // activation chunks of 16 k blocks per sub-tile
  constexpr uint32_t kHalf = NTOK * 128 * 16;                         // bytes of one activation chunk
  constexpr int P = NTOK + 1;
  if (ACT && warp == 0) {
    uint32_t au[2] = {c.au[0], c.au[1]};
    for (int st = 0; st < n_sub; ++st)
      for (int ci = 0; ci < n_chunks; ++ci) {
        const int h = ci & 1;
        mbar_wait_wd(b_misc(c, ACT_EMPTY + h), (au[h] & 1u) ^ 1u, 2);
        if (elect_one_sync()) {
          mbar_expect_tx(b_misc(c, ACT_FULL + h), kHalf);
#pragma unroll
          for (int kb = 0; kb < 16; ++kb)
            tma_load_2d(c.region + h * kHalf + kb * (NTOK * 128), act_map, b_misc(c, ACT_FULL + h), (ci * 16 + kb) * 64, 0);
        }
        __syncwarp();
        ++au[h];
      }
    if (lane == 0) RS_DBG(c.phase, 11);
  } else if (warp >= 8) {
    const int wc = warp - 8;
    const int rg = wc % g.rsplit, ksl = wc / g.rsplit, ksw = kConsumers / g.rsplit;
    const int rt0 = rg * g.tpw;                                         // first 16-row tile of this warp
    const int nt_used = (a.B + 7) >> 3;
    const bool active = rt0 < g.rtiles, two = g.tpw == 2 && rt0 + 1 < g.rtiles;
    uint32_t cnt = c.cnt;
    uint32_t au[2] = {c.au[0], c.au[1]};
    int r = g.r0;
    if (wc == 0 && lane == 0) RS_DBG(c.phase, 2);
    for (int st = 0; st < n_sub; ++st, r += g.box) {
      float acc[2][NTOK / 8][4];
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
        for (int nt = 0; nt < NTOK / 8; ++nt) { acc[h2][nt][0] = acc[h2][nt][1] = acc[h2][nt][2] = acc[h2][nt][3] = 0.f; }
      for (int grp = 0; grp < n_groups; ++grp, ++cnt) {
        const uint32_t s = cnt % c.n_stages, par = (cnt / c.n_stages) & 1u;
        const int kb0 = grp * g.kps;
        mbar_wait_wd(b_full(c, s), par, 5);
        if (ACT && (kb0 & 15) == 0) mbar_wait_wd(b_misc(c, ACT_FULL + ((kb0 >> 4) & 1)), au[(kb0 >> 4) & 1] & 1u, 6);   // kps divides 16
        if (st == 0 && grp == 0 && wc == 0 && lane == 0) RS_DBG(c.phase, 3);
        if (active) {
          // rows of this warp's tile(s); k blocks kb = kb0 + j with (kb % ksw) == ksl (a fixed partition: the sum order never changes)
          for (int j = (ksl - kb0) & (ksw - 1); j < g.kps; j += ksw) {        // ksw is a power of two
            const int kb = kb0 + j;
            const uint32_t slab = (uint32_t)(kb & 31) * (NTOK * 128);
            if (W8 && g.i8) {
              const uint8_t* tile = c.ring_g + (size_t)s * kStageBytes + (size_t)j * g.sub;
              kblock_i8<NTOK>(acc[0], tile, c.region_g + slab, rt0 * 16, lane, nt_used);
              if (two) kblock_i8<NTOK>(acc[1], tile, c.region_g + slab, rt0 * 16 + 16, lane, nt_used);
            } else {
              const uint32_t tile = c.ring + s * kStageBytes + j * g.sub;
              kblock_bf16<NTOK>(acc[0], tile, c.region + slab, rt0 * 16, lane, nt_used);
              if (two) kblock_bf16<NTOK>(acc[1], tile, c.region + slab, rt0 * 16 + 16, lane, nt_used);
            }
          }
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(b_empty(c, s));                                  // 8 consumer warps free a stage
          if (ACT && ((kb0 + g.kps) & 15) == 0) mbar_arrive(b_misc(c, ACT_EMPTY + (((kb0 + g.kps - 1) >> 4) & 1)));
        }
        if (ACT && ((kb0 + g.kps) & 15) == 0) ++au[((kb0 + g.kps - 1) >> 4) & 1];
      }
      if (wc == 0 && lane == 0) RS_DBG(c.phase, 4);
      // partial sums -> scratch[k-slice][row][token]
      if (active) {
        const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          if (h2 == 1 && !two) break;
          float* sp = c.scratch + ksl * (g.rtiles * 16 * P) + ((rt0 + h2) * 16 + gq) * P + 2 * tq;
#pragma unroll
          for (int nt = 0; nt < NTOK / 8; ++nt) {
            sp[nt * 8] = acc[h2][nt][0]; sp[nt * 8 + 1] = acc[h2][nt][1];
            sp[8 * P + nt * 8] = acc[h2][nt][2]; sp[8 * P + nt * 8 + 1] = acc[h2][nt][3];
          }
        }
      }
      sync_consumers();
      if (wc == 0 && lane == 0) RS_DBG(c.phase, 5);
      epilogue_rs<NTOK, EPI>(a, L, c, g, r, wscale, s_pos, s_rope, s_pick);
      if (wc == 0 && lane == 0) RS_DBG(c.phase, 6);
      if (st + 1 < n_sub || EPI == EPI_HEAD) sync_consumers();         // the scratch is rewritten by the next sub-tile
    }
  }
  // mirrored counters (every worker thread advances them identically)
  c.cnt += (uint32_t)(n_sub * n_groups);
  if (ACT) { c.au[0] += (uint32_t)(n_sub * ((n_chunks + 1) / 2)); c.au[1] += (uint32_t)(n_sub * (n_chunks / 2)); }
}

// ---- attention: item = (segment, kv head, chunk of 64 keys); partial (m, l, o) per item, merged by the item that arrives last
// at the (segment, kv head) counter, in chunk order (deterministic, independent of how many chunks the launch was given) ------
__device__ __forceinline__ void attention_phase_rs(const DecodeRsArgs& a, const RsLayer& L, uint8_t* scratch) {
  uint8_t* sK = scratch;                                                // kAttnCK x 272 B
  uint8_t* sV = scratch + kAttnCK * kAttnKRow;                          // kAttnCK x 256 B
  float* sQ = reinterpret_cast<float*>(sV + kAttnCK * RHD * 2);        // [4][128], pre-scaled
  float* sP = sQ + RG * RHD;                                            // [4][64]
  float* sM = sP + RG * kAttnCK;                                        // [16] per-warp maxima
  float* sL = sM + 16;                                                  // [16] per-warp sums
  int* s_last = reinterpret_cast<int*>(sL + 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = a.attn_chunks;
  const int n_items = a.B * RKVH * C;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int chunk = item % C, grp = item / C;
    const int seg = grp / RKVH, kvh = grp - seg * RKVH;
    const int pos = a.gs.ctx_len[seg], kv_len = pos + 1;
    const int n_chunks = (kv_len + kAttnCK - 1) / kAttnCK;
    float* ws = a.attn_ws + ((size_t)grp * C + chunk) * RG * (RHD + 2);
    if (chunk < n_chunks) {
      const bf16* kc = L.kc + ((size_t)seg * RKVH + kvh) * a.max_ctx * RHD;
      const bf16* vc = L.vc + ((size_t)seg * RKVH + kvh) * a.max_ctx * RHD;
      const int k0 = chunk * kAttnCK, nk = min(kAttnCK, kv_len - k0);
      if (tid < 64) {                                                   // q: 4 heads x 128 dims, 8 per thread
        const uint4 qv = __ldcg(reinterpret_cast<const uint4*>(a.q + (size_t)seg * RH + kvh * RG * RHD) + tid);
        const uint32_t w[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { sQ[tid * 8 + 2 * e] = bf_lo(w[e]) * a.scale; sQ[tid * 8 + 2 * e + 1] = bf_hi(w[e]) * a.scale; }
      }
      for (int i = tid; i < kAttnCK * 16; i += kRsWork) {
        const int r = i >> 4, c16 = i & 15;
        uint4 kk = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
        if (r < nk) {
          kk = __ldcg(reinterpret_cast<const uint4*>(kc + (size_t)(k0 + r) * RHD) + c16);
          vv = __ldcg(reinterpret_cast<const uint4*>(vc + (size_t)(k0 + r) * RHD) + c16);
        }
        *reinterpret_cast<uint4*>(sK + r * kAttnKRow + c16 * 16) = kk;
        *reinterpret_cast<uint4*>(sV + r * (RHD * 2) + c16 * 16) = vv;
      }
      sync_workers();
      // scores: pair p = tid / 2 -> (head, key); the two threads of a pair take 64 dims each
      const int p = tid >> 1, half = tid & 1, head = p >> 6, key = p & 63;
      float acc = 0.f;
      {
        const uint8_t* kr = sK + key * kAttnKRow + half * 128;
        const float* qh = sQ + head * RHD + half * 64;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          const uint4 kv = *reinterpret_cast<const uint4*>(kr + c8 * 16);
          const float4 q0 = *reinterpret_cast<const float4*>(qh + c8 * 8), q1 = *reinterpret_cast<const float4*>(qh + c8 * 8 + 4);
          acc = fmaf(bf_lo(kv.x), q0.x, acc); acc = fmaf(bf_hi(kv.x), q0.y, acc);
          acc = fmaf(bf_lo(kv.y), q0.z, acc); acc = fmaf(bf_hi(kv.y), q0.w, acc);
          acc = fmaf(bf_lo(kv.z), q1.x, acc); acc = fmaf(bf_hi(kv.z), q1.y, acc);
          acc = fmaf(bf_lo(kv.w), q1.z, acc); acc = fmaf(bf_hi(kv.w), q1.w, acc);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      const float sc = (key < nk) ? acc : -INFINITY;
      const float wmax = warp_max(sc);
      if (lane == 0) sM[warp] = wmax;                                   // warps 4h .. 4h+3 hold head h
      sync_workers();
      const float cmax = fmaxf(fmaxf(sM[head * 4], sM[head * 4 + 1]), fmaxf(sM[head * 4 + 2], sM[head * 4 + 3]));
      const float pr = (sc == -INFINITY) ? 0.f : expf(sc - cmax);
      if (half == 0) sP[head * kAttnCK + key] = pr;
      const float wsum = warp_sum(half == 0 ? pr : 0.f);
      if (lane == 0) sL[warp] = wsum;
      sync_workers();
      {
        const int ph = tid >> 7, d = tid & 127;                          // PV: thread = (head, dim)
        float o = 0.f;
        const float* pp = sP + ph * kAttnCK;
        const bf16* vcol = reinterpret_cast<const bf16*>(sV) + d;
        for (int j = 0; j < nk; ++j) o = fmaf(pp[j], __bfloat162float(vcol[j * RHD]), o);
        ws[(size_t)ph * (RHD + 2) + d] = o;
        if (d == 0) {
          ws[(size_t)ph * (RHD + 2) + RHD] = fmaxf(fmaxf(sM[ph * 4], sM[ph * 4 + 1]), fmaxf(sM[ph * 4 + 2], sM[ph * 4 + 3]));
          ws[(size_t)ph * (RHD + 2) + RHD + 1] = sL[ph * 4] + sL[ph * 4 + 1] + sL[ph * 4 + 2] + sL[ph * 4 + 3];
        }
      }
    }
    __threadfence();
    sync_workers();
    if (tid == 0) {
      const int prev = atomicAdd(a.attn_counters + grp, 1);
      *s_last = (prev == C - 1) ? 1 : 0;
      if (prev == C - 1) a.attn_counters[grp] = 0;
    }
    sync_workers();
    if (*s_last) {
      __threadfence();
      const int ph = tid >> 7, d = tid & 127;
      const float* wg = a.attn_ws + (size_t)grp * C * RG * (RHD + 2);
      float M = -INFINITY, Ls = 0.f, o = 0.f;
      if (n_chunks <= 16) {                                              // every partial in flight at once: one L2 round trip
        float mv[16], lv[16], ov[16];
#pragma unroll
        for (int cidx = 0; cidx < 16; ++cidx) {
          mv[cidx] = -INFINITY; lv[cidx] = 0.f; ov[cidx] = 0.f;
          if (cidx < n_chunks) {
            const float* pc = wg + ((size_t)cidx * RG + ph) * (RHD + 2);
            mv[cidx] = __ldcg(pc + RHD); lv[cidx] = __ldcg(pc + RHD + 1); ov[cidx] = __ldcg(pc + d);
          }
        }
#pragma unroll
        for (int cidx = 0; cidx < 16; ++cidx) M = fmaxf(M, mv[cidx]);
#pragma unroll
        for (int cidx = 0; cidx < 16; ++cidx) {                          // chunk order: deterministic
          if (cidx < n_chunks) { const float w = expf(mv[cidx] - M); Ls += lv[cidx] * w; o += ov[cidx] * w; }
        }
      } else {
        for (int cidx = 0; cidx < n_chunks; ++cidx) M = fmaxf(M, __ldcg(wg + ((size_t)cidx * RG + ph) * (RHD + 2) + RHD));
        for (int cidx = 0; cidx < n_chunks; ++cidx) {
          const float* pc = wg + ((size_t)cidx * RG + ph) * (RHD + 2);
          const float w = expf(__ldcg(pc + RHD) - M);
          Ls += __ldcg(pc + RHD + 1) * w;
          o += __ldcg(pc + d) * w;
        }
      }
      a.attn[(size_t)seg * RH + (size_t)(kvh * RG + ph) * RHD + d] = __float2bfloat16_rn(o / Ls);
    }
    sync_workers();
  }
}

#define RS_STAMP()                                                                                  \
  do {                                                                                              \
    if (a.timestamps && blockIdx.x == 0 && threadIdx.x == 0) a.timestamps[n_stamp++] = gtime();     \
  } while (0)

template <int NTOK, bool W8>
struct RsSmem {
  static constexpr int kRegion = NTOK * RH * 2;                          // operand region: 64 KB / 128 KB
  static constexpr int kScratch = 256 * (NTOK + 1) * 4;                  // consumer partial sums: max over matrices of k-slices x rows (2 x 128)
  static constexpr int kSmall = 1024 + kRopePairs * NTOK * 8 + kScratch; // barriers, positions, pick state, RoPE table, scratch
  static constexpr int kStages = (227 * 1024 - 1024 - kRegion - kSmall) / kStageBytes;
  static constexpr int kTotal = 1024 + kStages * kStageBytes + kRegion + kSmall;
  static_assert(kStages >= 3, "ring too shallow");
  static_assert(kRegion >= kAttnCK * kAttnKRow + kAttnCK * RHD * 2 + (RG * RHD + RG * kAttnCK + 40) * 4, "attention scratch does not fit the operand region");
};

template <int NTOK, bool W8>
__global__ void __launch_bounds__(kRsThreads, 1) decode_rs_kernel(DecodeRsArgs a) {
  using SM = RsSmem<NTOK, W8>;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sg = smem_raw + (base - raw);
  RsCtx c;
  c.n_stages = SM::kStages;
  c.ring = base; c.ring_g = sg;
  c.region = base + SM::kStages * kStageBytes; c.region_g = sg + SM::kStages * kStageBytes;
  uint8_t* small = c.region_g + SM::kRegion;
  c.bars = c.region + SM::kRegion;
  static_assert(8 * (2 * SM::kStages + N_MISC) + 256 + NTOK * 3 * 4 <= 1024, "small area overflow");
  int* s_pos = reinterpret_cast<int*>(small + 8 * (2 * SM::kStages + N_MISC));                   // [64] token positions
  float* s_pick = reinterpret_cast<float*>(s_pos + 64);                                          // [NTOK][3] running argmax state
  float2* s_rope = reinterpret_cast<float2*>(small + 1024);                                      // [kRopePairs][NTOK] (cos, sin)
  c.scratch = reinterpret_cast<float*>(small + 1024 + kRopePairs * NTOK * 8);
  c.cnt = 0; c.au[0] = 0; c.au[1] = 0; c.layer = -1; c.phase = 0;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < SM::kStages; ++s) { mbar_init(b_full(c, s), 1); mbar_init(b_empty(c, s), kConsumers); }
    for (int i = 0; i < 2; ++i) { mbar_init(b_misc(c, ACT_FULL + i), 1); mbar_init(b_misc(c, ACT_EMPTY + i), kConsumers); }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == 16) {                                                     // the weight stream never waits for a grid barrier
    if (elect_one_sync()) producer_loop<W8>(a, c);
    return;
  }

  unsigned epoch = 0;
  int n_stamp = 0;
  RS_STAMP();
  const CUtensorMap* wmaps = reinterpret_cast<const CUtensorMap*>(a.wmaps);
  const CUtensorMap* amaps = reinterpret_cast<const CUtensorMap*>(a.amaps);       // [0] attention output, [1] SwiGLU output
  // positions and the RoPE table of this CTA's q / k rows (same rows, same angles in every layer)
  if (tid < a.B) s_pos[tid] = a.gs.ctx_len[tid];
  GammaRegs gm;
  load_gamma(gm, a.layers[0].g1, a.B);
  sync_workers();
  {
    const Geom gq = make_geom<W8>(wmaps, a.n_layers, 0, MAT_QKV);
    const int pairs = (gq.r1 - gq.r0) >> 1;
    if (pairs <= kRopePairs) {
      for (int idx = tid; idx < pairs * a.B; idx += kRsWork) {
        const int p = idx / a.B, t = idx - p * a.B, row = gq.r0 + 2 * p;
        if (row < RQKV - RKVH * RHD) {
          const int j = (row & 127) >> 1, pos = s_pos[t];
          s_rope[p * NTOK + t] = make_float2(bf16r(__ldg(a.cos_t + (size_t)pos * (RHD / 2) + j)), bf16r(__ldg(a.sin_t + (size_t)pos * (RHD / 2) + j)));
        }
      }
    }
  }
  for (int l = 0; l < a.n_layers; ++l) {
    const RsLayer L = a.layers[l];
    c.layer = l;
    unsigned long long* bd = (a.dbg && (int)blockIdx.x == a.dbg_cta && l == a.dbg_layer) ? a.dbg : nullptr;
    c.phase = 0;
    if (tid == 0) RS_DBG(0, 0);
    build_norm_operand<NTOK>(a, c, gm, l == 0);
    if (tid == 0) RS_DBG(0, 1);
    gemm_phase_rs<NTOK, W8, EPI_QKV, false>(a, L, c, make_geom<W8>(wmaps, a.n_layers, l, MAT_QKV), nullptr, L.s_qkv, s_pos, s_rope, s_pick);
    grid_barrier_rs<false>(a.bar, epoch, bd ? bd + 7 : nullptr); RS_STAMP();
    c.phase = 1;
    if (tid == 0) RS_DBG(1, 0);
    attention_phase_rs(a, L, c.region_g);
    grid_barrier_rs<true>(a.bar, epoch, bd ? bd + 16 + 7 : nullptr); RS_STAMP();      // attn is read by TMA
    c.phase = 2;
    if (tid == 0) RS_DBG(2, 0);
    gemm_phase_rs<NTOK, W8, EPI_RESID, true>(a, L, c, make_geom<W8>(wmaps, a.n_layers, l, MAT_O), amaps, L.s_o, s_pos, s_rope, s_pick);
    load_gamma(gm, L.g2, a.B);
    grid_barrier_rs<false>(a.bar, epoch, bd ? bd + 32 + 7 : nullptr); RS_STAMP();
    c.phase = 3;
    if (tid == 0) RS_DBG(3, 0);
    build_norm_operand<NTOK>(a, c, gm, false);
    if (tid == 0) RS_DBG(3, 1);
    gemm_phase_rs<NTOK, W8, EPI_SWIGLU, false>(a, L, c, make_geom<W8>(wmaps, a.n_layers, l, MAT_GU), nullptr, L.s_gu, s_pos, s_rope, s_pick);
    grid_barrier_rs<true>(a.bar, epoch, bd ? bd + 48 + 7 : nullptr); RS_STAMP();      // act is read by TMA
    c.phase = 4;
    if (tid == 0) RS_DBG(4, 0);
    gemm_phase_rs<NTOK, W8, EPI_RESID, true>(a, L, c, make_geom<W8>(wmaps, a.n_layers, l, MAT_DOWN), amaps + 1, L.s_down, s_pos, s_rope, s_pick);
    load_gamma(gm, (l + 1 < a.n_layers) ? a.layers[l + 1].g1 : a.final_norm_bf, a.B);
    grid_barrier_rs<false>(a.bar, epoch, bd ? bd + 64 + 7 : nullptr); RS_STAMP();
  }
  c.layer = -1; c.phase = 5;
  // ---- lm_head with the argmax as epilogue
  if (tid < NTOK) { s_pick[tid * 3] = -INFINITY; s_pick[tid * 3 + 1] = -INFINITY; s_pick[tid * 3 + 2] = __int_as_float(0x7fffffff); }
  build_norm_operand<NTOK>(a, c, gm, false);
  {
    const RsLayer L0 = a.layers[0];
    gemm_phase_rs<NTOK, W8, EPI_HEAD, false>(a, L0, c, make_geom<W8>(wmaps, a.n_layers, 0, MAT_HEAD), nullptr, nullptr, s_pos, s_rope, s_pick);
    gemm_phase_rs<NTOK, W8, EPI_HEAD, false>(a, L0, c, make_geom<W8>(wmaps, a.n_layers, 0, MAT_HEAD_TAIL), nullptr, nullptr, s_pos, s_rope, s_pick);
  }
  if (warp >= 8) {
    sync_consumers();
    for (int t = tid - 256; t < a.B; t += 32 * kConsumers) {
      float* pp = a.pick_scratch + ((size_t)t * gridDim.x + blockIdx.x) * 4;
      pp[0] = s_pick[t * 3]; pp[1] = s_pick[t * 3 + 1]; pp[2] = s_pick[t * 3 + 2];
    }
  }
  grid_barrier_rs<false>(a.bar, epoch); RS_STAMP();
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    if (tid < 32) {
      float best = -INFINITY, second = -INFINITY;
      int bi = 0x7fffffff;
      for (int cta = tid; cta < (int)gridDim.x; cta += 32) {
        const float* pp = a.pick_scratch + ((size_t)b * gridDim.x + cta) * 4;
        pick_merge2(best, second, bi, __ldcg(pp), __ldcg(pp + 1), __float_as_int(__ldcg(pp + 2)));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        pick_merge2(best, second, bi, ob, os, oi);
      }
      if (tid == 0) {
        const int step = a.step;
        if (bi < 0 || bi >= RV) bi = a.gs.eos[0];             // all-NaN logits: emit EOS rather than an out-of-range id
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
  if (blockIdx.x == 0 && tid == 0) *a.gs.step = a.step + 1;      // the other decode paths read the step from device memory
  RS_STAMP();
}

typedef void (*RsKernel)(DecodeRsArgs);
struct RsVariant { RsKernel fn; int smem; int ntok; bool w8; };
const RsVariant kRsVariants[3] = {
    {decode_rs_kernel<16, false>, RsSmem<16, false>::kTotal, 16, false},
    {decode_rs_kernel<32, false>, RsSmem<32, false>::kTotal, 32, false},
    {decode_rs_kernel<16, true>, RsSmem<16, true>::kTotal, 16, true},
};
const RsVariant* rs_variant_for(bool w8, int B) {
  if (w8) return B <= 16 ? &kRsVariants[2] : nullptr;
  return B <= 16 ? &kRsVariants[0] : (B <= 32 ? &kRsVariants[1] : nullptr);
}

// rows of the fused qkv matrix: the q and k heads are re-ordered so that the RoPE pair (j, j + 64) of a head sits in rows
// (2j, 2j + 1); v rows keep their place.  dst row r <- src row decode_rs_qkv_src_row(r).
__host__ __device__ inline int qkv_src_row(int r) {
  if (r >= RQKV - RKVH * RHD) return r;
  const int hb = r >> 7, i = r & 127;
  return hb * RHD + (i >> 1) + 64 * (i & 1);
}
__global__ void permute_qkv_rows_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int row_bytes) {
  const int r = blockIdx.x;
  const uint4* s = reinterpret_cast<const uint4*>(src + (size_t)qkv_src_row(r) * row_bytes);
  uint4* d = reinterpret_cast<uint4*>(dst + (size_t)r * row_bytes);
  for (int i = threadIdx.x; i < row_bytes / 16; i += blockDim.x) d[i] = s[i];
}
__global__ void permute_qkv_scale_kernel(const float* __restrict__ src, float* __restrict__ dst) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < RQKV) dst[r] = src[qkv_src_row(r)];
}

}  // namespace

bool decode_rs_supports(bool w8, int B) { return B >= 1 && rs_variant_for(w8, B) != nullptr; }
int decode_rs_tokens(bool w8, int B) { const RsVariant* v = rs_variant_for(w8, B); return v ? v->ntok : 0; }
int decode_rs_box_rows(int kind) { return rs_box(kind); }

cudaError_t decode_rs_configure() {
  for (const RsVariant& v : kRsVariants) SONIC_CUDA_TRY(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, v.smem));
  return cudaSuccess;
}
int decode_rs_occupancy() {
  int worst = 1 << 30;
  for (const RsVariant& v : kRsVariants) {
    int per_sm = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v.fn, kRsThreads, v.smem);
    if (per_sm < worst) worst = per_sm;
  }
  return worst;
}

cudaError_t decode_rs_permute_qkv(const void* src, void* dst, int row_bytes, const float* scale_src, float* scale_dst, cudaStream_t st) {
  permute_qkv_rows_kernel<<<RQKV, 128, 0, st>>>(reinterpret_cast<const uint8_t*>(src), reinterpret_cast<uint8_t*>(dst), row_bytes);
  SONIC_LAUNCH_CHECK();
  if (scale_src) {
    permute_qkv_scale_kernel<<<(RQKV + 255) / 256, 256, 0, st>>>(scale_src, scale_dst);
    SONIC_LAUNCH_CHECK();
  }
  return cudaSuccess;
}

cudaError_t launch_decode_rs(const DecodeRsArgs& a, bool w8, int grid, cudaStream_t st, int* mode) {
  const RsVariant* v = rs_variant_for(w8, a.B);
  if (!v || grid < 1) return cudaErrorInvalidConfiguration;
  SONIC_CUDA_TRY(cudaMemsetAsync(a.bar, 0, sizeof(unsigned), st));
  cudaError_t e = cudaErrorUnknown;
  for (; *mode < 2; ++*mode) {
    if (*mode == 0) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kRsThreads); cfg.dynamicSmemBytes = v->smem; cfg.stream = st;
      cudaLaunchAttribute attr[1];
      memset(attr, 0, sizeof(attr));
      attr[0].id = cudaLaunchAttributeCooperative;
      attr[0].val.cooperative = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      e = cudaLaunchKernelEx(&cfg, v->fn, a);
    } else {
      DecodeRsArgs copy = a;
      void* args[1] = {&copy};
      e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(v->fn), dim3(grid), dim3(kRsThreads), args, v->smem, st);
    }
    if (e == cudaSuccess) return e;
    fprintf(stderr, "[sonicscribe_b200] decode_rs launch mode %d failed: %s (%s)\n", *mode, cudaGetErrorName(e), cudaGetErrorString(e));
    cudaGetLastError();
  }
  return e;
}

}  // namespace sonic

// ---- tcgen05.mma issue/execute rate microbenchmark (one CTA per SM, operands already in shared memory) ---------------------
// Measures the clocks per MMA for M in {64,128}, N = ntok, K = 16 (kind::f16, SWIZZLE_128B K-major operands), cycling through
// `n_acc` accumulators and `n_tiles` different A tiles; results feed the decode kernels' cost model.
namespace sonic {
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int m, int ntok, int n_mma, int n_acc, int n_tiles, unsigned long long* out) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (n_tiles * 16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<512>(smem_u32(&slot));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    unsigned long long t0 = 0, t1 = 0;
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc_bf16(m, ntok);
      const uint32_t bop = base + n_tiles * 16384;
      t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const uint64_t da = make_sw128_desc(base + (uint32_t)(i % n_tiles) * 16384) + (uint64_t)(2 * (i & 3));
        const uint64_t db = make_sw128_desc(bop) + (uint64_t)(2 * (i & 3));
        tc_mma_bf16(tmem + (uint32_t)(i % n_acc) * ntok, da, db, idesc, i >= n_acc ? 1u : 0u);
      }
      tc_commit(smem_u32(&bar));
      t1 = clock64();
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    const unsigned long long t2 = clock64();
    if (elect_one_sync() && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

cudaError_t bench_mma_rate(int m, int ntok, int n_mma, int n_acc, int n_tiles, float* issue_clk, float* total_clk, cudaStream_t st) {
  unsigned long long* d = nullptr;
  SONIC_CUDA_TRY(cudaMalloc(&d, 16));
  const int smem = 1024 + n_tiles * 16384 + 32768;
  cudaError_t e = cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) {
    mma_rate_kernel<<<1, 128, smem, st>>>(m, ntok, n_mma, n_acc, n_tiles, d);        // warm-up (instruction cache)
    mma_rate_kernel<<<1, 128, smem, st>>>(m, ntok, n_mma, n_acc, n_tiles, d);
    e = cudaStreamSynchronize(st);
  }
  unsigned long long h[2] = {0, 0};
  if (e == cudaSuccess) e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  *issue_clk = (float)h[0] / n_mma;
  *total_clk = (float)h[1] / n_mma;
  return e;
}
}  // namespace sonic
