// Internal launcher prototypes (device code lives in the .cu files next to this header).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>

#define SONIC_MEL_PEAK_NORM 1
#define SONIC_MEL_PCM16 2
#define SONIC_MEL_S16 4      // PCM buffer holds int16 samples (scaled by 1/32768 on load)

namespace sonic {

typedef __nv_bfloat16 bf16;

// ---- mel.cu ------------------------------------------------------------------------------------------------------
size_t mel_tables_bytes();
void mel_build_tables(void* host_out, const int* tap_start, const int* tap_count, const float* tapw);
cudaError_t mel_setup();
template <typename T>
cudaError_t launch_mel(const float* pcm, const long long* offs, const int* lens, int batch, int max_len, int flags,
                       const void* tables, unsigned* peak_bits, unsigned* gmax_bits, float* tile_min /*[batch][mel_tiles_per_segment()]*/,
                       float* feat, T* feat_tm, cudaStream_t st);
int mel_tiles_per_segment();

// ---- gemm: C[b][M,N] = epilogue(A[b][M,K] * W[N,K]^T) ---------------------------------------------------------------
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_SWIGLU = 2 };   // SWIGLU: columns are (gate,up) pairs; writes N/2 columns
struct GemmArgs {
  const void* A; long long lda; long long a_bstride;       // elements
  const void* W; long long ldw;                             // [N,K] row-major ("K-major")
  void* C; long long ldc; long long c_bstride; long long c_row0;   // output row offset inside each batch slab
  const float* bias;                                        // [N] or null (fp32)
  const void* resid; long long ldr; long long r_bstride;    // same dtype as C; may alias C
  int M, N, K, batch;
  int act;
  int out_f32;                                              // 1: C is float regardless of the activation dtype
  // implicit conv1d(k=3, pad=1, stride 1|2) over a zero-padded time-major input [batch][conv_rows_pad, conv_cin]:
  // A points at padded row 0, K = 3*conv_cin, lda = conv_stride*conv_cin (the overlapping-row view the SIMT kernel uses
  // directly; the TMA kernel walks the three taps with a non-overlapping tensor map).  conv_cin = 0 => plain GEMM.
  int conv_cin, conv_stride, conv_rows_pad;
  // split-K scratch of the tcgen05 decode orientation: fp32 partials + per-tile arrival counters (zero-initialised,
  // >= 2048 ints).  Null => no split.  One scratch per stream: launches that share it must be stream-ordered.
  float* splitk_ws; size_t splitk_ws_bytes; int* splitk_counters;
  int w_int8; const float* wscale;   // tcgen05 path: W is int8 [N,K] with per-row fp32 scale (weight-only quantisation)
  const float* rope_cos; const float* rope_sin; int rope_T; int rope_ncols;   // tcgen05 normal mode: fused encoder RoPE (see gemm_tc.cu)
  int pdl;   // launch with programmatic stream serialization (weights are prefetched before the dependency wait)
};
template <typename T> cudaError_t launch_gemm_simt(const GemmArgs& g, cudaStream_t st);       // gemm_simt.cu

// ---- ops.cu ------------------------------------------------------------------------------------------------------
template <typename T>
cudaError_t launch_layernorm(const T* x, T* y, const float* gamma, const float* beta, int rows, int H, float eps, cudaStream_t st);
template <typename T>
cudaError_t launch_rmsnorm(const T* x, T* y, const float* gamma, int rows, int H, float eps, cudaStream_t st, bool pdl = false);
// rows_idx: optional gather of input rows (used for the last-position final norm)
template <typename T>
cudaError_t launch_rmsnorm_rows(const T* x, const int* rows_idx, T* y, const float* gamma, int rows, int H, float eps, cudaStream_t st, bool pdl = false);

// encoder RoPE on the fused QKV buffer [rows, 3*H]: rotate the first rot dims of every q and k head. pos = row % T.
template <typename T>
cudaError_t launch_rope_enc(T* qkv, const float* cos_t, const float* sin_t, int rows, int T_len, int heads, int hd, int rot, cudaStream_t st);
// decoder RoPE + KV append. qkv [rows, (H + 2*KV)*hd]; row r belongs to segment row_seg[r] at position row_pos[r]
// (null row_seg => decode step: segment = r, position = ctx_len[r]). q rotated in place; k (rotated) and v stored into the caches.
template <typename T>
cudaError_t launch_rope_dec_kv(T* qkv, const float* cos_t, const float* sin_t, const int* row_seg, const int* row_pos,
                               const int* ctx_len, T* kcache, T* vcache, int rows, int heads, int kv_heads, int hd,
                               int max_ctx, cudaStream_t st);
template <typename T>
cudaError_t launch_embed(const int* ids, const int* audio_src, const T* table, const T* audio_embeds, T* x, int rows, int H, cudaStream_t st, bool pdl = false);
template <typename T>
cudaError_t launch_embed_next(const int* cur_tok, const T* table, T* x, int rows, int H, cudaStream_t st, bool pdl = false);
// greedy step bookkeeping: argmax(+top-2 margin) over logits [B,V]; appends to out_ids unless finished; updates state.
struct GreedyState {
  int* cur_tok;      // [B] next input token
  int* ctx_len;      // [B] tokens in the KV cache
  int* finished;     // [B]
  int* n_out;        // [B]
  int* out_ids;      // [B][max_new]
  float* margins;    // [B][max_new]
  int* step;         // [1]
  int* n_unfinished; // [1]
  int* step_arrivals;   // [1] zero-initialised
  void* pick_partials;  // greedy_pick_scratch_bytes(max_batch)
  int* pick_counters;   // [max_batch + 1] zero-initialised
  int max_new;
  int eos[4];
  int n_eos;
};
size_t greedy_pick_scratch_bytes(int max_batch);
cudaError_t launch_greedy_pick(const float* logits, int B, int V, GreedyState gs, int advance_ctx, cudaStream_t st, bool pdl = false);
void rope_table_host(float* cos_t, float* sin_t, int positions, int rot_dim, float theta);

// ---- attention.cu ---------------------------------------------------------------------------------------------
struct AttnArgs {
  const void* q; long long q_row_stride;                 // q[(row)*q_row_stride + h*hd + d]
  const void* k; long long k_tok_stride, k_head_stride, k_seg_stride;
  const void* v; long long v_tok_stride, v_head_stride, v_seg_stride;
  void* o; long long o_row_stride;
  const int* q_off;     // [B+1] first q row of each segment (null => b*q_len_fixed)
  const int* kv_len;    // [B] keys visible per segment (null => kv_len_fixed)
  int q_len_fixed, kv_len_fixed;
  int causal;           // 1: key j visible to the i-th query of a segment iff j <= (kv_len - q_len + i)
  int decode;           // 1: one query per segment (row b), kv_len read AFTER the append (ctx_len+1 handled by caller)
  int heads, kv_heads, hd, batch, max_q;
  float scale;
};
template <typename T> cudaError_t launch_attention_simt(const AttnArgs& a, cudaStream_t st);

// ---- decode_attn.cu: fused RoPE + KV append + split-KV attention + combine for one new token per segment (bf16) -----
struct DecodeAttnArgs {
  const bf16* qkv;            // [segments, (heads + 2*kv_heads) * 128] un-rotated q | k | v of the new token
  const float* cos_t; const float* sin_t;   // [max_ctx, 64]
  const int* ctx_len;         // [segments] tokens already in the cache == position of the new token
  bf16* kcache; bf16* vcache; // this layer: [segments(max_batch stride)][kv_heads][max_ctx][128]
  bf16* out;                  // [segments, heads*128]
  float* ws; int* counters;   // partials [segments*kv_heads][max_chunks][4][130], arrival counters [segments*kv_heads]
  int kv_heads, max_ctx, max_chunks;
  float scale;
};
cudaError_t launch_decode_attn(const DecodeAttnArgs& a, int batch, int n_chunks, cudaStream_t st, bool pdl = false);

// ---- decode_persist.cu: one cooperative kernel per greedy step (all layers + lm_head + pick), bf16 ------------------------
struct DecLayerDev {
  const void *wqkv, *wo, *wgu, *wdown;            // bf16 [N][K], or int8 [N][K] when DecodePersistArgs::w8
  const float *sqkv, *so, *sgu, *sdown;           // int8 mode: per-row scales
  const float *rms1, *rms2; bf16 *kc, *vc;
};
struct DecodePersistArgs {
  const DecLayerDev* layers; int n_layers;
  const bf16* embed; const bf16* lm_head; const float* final_norm;
  const float* cos_t; const float* sin_t;
  bf16 *x, *u, *attn, *act;       // [B][2048], [B][2048], [B][2048], [B][6144]
  float* part;                    // split-K partials, decode_persist_part_floats(Bpad) floats
  float* logits_out;              // optional [B][vocab]
  float* pick_scratch;            // decode_persist_pick_floats(max_batch, num_sms)
  float* attn_ws; int* attn_counters; int attn_chunks;   // split-KV partials [B*4][attn_chunks][4][130], counters [B*4] (zeroed)
  int attn_chunk_keys;            // keys per split-KV item when attn_chunks > 1: 64 or 128
  GreedyState gs;
  unsigned* bar;                  // grid barrier counter
  unsigned long long* timestamps; // optional: %globaltimer after every grid barrier (CTA 0), 1 + 8*layers + 3 entries
  int B, Bpad, max_ctx;
  int w8;                         // 1: decoder linears are int8 weight-only (lm_head / embedding stay bf16)
  // batch class 33..64, bf16 weights: the GEMM phases run on tcgen05 fed by TMA.  Device array of CUtensorMap (128 B each):
  // [4*l + {0: qkv, 1: o, 2: gate/up, 3: down}] weight maps (box 64 k x 128 rows; gate/up: x kPersistGuTileRows), [4*n_layers] lm_head,
  // [4*n_layers + 1 + 3*w + {0: u, 1: attn, 2: act}] activation maps (box 64 k x {64, 32, 16, 128, 256}[w] token rows).  nullptr: mma.sync phases.
  const void* tmaps;
  // attention phase: device array of two CUtensorMap over the whole K and V caches ({128 dims, layers*batch*4*max_ctx key rows},
  // box 64 dims x 64 keys, 128B swizzle); kc_base = first element of the K cache (row 0 of the maps)
  const void* kv_maps; const bf16* kc_base;
  int tc_ntok;                    // tcgen05 classes: token rows per activation tile = MMA N: 16, 32, 64, 128 or 256, >= B
  int tc_pre_depth;               // tcgen05 classes: ring stages of the NEXT phase's weights put in flight before each grid barrier
  int dbg_flags;                  // timing experiments (SONIC_PERSIST_DBGFLAGS; results are wrong with them): 1 / 2 skip the proxy fence before / after the gate/up barrier, 4 skip the SwiGLU stores
  int dbg_cta;                    // debug handles: CTA whose gate/up phase of layer 1 writes fine-grained stamps to timestamps + 1024 (-1: none)
  float eps, scale;
};
// weight rows per gate/up item of the tcgen05 decode classes = box height of the gate/up weight map: 147 items of 84 rows cover
// the 12288 interleaved (gate, up) rows, one whole-K item per CTA of a 148-SM grid (128-row tiles would occupy 96 CTAs)
static constexpr int kPersistGuTileRows = 84;
static constexpr int kPersistQkvTileRows = 128, kPersistOTileRows = 128, kPersistDownTileRows = 128;   // x K splits 6 / 9 / 9 (decode_persist.cu)
static constexpr int kPersistLmTileRows = 104;  // lm_head: 570 items = 3.85 per CTA
size_t decode_persist_smem_bytes();
size_t decode_persist_part_floats(int Bpad);
size_t decode_persist_pick_floats(int max_batch, int num_sms);
cudaError_t decode_persist_configure();
int decode_persist_max_grid(int num_sms);
int decode_persist_occupancy();
// *mode: per-handle launch API state (0 on first use); advanced when a launch API is refused
cudaError_t launch_decode_persist(const DecodePersistArgs& a, int grid, cudaStream_t st, int* mode);
static constexpr int kPersistTcTokens = 64;   // token rows of the widest activation tensor maps of the row-sliced kernel (decode_rs.cu)
static constexpr int kPersistTcMaxTokens = 256;   // largest live batch of the tcgen05 decode classes (activation tiles of 16 ... 256 token rows; int8: 128)

// ---- decode_rs.cu: row-sliced persistent decode step for <= 32 segments (bf16) / <= 16 segments (int8 weight-only) ---------
struct RsLayer {
  const bf16 *g1, *g2;                            // input / post-attention RMSNorm weights as bf16 (the fp32 vectors hold bf16-rounded values)
  const float *s_qkv, *s_o, *s_gu, *s_down;       // int8 mode: per-row scales (s_qkv in the interleaved row order of wqkv_il)
  bf16 *kc, *vc;                                  // this layer's K / V cache [max_batch][4][max_ctx][128]
};
struct DecodeRsArgs {
  const RsLayer* layers; int n_layers;
  const bf16* embed; const bf16* final_norm_bf;
  const float* cos_t; const float* sin_t;
  bf16 *x, *q, *attn, *act;       // residual stream [B][2048], rotated queries [B][2048], attention output [B][2048], SwiGLU output [B][6144]
  // device array of CUtensorMap: [4*l + {0: qkv (q/k rows interleaved), 1: o, 2: gate/up, 3: down}] with box rows decode_rs_box_rows(kind),
  // [4*n_layers] lm_head (128-row boxes), [4*n_layers + 1] lm_head (32-row boxes).  bf16 maps use 128B swizzle; int8 layer maps none.
  const void* wmaps;
  const void* amaps;              // CUtensorMap[2]: attn [64 rows][2048] and act [64 rows][6144], box 64 k x (16 | 32) token rows
  float* pick_scratch;            // [max_batch][grid][4]
  float* logits_out;              // optional [B][vocab]
  float* attn_ws; int* attn_counters; int attn_chunks;   // split-KV partials [B*4][attn_chunks][4][130]; counters [B*4] (zeroed); chunks of 64 keys
  GreedyState gs;
  unsigned* bar;
  unsigned long long* timestamps; // optional: %globaltimer of CTA 0 after every grid barrier (5 * layers + 3 entries)
  int l2_prefetch;                // 1: the producer prefetches the coming weight boxes into L2 ahead of the ring
  unsigned long long* dbg; int dbg_cta, dbg_layer;   // optional: 80 fine-grained stamps of one layer on one CTA (decode_rs.cu RS_DBG)
  int B, max_ctx, step;           // step: index of the token this launch produces
  float eps, scale;
};
bool decode_rs_supports(bool w8, int B);
int decode_rs_tokens(bool w8, int B);              // token rows of the activation boxes the class for this batch uses (16 | 32)
int decode_rs_box_rows(int kind);                  // kind: 0 qkv, 1 o, 2 gate/up, 3 down, 4 lm_head, 5 lm_head tail
cudaError_t decode_rs_configure();
int decode_rs_occupancy();
// dst <- fused qkv matrix with the q / k head rows interleaved for the RoPE epilogue (and the int8 row scales likewise)
cudaError_t decode_rs_permute_qkv(const void* src, void* dst, int row_bytes, const float* scale_src, float* scale_dst, cudaStream_t st);
cudaError_t launch_decode_rs(const DecodeRsArgs& a, bool w8, int grid, cudaStream_t st, int* mode);
// clocks per tcgen05.mma (M x ntok x 16, operands in shared memory): issue time of the loop and time to completion
cudaError_t bench_mma_rate(int m, int ntok, int n_mma, int n_acc, int n_tiles, float* issue_clk, float* total_clk, cudaStream_t st);

}  // namespace sonic
