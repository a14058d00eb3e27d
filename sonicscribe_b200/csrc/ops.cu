// Row-wise and bookkeeping kernels of the GLM-ASR path: LayerNorm, RMSNorm, RoPE (+KV append), embedding gather with
// audio-embedding scatter, greedy pick.  All math in fp32; T is the activation storage type (float or bf16).
#include "common.cuh"
#include "kernels.h"
#include <math.h>

namespace sonic {

// ---- vectorised row access: 8 contiguous elements per thread (16 B for bf16, 32 B for fp32) -------------------------------
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  const uint4 a = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 o;
  __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
  o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
  o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
  *reinterpret_cast<uint4*>(p) = o;
}

// ---- LayerNorm (modeling_glmasr.py:250-251,308; eps 1e-5, affine).  One row per CTA, H/8 <= 256 active threads ------------
template <typename T>
__global__ void __launch_bounds__(256) layernorm_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int H, float eps) {
  __shared__ float red[32];
  const size_t row = blockIdx.x;
  const int c0 = threadIdx.x * 8;
  const bool on = c0 < H;
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (on) load8(x + row * H + c0, v);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mean = block_sum(s, red) / H;
  float q = 0.f;
  if (on) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q += d * d; }
  }
  const float rstd = rsqrtf(block_sum(q, red) / H + eps);
  if (on) {
    float g[8], b[8], o[8];
    load8(gamma + c0, g);
    load8(beta + c0, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = (v[i] - mean) * rstd * g[i] + b[i];
    store8(y + row * H + c0, o);
  }
}
// One row per WARP (8 rows per CTA): the whole row lives in registers (CH chunks of 8 per lane), both reductions are warp
// shuffles, no shared memory and no CTA barrier.  96 000 encoder rows per launch at 64 segments: the one-row-per-CTA kernel
// above spent its time in two block reductions per row (2.8 TB/s); this one streams.
template <typename T, int CH>
__global__ void __launch_bounds__(256) layernorm_warp_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, int rows, int H, float eps) {
  const int lane = threadIdx.x & 31;
  const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (size_t)rows) return;
  float v[CH][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int c0 = (lane + 32 * c) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[c][i] = 0.f;
    if (c0 < H) load8(x + row * H + c0, v[c]);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[c][i];
  }
  const float mean = warp_sum(s) / H;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    if ((lane + 32 * c) * 8 < H) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = v[c][i] - mean; q += d * d; }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / H + eps);
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int c0 = (lane + 32 * c) * 8;
    if (c0 < H) {
      float g[8], b[8], o[8];
      load8(gamma + c0, g);
      load8(beta + c0, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (v[c][i] - mean) * rstd * g[i] + b[i];
      store8(y + row * H + c0, o);
    }
  }
}
template <typename T>
cudaError_t launch_layernorm(const T* x, T* y, const float* gamma, const float* beta, int rows, int H, float eps, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  if (H > 2048 || H % 8 != 0) return cudaErrorInvalidValue;
  const int grid = (rows + 7) / 8;
  if (H <= 1280) layernorm_warp_kernel<T, 5><<<grid, 256, 0, st>>>(x, y, gamma, beta, rows, H, eps);
  else layernorm_warp_kernel<T, 8><<<grid, 256, 0, st>>>(x, y, gamma, beta, rows, H, eps);
  return cudaGetLastError();
}

// ---- RMSNorm (modeling_llama.py:62-67): w * (x_f32 * rsqrt(mean(x_f32^2)+eps)).to(dtype) -------------------------------
template <typename T>
__global__ void __launch_bounds__(256) rmsnorm_kernel(const T* __restrict__ x, const int* __restrict__ rows_idx, T* __restrict__ y,
                                                      const float* __restrict__ gamma, int H, float eps) {
  __shared__ float red[32];
  pdl_launch_dependents();
  pdl_wait();
  const size_t row = blockIdx.x;
  const size_t src = rows_idx ? (size_t)rows_idx[row] : row;
  const int c0 = threadIdx.x * 8;
  const bool on = c0 < H;
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (on) load8(x + src * H + c0, v);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i] * v[i];
  const float rstd = rsqrtf(block_sum(s, red) / H + eps);
  if (on) {
    float g[8], o[8];
    load8(gamma + c0, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = g[i] * to_f32(from_f32<T>(v[i] * rstd));   // the reference rounds to the model dtype before the weight
    store8(y + row * H + c0, o);
  }
}
template <typename T>
cudaError_t launch_rmsnorm_rows(const T* x, const int* rows_idx, T* y, const float* gamma, int rows, int H, float eps, cudaStream_t st, bool pdl) {
  if (rows <= 0) return cudaSuccess;
  if (H > 2048 || H % 8 != 0) return cudaErrorInvalidValue;
  return launch_ex(rmsnorm_kernel<T>, dim3(rows), dim3(256), 0, st, pdl, x, rows_idx, y, gamma, H, eps);
}
template <typename T>
cudaError_t launch_rmsnorm(const T* x, T* y, const float* gamma, int rows, int H, float eps, cudaStream_t st, bool pdl) {
  return launch_rmsnorm_rows<T>(x, nullptr, y, gamma, rows, H, eps, st, pdl);
}

// ---- RoPE tables (modeling_glmasr.py:66-109 / modeling_llama.py:96-136): fp32, angle = pos * inv_freq -------------------
void rope_table_host(float* cos_t, float* sin_t, int positions, int rot_dim, float theta) {
  const int half = rot_dim / 2;
  for (int j = 0; j < half; ++j) {
    const float e = (float)(2 * j) / (float)rot_dim;
    const float inv = 1.0f / (float)pow((double)theta, (double)e);
    for (int p = 0; p < positions; ++p) {
      const float a = (float)p * inv;
      cos_t[(size_t)p * half + j] = (float)cos((double)a);
      sin_t[(size_t)p * half + j] = (float)sin((double)a);
    }
  }
}

// encoder: rotate first `rot` dims of each q/k head in the fused [rows, 3*heads*hd] buffer (NeoX halves: j <-> j+rot/2)
template <typename T>
__global__ void rope_enc_kernel(T* __restrict__ qkv, const float* __restrict__ cos_t, const float* __restrict__ sin_t, int rows,
                                int T_len, int heads, int hd, int rot) {
  const int half = rot / 2;
  const int per_row = 2 * heads * half;                 // q and k
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * per_row) return;
  const int row = (int)(idx / per_row);
  int r = (int)(idx - (long long)row * per_row);
  const int which = r / (heads * half);                 // 0 q, 1 k
  r -= which * heads * half;
  const int h = r / half, j = r - h * half;
  const int pos = row % T_len;
  // the reference casts cos/sin to the model dtype (modeling_glmasr.py:109)
  const float c = to_f32(from_f32<T>(cos_t[pos * half + j])), s = to_f32(from_f32<T>(sin_t[pos * half + j]));
  T* p = qkv + (size_t)row * (3 * heads * hd) + (size_t)which * heads * hd + h * hd;
  const float a = to_f32(p[j]), b = to_f32(p[j + half]);
  p[j] = from_f32<T>(a * c - b * s);
  p[j + half] = from_f32<T>(b * c + a * s);
}
template <typename T>
cudaError_t launch_rope_enc(T* qkv, const float* cos_t, const float* sin_t, int rows, int T_len, int heads, int hd, int rot, cudaStream_t st) {
  const long long total = (long long)rows * 2 * heads * (rot / 2);
  if (total <= 0) return cudaSuccess;
  rope_enc_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(qkv, cos_t, sin_t, rows, T_len, heads, hd, rot);
  return cudaGetLastError();
}

// decoder: full-dim RoPE on q (in place) and k; k,v appended to the caches [seg][kv_head][pos][hd]
template <typename T>
__global__ void rope_dec_kv_kernel(T* __restrict__ qkv, const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                   const int* __restrict__ row_seg, const int* __restrict__ row_pos, const int* __restrict__ ctx_len,
                                   T* __restrict__ kcache, T* __restrict__ vcache, int heads, int kv_heads, int hd, int max_ctx) {
  const int row = blockIdx.x;
  const int seg = row_seg ? row_seg[row] : row;
  const int pos = row_seg ? row_pos[row] : ctx_len[row];
  const int half = hd / 2;
  const int width = (heads + 2 * kv_heads) * hd;
  T* base = qkv + (size_t)row * width;
  const int n_rot = (heads + kv_heads) * half;
  for (int i = threadIdx.x; i < n_rot; i += blockDim.x) {
    const int h = i / half, j = i - h * half;              // h < heads: q head, else k head
    const float c = to_f32(from_f32<T>(cos_t[(size_t)pos * half + j])), s = to_f32(from_f32<T>(sin_t[(size_t)pos * half + j]));
    T* p = base + (size_t)h * hd;
    const float a = to_f32(p[j]), b = to_f32(p[j + half]);
    const T ra = from_f32<T>(a * c - b * s), rb = from_f32<T>(b * c + a * s);
    if (h < heads) {
      p[j] = ra;
      p[j + half] = rb;
    } else {
      T* kc = kcache + (((size_t)seg * kv_heads + (h - heads)) * max_ctx + pos) * hd;
      kc[j] = ra;
      kc[j + half] = rb;
    }
  }
  const T* vsrc = base + (size_t)(heads + kv_heads) * hd;
  for (int i = threadIdx.x; i < kv_heads * hd; i += blockDim.x) {
    const int h = i / hd, d = i - h * hd;
    vcache[(((size_t)seg * kv_heads + h) * max_ctx + pos) * hd + d] = vsrc[i];
  }
}
template <typename T>
cudaError_t launch_rope_dec_kv(T* qkv, const float* cos_t, const float* sin_t, const int* row_seg, const int* row_pos,
                               const int* ctx_len, T* kcache, T* vcache, int rows, int heads, int kv_heads, int hd,
                               int max_ctx, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  rope_dec_kv_kernel<T><<<rows, 256, 0, st>>>(qkv, cos_t, sin_t, row_seg, row_pos, ctx_len, kcache, vcache, heads, kv_heads, hd, max_ctx);
  return cudaGetLastError();
}

// ---- embedding gather + audio scatter (modeling_glmasr.py:473-483) ------------------------------------------------------
template <typename T>
__global__ void embed_kernel(const int* __restrict__ ids, const int* __restrict__ audio_src, const T* __restrict__ table,
                             const T* __restrict__ audio, T* __restrict__ x, int H) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;
  const int a = audio_src ? audio_src[row] : -1;
  const T* src = (a >= 0) ? audio + (size_t)a * H : table + (size_t)ids[row] * H;
  for (int i = threadIdx.x; i < H; i += blockDim.x) x[(size_t)row * H + i] = src[i];
}
template <typename T>
cudaError_t launch_embed(const int* ids, const int* audio_src, const T* table, const T* audio_embeds, T* x, int rows, int H, cudaStream_t st, bool pdl) {
  if (rows <= 0) return cudaSuccess;
  return launch_ex(embed_kernel<T>, dim3(rows), dim3(256), 0, st, pdl, ids, audio_src, table, audio_embeds, x, H);
}
template <typename T>
cudaError_t launch_embed_next(const int* cur_tok, const T* table, T* x, int rows, int H, cudaStream_t st, bool pdl) {
  return launch_embed<T>(cur_tok, nullptr, table, nullptr, x, rows, H, st, pdl);
}

// ---- greedy pick (generation/utils.py:2793-2805): argmax of fp32 logits, first index on ties; EOS bookkeeping ---------
// grid (kPickSlices, B): every CTA scans one slice of the vocabulary; the CTA arriving last at the segment's counter merges
// the slice results in slice order and does the bookkeeping (also advances the step counter once per launch).
static constexpr int kPickSlices = 16;
struct PickPartial { float best, second; int idx; int pad; };

__device__ __forceinline__ void pick_merge(float& mb, float& ms, int& mi, float ob, float os, int oi) {
  if (ob > mb || (ob == mb && oi < mi)) { ms = fmaxf(fmaxf(ms, os), mb); mb = ob; mi = oi; }
  else { ms = fmaxf(ms, ob); }
}

__global__ void __launch_bounds__(256) greedy_pick_kernel(const float* __restrict__ logits, int V, GreedyState gs, int advance_ctx,
                                                          PickPartial* __restrict__ partials, int* __restrict__ counters) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y, slice = blockIdx.x;
  const int per = (V + kPickSlices - 1) / kPickSlices;
  const int lo = slice * per, hi = min(V, lo + per);
  const float* l = logits + (size_t)b * V;
  float best = -INFINITY, second = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float v = l[i];
    if (v > best) { second = best; best = v; bi = i; }
    else if (v > second) second = v;
  }
  // warp then block merge
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    pick_merge(best, second, bi, ob, os, oi);
  }
  __shared__ float s_best[8], s_second[8];
  __shared__ int s_idx[8], s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_best[warp] = best; s_second[warp] = second; s_idx[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) pick_merge(best, second, bi, s_best[w], s_second[w], s_idx[w]);
    PickPartial p; p.best = best; p.second = second; p.idx = bi; p.pad = 0;
    partials[b * kPickSlices + slice] = p;
    __threadfence();
    const int prev = atomicAdd(counters + b, 1);
    s_last = (prev == kPickSlices - 1);
    if (s_last) counters[b] = 0;
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  float mb = -INFINITY, ms = -INFINITY;
  int mi = 0x7fffffff;
  for (int s = 0; s < kPickSlices; ++s) {
    const volatile PickPartial* p = partials + b * kPickSlices + s;
    pick_merge(mb, ms, mi, p->best, p->second, p->idx);
  }
  // a row of NaN logits (NaN PCM, poisoned cache) never updates the running maximum: emit EOS instead of an out-of-range id
  const bool bad = (mi < 0 || mi >= V);
  const int tok = bad ? gs.eos[0] : mi;
  const int step = *gs.step;
  if (advance_ctx) gs.ctx_len[b] += 1;                 // the token just consumed is now in the cache
  if (!gs.finished[b]) {
    gs.out_ids[(size_t)b * gs.max_new + step] = tok;
    if (gs.margins) gs.margins[(size_t)b * gs.max_new + step] = mb - ms;
    gs.n_out[b] = step + 1;
    bool eos = false;
    for (int e = 0; e < gs.n_eos; ++e) eos |= (tok == gs.eos[e]);
    if (eos || step + 1 >= gs.max_new) { gs.finished[b] = 1; atomicSub(gs.n_unfinished, 1); }
  }
  gs.cur_tok[b] = tok;
  // the segment finishing last in this launch advances the shared step counter
  __threadfence();
  const int arrived = atomicAdd(gs.step_arrivals, 1);
  if (arrived == (int)gridDim.y - 1) { *gs.step_arrivals = 0; *gs.step = step + 1; }
}

cudaError_t launch_greedy_pick(const float* logits, int B, int V, GreedyState gs, int advance_ctx, cudaStream_t st, bool pdl) {
  if (B <= 0) return cudaSuccess;
  return launch_ex(greedy_pick_kernel, dim3(kPickSlices, B), dim3(256), 0, st, pdl, logits, V, gs, advance_ctx,
                   reinterpret_cast<PickPartial*>(gs.pick_partials), gs.pick_counters);
}
size_t greedy_pick_scratch_bytes(int max_batch) { return (size_t)max_batch * kPickSlices * sizeof(PickPartial); }

#define INST(T)                                                                                                           \
  template cudaError_t launch_layernorm<T>(const T*, T*, const float*, const float*, int, int, float, cudaStream_t);       \
  template cudaError_t launch_rmsnorm<T>(const T*, T*, const float*, int, int, float, cudaStream_t, bool);                        \
  template cudaError_t launch_rmsnorm_rows<T>(const T*, const int*, T*, const float*, int, int, float, cudaStream_t, bool);       \
  template cudaError_t launch_rope_enc<T>(T*, const float*, const float*, int, int, int, int, int, cudaStream_t);           \
  template cudaError_t launch_rope_dec_kv<T>(T*, const float*, const float*, const int*, const int*, const int*, T*, T*,    \
                                             int, int, int, int, int, cudaStream_t);                                       \
  template cudaError_t launch_embed<T>(const int*, const int*, const T*, const T*, T*, int, int, cudaStream_t, bool);             \
  template cudaError_t launch_embed_next<T>(const int*, const T*, T*, int, int, cudaStream_t, bool);
INST(float)
INST(bf16)

}  // namespace sonic
