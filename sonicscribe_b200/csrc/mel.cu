// 16 kHz log-mel front end for sm_100a.
//
// Replaces (reference call chain): backend/asr.py:248-276 (peak-normalise + PCM_16 WAV round trip) and
// transformers/models/whisper/feature_extraction_whisper.py:135-164,296-337 (pad to 30 s, reflect-centred STFT
// n_fft=400 hop=160 periodic Hann, |.|^2, 128-bin slaney mel, log10 clamp, max-8 floor, (x+4)/4).
//
// Kernels
//   mel_peak_kernel     per-segment max|x| (the peak-normalise reduction).  Launched per group of <= 64 segments right before
//                       that group's frames kernel, so the second read of the PCM comes from the 126 MB L2, not from HBM.
//   mel_frames_kernel   persistent CTAs over 32-frame tiles.  Per tile:
//                         load    the 5360-sample span: float4 loads issued during the PREVIOUS tile's mel phase (registers),
//                                 pre-step (exact division by two FMA corrections of v * rcp(peak), PCM16 rint) -> smem
//                         step 1  frame PAIRS as one complex transform (z = w fA + i w fB), 400 = 16 x 25 Cooley-Tukey:
//                                 16-point DFTs in registers (packed fp32x2 butterflies), twiddled, float2 rows to smem
//                         step 2  25-point DFTs in registers; the two real spectra are separated WITHOUT a pass over smem:
//                                 bin k of a pair lives in thread (k mod 16), bin 400 - k in thread (16 - k mod 16) of the same
//                                 half-warp, so one shuffle per value fetches the partner and every thread writes one power
//                                 (frame A for k <= 200, frame B for the mirrored bins) in natural bin order
//                         mel     lane = frame, warp strides over mel bins: taps in float4 groups, log10 via lg2, (x+4)/4
//                                 stored straight to the fp32 [B,128,3000] features (coalesced 128 B rows) and / or staged for
//                                 the time-major encoder input [B,3002,128]; per-tile minimum, per-segment maximum on the side
//   mel_fixup_kernel    the max(x, gmax-8) clamp needs the segment maximum, known only after the last frame: this pass
//                       fills the frames that only see zero padding (never transformed: exactly log10(1e-10) = -10) and
//                       re-touches only the 32-frame tiles whose minimum is below gmax-8 (silence), instead of streaming a
//                       raw fp32 copy out and back in (round 1: 7.2 MB of traffic per segment against 2.8 MB algorithmic).
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"
#include "kernels.h"

namespace sonic {

static constexpr int kNfft = 400, kHop = 160, kBins = 201, kMels = 128, kFrames = 3000, kWin = 480000;
static constexpr int kPairs = 16;                 // complex transforms per CTA tile
static constexpr int kTileFrames = 2 * kPairs;    // 32 frames per tile
static constexpr int kSpan = (kTileFrames - 1) * kHop + kNfft;   // 5360 samples
static constexpr int kPStride = kBins;            // power rows [frame][bin]: odd stride, lanes (frames) hit distinct banks
static constexpr int kPFloats = kTileFrames * kPStride + 16;     // + slack for the zero-weight taps of the last float4 group
static constexpr int kMelThreads = 256;
static constexpr int kMaxTaps = 12;
static constexpr int kPer4 = (kSpan / 4 + kMelThreads - 1) / kMelThreads;    // 1340 float4 -> 6 per thread
static constexpr int kPer1 = (kSpan + kMelThreads - 1) / kMelThreads;        // 21 scalars per thread (edge tiles)

#include "mel_twiddles.inc"   // __constant__ float2 c_w16[16], c_w25[25]  (e^{-2 pi i k/n})

struct MelTables {             // built once on the host (api.cu) from the slaney filter bank, uploaded to global
  float window[kNfft];
  float2 tw[16 * 25];        // step-1 twiddles W400^(n2*k1) laid out [k1][n2]: consecutive lanes (n2) read consecutive entries
  float tapw[kMels * kMaxTaps];
  int tap_start[kMels];
  int tap_count[kMels];
  int meta[kMels];           // mel bins ordered by their number of float4 tap groups (1, 2, 3): bin | first tap << 8
  int cls_end[4];            // [c]: entries of `meta` with at most c + 1 groups
};

// ---- packed fp32x2 arithmetic (one issue slot per complex add / scaled add) ----------------------------------------------
__device__ __forceinline__ float2 padd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 psub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 pmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 cswap(float2 a) { return make_float2(a.y, a.x); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// forward 4-point DFT in place: (a0..a3) -> (X0..X3)
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 t0 = padd(a0, a2), t1 = psub(a0, a2), t2 = padd(a1, a3), t3 = psub(a1, a3);
  const float2 sw = cswap(t3);                                    // -i t3 = (t3.y, -t3.x)
  a0 = padd(t0, t2);
  a2 = psub(t0, t2);
  a1 = pfma(sw, make_float2(1.f, -1.f), t1);
  a3 = pfma(sw, make_float2(-1.f, 1.f), t1);
}
// forward 5-point DFT in place
__device__ __forceinline__ void dft5(float2& a0, float2& a1, float2& a2, float2& a3, float2& a4) {
  const float c1 = 0.30901699437494745f, c2 = -0.8090169943749473f, s1 = 0.9510565162951535f, s2 = 0.5877852522924732f;
  const float2 C1 = make_float2(c1, c1), C2 = make_float2(c2, c2), S1 = make_float2(s1, s1), S2 = make_float2(s2, s2), NS1 = make_float2(-s1, -s1);
  const float2 p = padd(a1, a4), q = padd(a2, a3), d1 = psub(a1, a4), d2 = psub(a2, a3);
  const float2 x0 = padd(padd(a0, p), q);
  const float2 p1 = pfma(q, C2, pfma(p, C1, a0));
  const float2 p2 = pfma(q, C1, pfma(p, C2, a0));
  const float2 q1 = cswap(pfma(d2, S2, pmul(d1, S1)));           // (q1.y, q1.x)
  const float2 q2 = cswap(pfma(d2, NS1, pmul(d1, S2)));
  a0 = x0;
  a1 = pfma(q1, make_float2(1.f, -1.f), p1);                     // p1 - i q1 = (p1.x + q1.y, p1.y - q1.x)
  a4 = pfma(q1, make_float2(-1.f, 1.f), p1);                     // p1 + i q1
  a2 = pfma(q2, make_float2(1.f, -1.f), p2);
  a3 = pfma(q2, make_float2(-1.f, 1.f), p2);
}
// 16-point forward DFT. Input x[n], n = 4*m1 + m2. Output X[j1 + 4*j2] is left in slot 4*j1 + j2.
__device__ __forceinline__ void dft16(float2 (&x)[16]) {
#pragma unroll
  for (int m2 = 0; m2 < 4; ++m2) dft4(x[m2], x[4 + m2], x[8 + m2], x[12 + m2]);   // slot 4*j1+m2 = Y[m2][j1]
#pragma unroll
  for (int j1 = 1; j1 < 4; ++j1)
#pragma unroll
    for (int m2 = 1; m2 < 4; ++m2) x[4 * j1 + m2] = cmul(x[4 * j1 + m2], c_w16[(m2 * j1) & 15]);
#pragma unroll
  for (int j1 = 0; j1 < 4; ++j1) dft4(x[4 * j1], x[4 * j1 + 1], x[4 * j1 + 2], x[4 * j1 + 3]);
}
// 25-point forward DFT. Input x[n], n = 5*m1 + m2. Output X[j1 + 5*j2] is left in slot 5*j1 + j2.
__device__ __forceinline__ void dft25(float2 (&x)[25]) {
#pragma unroll
  for (int m2 = 0; m2 < 5; ++m2) dft5(x[m2], x[5 + m2], x[10 + m2], x[15 + m2], x[20 + m2]);
#pragma unroll
  for (int j1 = 1; j1 < 5; ++j1)
#pragma unroll
    for (int m2 = 1; m2 < 5; ++m2) x[5 * j1 + m2] = cmul(x[5 * j1 + m2], c_w25[(m2 * j1) % 25]);
#pragma unroll
  for (int j1 = 0; j1 < 5; ++j1) dft5(x[5 * j1], x[5 * j1 + 1], x[5 * j1 + 2], x[5 * j1 + 3], x[5 * j1 + 4]);
}
__host__ __device__ constexpr int dft25_slot(int k2) { return 5 * (k2 % 5) + k2 / 5; }      // slot that holds X[k2]

// ------------------------------------------------------------------------------------------------------------
// sample j of a segment whose first sample is element `base` of the PCM buffer: float32, or int16 scaled by 1/32768 (the
// wire format of the realtime path; backend/transcription_manager.py:45-51 does the same division on the host)
__device__ __forceinline__ float load_pcm(const float* __restrict__ pcm, long long base, int j, int flags) {
  if (flags & SONIC_MEL_S16) return (float)__ldg(reinterpret_cast<const short*>(pcm) + base + j) * (1.0f / 32768.0f);
  return __ldg(pcm + base + j);
}

__global__ void mel_peak_kernel(const float* __restrict__ pcm, const long long* __restrict__ offs,
                                const int* __restrict__ lens, unsigned* __restrict__ peak_bits, int flags) {
  const int b = blockIdx.y;
  const int n = min(lens[b], INT_MAX);
  const long long base = offs[b];
  float m = 0.f;
  // the reference normalises by the peak of the WHOLE input segment (asr.py:265), before the 30 s truncation
  if (!(flags & SONIC_MEL_S16) && ((base & 3) == 0)) {
    const float4* x4 = reinterpret_cast<const float4*>(pcm + base);
    const int n4 = n >> 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
      const float4 v = __ldg(x4 + i);
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (int i = 4 * n4 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(__ldg(pcm + base + i)));
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(load_pcm(pcm, base, i, flags)));
  }
  __shared__ float red[32];
  m = block_max(m, red);
  if (threadIdx.x == 0 && m > 0.f) atomicMax(peak_bits + b, __float_as_uint(m));   // non-negative floats order as uints
}

// the reference pre-step on one sample: v / peak (asr.py:266-267; IEEE division reproduced as q = v * RN(1/peak) followed by
// two FMA corrections — the second one is Markstein's correctly-rounding step) and the PCM_16 write + float read (asr.py:276)
__device__ __forceinline__ float prestep(float v, float peak, float rpeak, bool norm, bool pcm16) {
  if (norm) {
    float q = v * rpeak;
    q = fmaf(fmaf(-q, peak, v), rpeak, q);
    q = fmaf(fmaf(-q, peak, v), rpeak, q);
    v = q;
  }
  if (pcm16) v = rintf(v * 32767.0f) * (1.0f / 32768.0f);
  return v;
}

__device__ __forceinline__ void store_pair(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void store_pair(bf16* p, float a, float b) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b); }

struct TileItem { int b, tile, n, n_active; long long base; float peak; bool interior; };

template <typename TM>
__global__ void __launch_bounds__(kMelThreads, 2)
mel_frames_kernel(const float* __restrict__ pcm, const long long* __restrict__ offs, const int* __restrict__ lens,
                  const unsigned* __restrict__ peak_bits, const MelTables* __restrict__ tab, int flags, int batch, int tiles_per_seg, int tile_min_stride,
                  float* __restrict__ feat /*[B][128][3000] or null*/, TM* __restrict__ feat_tm /*[B][3002][128] or null*/,
                  float* __restrict__ tile_min /*[B][94]*/, unsigned* __restrict__ gmax_bits /*[B]*/) {
  extern __shared__ __align__(16) float smem[];
  float* s_samp = smem;                                            // kSpan
  float2* s_z = reinterpret_cast<float2*>(s_samp + kSpan);         // [pair][k1][n2]  (16 x 400 float2)
  float* s_P = reinterpret_cast<float*>(s_z + kPairs * kNfft);     // [frame][bin] powers
  float* s_win = s_P + kPFloats;                                   // 400
  float2* s_tw = reinterpret_cast<float2*>(s_win + kNfft);         // [k1][n2]
  float4* s_tapw = reinterpret_cast<float4*>(s_tw + kNfft);        // [128][3] float4 groups (zero beyond the tap count)
  int* s_meta = reinterpret_cast<int*>(s_tapw + kMels * (kMaxTaps / 4));
  float* s_out = reinterpret_cast<float*>(s_z);                    // [32 frames][129] staging of the time-major copy (s_z is dead after step 2)
  __shared__ unsigned s_ext[2][2];                                 // [slot][max, min] of a tile, order-preserving encoding

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kNfft; i += kMelThreads) { s_win[i] = tab->window[i]; s_tw[i] = tab->tw[i]; }
  for (int i = tid; i < kMels * kMaxTaps; i += kMelThreads) reinterpret_cast<float*>(s_tapw)[i] = tab->tapw[i];
  for (int i = tid; i < kMels; i += kMelThreads) s_meta[i] = tab->meta[i];
  const int cls1 = tab->cls_end[0], cls2 = tab->cls_end[1];
  for (int i = tid; i < 16; i += kMelThreads) s_P[kTileFrames * kPStride + i] = 0.f;

  // the next group's peak kernel (launched with programmatic stream serialization; it reads nothing this kernel writes) may
  // take the SMs this grid's CTAs leave at its tail instead of waiting for the whole grid
  pdl_launch_dependents();
  const bool norm_flag = (flags & SONIC_MEL_PEAK_NORM) != 0, pcm16 = (flags & SONIC_MEL_PCM16) != 0;
  const int total = batch * tiles_per_seg;
  // work item = (segment, tile of 32 frames); tiles that hold no transformed frame are skipped (CTA-uniform)
  auto describe = [&](int work, TileItem& it) -> bool {
    it.b = work / tiles_per_seg; it.tile = work - it.b * tiles_per_seg;
    it.n = min(lens[it.b], kWin);
    it.n_active = min(kFrames, (it.n + 200 + kHop - 1) / kHop);
    if (it.tile * kTileFrames >= it.n_active) return false;
    it.base = offs[it.b];
    it.peak = __uint_as_float(peak_bits[it.b]);
    const int j0 = it.tile * kTileFrames * kHop - 200;
    it.interior = j0 >= 0 && j0 + kSpan <= it.n && !(flags & SONIC_MEL_S16) && (((it.base + j0) & 3) == 0);
    return true;
  };
  auto next_item = [&](int work, TileItem& it) -> int {
    while (work < total && !describe(work, it)) work += gridDim.x;
    return work;
  };
  // span loads in two halves: issue (global -> registers, all in flight together) and commit (pre-step -> smem)
  float rawv[4 * kPer4];
  auto issue_loads = [&](const TileItem& it) {
    const int j0 = it.tile * kTileFrames * kHop - 200;
    if (it.interior) {      // the whole span lies inside the segment and is 16 B aligned: float4 loads
      const float4* x4 = reinterpret_cast<const float4*>(pcm + it.base + j0);
#pragma unroll
      for (int q = 0; q < kPer4; ++q) {
        const int i = tid + q * kMelThreads;
        const float4 v = (i < kSpan / 4) ? __ldg(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        rawv[4 * q] = v.x; rawv[4 * q + 1] = v.y; rawv[4 * q + 2] = v.z; rawv[4 * q + 3] = v.w;
      }
    } else {                // edges (reflection, zero padding), int16 input or unaligned base: scalar loads
#pragma unroll
      for (int q = 0; q < kPer1; ++q) {
        const int i = tid + q * kMelThreads;
        int j = j0 + i;
        if (j < 0) j = -j;
        if (j >= kWin) j = 2 * (kWin - 1) - j;
        rawv[q] = (i < kSpan && j < it.n) ? load_pcm(pcm, it.base, j, flags) : 0.f;
      }
    }
  };
  auto commit_loads = [&](const TileItem& it) {
    const bool norm = norm_flag && it.peak > 1e-6f;
    const float rpeak = norm ? __frcp_rn(it.peak) : 0.f;
    if (it.interior) {
#pragma unroll
      for (int q = 0; q < kPer4; ++q) {
        const int i = tid + q * kMelThreads;
        float4 v;
        v.x = prestep(rawv[4 * q], it.peak, rpeak, norm, pcm16); v.y = prestep(rawv[4 * q + 1], it.peak, rpeak, norm, pcm16);
        v.z = prestep(rawv[4 * q + 2], it.peak, rpeak, norm, pcm16); v.w = prestep(rawv[4 * q + 3], it.peak, rpeak, norm, pcm16);
        if (i < kSpan / 4) *reinterpret_cast<float4*>(s_samp + 4 * i) = v;
      }
    } else {
#pragma unroll
      for (int q = 0; q < kPer1; ++q) {
        const int i = tid + q * kMelThreads;
        if (i < kSpan) s_samp[i] = prestep(rawv[q], it.peak, rpeak, norm, pcm16);
      }
    }
  };

  TileItem cur, nxt;
  int prev_b = -1, prev_tile = 0, slot = 0;
  if (tid == 0) { s_ext[0][0] = s_ext[1][0] = f32_to_ordered(-10.0f); s_ext[0][1] = s_ext[1][1] = f32_to_ordered(0.f); }
  int work = next_item(blockIdx.x, cur);
  if (work < total) { issue_loads(cur); commit_loads(cur); }
  while (work < total) {
    // the next candidate item's segment fields are fetched now and looked at after step 2 (no dependent-load stall per tile)
    const int cand = work + gridDim.x, cand_b = cand / tiles_per_seg;
    int c_len = 0; long long c_off = 0; unsigned c_peak = 0;
    if (cand < total) { c_len = lens[cand_b]; c_off = offs[cand_b]; c_peak = peak_bits[cand_b]; }
    const int b = cur.b, tile = cur.tile, n_active = cur.n_active, t0 = tile * kTileFrames;
    __syncthreads();                                  // the span is in smem; the previous item's readers of s_P / s_out are done
    if (tid == 0) {
      if (prev_b >= 0) {                              // publish the previous tile's extrema
        atomicMax(gmax_bits + prev_b, s_ext[slot ^ 1][0]);
        tile_min[(size_t)prev_b * tile_min_stride + prev_tile] = ordered_to_f32(s_ext[slot ^ 1][1]);
      }
      // re-arm the slot the NEXT tile will use (its last reader ran above one iteration ago): maximum -10, minimum 0
      s_ext[slot ^ 1][0] = f32_to_ordered(-10.0f); s_ext[slot ^ 1][1] = f32_to_ordered(0.f);
    }

    // ---- step 1: for each (pair, n2): 16-point DFT over n1 of z[25*n1+n2], z = w*fA + i*w*fB; twiddle W400^{n2*k1}
    for (int it = tid; it < kPairs * 25; it += kMelThreads) {
      const int tr = it / 25, n2 = it - tr * 25;
      const float* fa = s_samp + (2 * tr) * kHop + n2;
      const float* wp = s_win + n2;
      float2 v[16];
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const float w = wp[25 * n1];
        v[n1] = make_float2(fa[25 * n1] * w, fa[25 * n1 + kHop] * w);
      }
      dft16(v);
      float2* zo = s_z + tr * kNfft + n2;
      const float2* twp = s_tw + n2;
#pragma unroll
      for (int s = 0; s < 16; ++s) {
        const int k1 = (s >> 2) + 4 * (s & 3);
        float2 y = v[s];
        if (k1 != 0) y = cmul(y, twp[k1 * 25]);        // W400^(n2 k1) (n2 = 0: exactly 1); [k1][n2] layout: no bank conflicts across n2
        zo[k1 * 25] = y;
      }
    }
    __syncthreads();
    // ---- step 2: thread (pair, k1): 25-point DFT over n2 -> Z[k1 + 16*k2]; then the split of the two real spectra.
    // Z = X_A + i X_B, so |X_A[k]|^2 = |Z[k] + conj Z[400-k]|^2 / 4 and |X_B[k]|^2 = |Z[k] - conj Z[400-k]|^2 / 4.  Bin 400 - k
    // sits in thread (16 - k1) & 15 of the same half-warp, slot 24 - k2 (k1 = 0: own slot (25 - k2) % 25).  The holder of
    // k <= 200 writes frame A's power of bin k, the holder of k > 200 frame B's power of bin 400 - k (same modulus seen from
    // the mirrored side); the self-mirrored bins 0 and 200 write both.
    {
      const int tr = tid >> 4, k1 = tid & 15;
      const float2* zi = s_z + tr * kNfft + k1 * 25;
      float2 v[25];
#pragma unroll
      for (int i = 0; i < 25; ++i) v[i] = zi[i];
      dft25(v);
      const int partner = (lane & 16) | ((16 - k1) & 15);
      float* PA = s_P + (2 * tr) * kPStride;
      float* PB = PA + kPStride;
#pragma unroll
      for (int k2 = 0; k2 < 25; ++k2) {
        const float2 z = v[dft25_slot(k2)];
        const float2 off = v[dft25_slot(24 - k2)];                         // what the partner wants from this thread
        float pr = __shfl_sync(0xffffffffu, off.x, partner), pi = __shfl_sync(0xffffffffu, off.y, partner);
        if (k1 == 0) { const float2 own = v[dft25_slot((25 - k2) % 25)]; pr = own.x; pi = own.y; }
        const int k = k1 + 16 * k2;
        const bool is_a = (k2 < 12) || (k2 == 12 && k1 <= 8);              // k <= 200
        const float sg = is_a ? 1.f : -1.f;
        const float ar = fmaf(sg, pr, z.x), ai = fmaf(-sg, pi, z.y);
        const float pw = 0.25f * (ar * ar + ai * ai);
        if (is_a) PA[k] = pw; else PB[kNfft - k] = pw;
        if ((k2 == 0 && k1 == 0) || (k2 == 12 && k1 == 8)) {               // bins 0 and 200 mirror onto themselves
          const float br = z.x - pr, bi = z.y + pi;
          PB[k] = 0.25f * (br * br + bi * bi);
        }
      }
    }
    __syncthreads();
    // ---- the next item's PCM loads fly during the mel phase (the span buffer is free: step 1 was its last reader)
    int work_next = cand;
    if (cand < total) {
      nxt.b = cand_b; nxt.tile = cand - cand_b * tiles_per_seg;
      nxt.n = min(c_len, kWin);
      nxt.n_active = min(kFrames, (nxt.n + 200 + kHop - 1) / kHop);
      if (nxt.tile * kTileFrames >= nxt.n_active) work_next = next_item(cand + gridDim.x, nxt);     // skipped tile: look further (rare)
      else {
        nxt.base = c_off; nxt.peak = __uint_as_float(c_peak);
        const int j0 = nxt.tile * kTileFrames * kHop - 200;
        nxt.interior = j0 >= 0 && j0 + kSpan <= nxt.n && !(flags & SONIC_MEL_S16) && (((nxt.base + j0) & 3) == 0);
      }
    }
    if (work_next < total) issue_loads(nxt);
    // ---- sparse mel + log10: lane = frame of the tile; a warp takes every 8th entry of the bin list, which is ordered by the
    // number of float4 tap groups so that each of the three loops below is fully unrolled
    float lmax = -10.0f, lmin = 0.f;
    {
      const int t = t0 + lane;
      const bool valid = t < n_active, inrange = t < kFrames;
      const float* P = s_P + lane * kPStride;
      float* frow = (feat && inrange) ? feat + ((size_t)b * kMels) * kFrames + t : nullptr;
      float* orow = s_out + lane * 129;
      auto one = [&](int i, auto groups) {
        constexpr int G = decltype(groups)::value;
        const int meta = s_meta[i], m = meta & 127;
        const float* p = P + (meta >> 8);
        const float4* w4 = s_tapw + m * (kMaxTaps / 4);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < G; ++c) {
          const float4 w = w4[c];
          acc = fmaf(w.x, p[4 * c], acc); acc = fmaf(w.y, p[4 * c + 1], acc);
          acc = fmaf(w.z, p[4 * c + 2], acc); acc = fmaf(w.w, p[4 * c + 3], acc);
        }
        // log10 via MUFU lg2 (relative error 2^-22: < 2e-7 in the log, far below the 1e-4 parity bar); the argument is >= 1e-10
        float l;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fmaxf(acc, 1e-10f)));
        l = valid ? l * 0.30102999566398120f : -10.0f;     // frames of the tile that only see zero padding: log10(1e-10)
        lmax = fmaxf(lmax, l);
        lmin = fminf(lmin, inrange ? l : 0.f);
        const float y = fmaf(l, 0.25f, 1.0f);              // (l + 4) / 4: the same value (scaling by 1/4 commutes with the rounding)
        if (frow) frow[(size_t)m * kFrames] = y;           // a warp writes 32 consecutive frames of one mel row
        if (feat_tm) orow[m] = y;
      };
      int i = warp;
      for (; i < cls1; i += kMelThreads / 32) one(i, std::integral_constant<int, 1>());
      for (; i < cls2; i += kMelThreads / 32) one(i, std::integral_constant<int, 2>());
      for (; i < kMels; i += kMelThreads / 32) one(i, std::integral_constant<int, 3>());
    }
    if (work_next < total) commit_loads(nxt);
    if (feat_tm) {                                         // [t][m]: 128 consecutive mels of one frame, two per thread and store
      __syncthreads();
      TM* dst = feat_tm + ((size_t)b * (kFrames + 2) + 1 + t0) * kMels;
#pragma unroll
      for (int q = 0; q < kTileFrames * (kMels / 2) / kMelThreads; ++q) {
        const int j = tid + q * kMelThreads, r = j >> 6, mp = j & 63;
        if (t0 + r < kFrames) store_pair(dst + r * kMels + 2 * mp, s_out[r * 129 + 2 * mp], s_out[r * 129 + 2 * mp + 1]);
      }
    }
    // tile extrema: warp shuffles, then one shared-memory atomic per warp into this tile's slot (two alternating slots: the
    // result is published after the next tile's first barrier, or after the loop — no extra barrier per tile)
    lmax = warp_max(lmax);
    lmin = -warp_max(-lmin);
    if (lane == 0) { atomicMax(&s_ext[slot][0], f32_to_ordered(lmax)); atomicMin(&s_ext[slot][1], f32_to_ordered(lmin)); }
    prev_b = b; prev_tile = tile;
    slot ^= 1;
    cur = nxt;
    work = work_next;
  }
  __syncthreads();
  if (tid == 0 && prev_b >= 0) {
    atomicMax(gmax_bits + prev_b, s_ext[slot ^ 1][0]);
    tile_min[(size_t)prev_b * tile_min_stride + prev_tile] = ordered_to_f32(s_ext[slot ^ 1][1]);
  }
}

// The clamp max(x, gmax - 8) in the written domain y = (x + 4) / 4 (a monotone map: clamping y at (gmax - 8 + 4) / 4 gives the
// same bits as clamping x first).  grid (16, segments): CTA (s, b) fills the never-transformed frames [pad0, 3000) of mel rows
// 8 s .. 8 s + 7 with 16 B stores (pad0 = first tile without a transformed frame), one sixteenth of the contiguous padded block
// of the time-major copy, and re-touches the transformed tiles s, s + 16, ... whose minimum lies below the floor.
template <typename T>
__global__ void __launch_bounds__(256) mel_fixup_kernel(const unsigned* __restrict__ gmax_bits, const int* __restrict__ lens,
                                                        const float* __restrict__ tile_min, int tiles_per_seg, float* __restrict__ feat, T* __restrict__ feat_tm) {
  const int b = blockIdx.y, s = blockIdx.x, ns = gridDim.x;
  const int n = min(lens[b], kWin);
  const int n_active = min(kFrames, (n + 200 + kHop - 1) / kHop);
  const int active_tiles = (n_active + kTileFrames - 1) / kTileFrames;
  const int pad0 = min(kFrames, active_tiles * kTileFrames);
  const float fl = ordered_to_f32(gmax_bits[b]) - 8.0f;
  const float y_floor = (fl + 4.0f) * 0.25f;
  const float y_pad = (fmaxf(-10.0f, fl) + 4.0f) * 0.25f;   // the constant value of silence after the clamp
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (feat_tm && s == 0 && tid < kMels) {                  // conv padding rows of the time-major copy
    feat_tm[((size_t)b * (kFrames + 2)) * kMels + tid] = from_f32<T>(0.f);
    feat_tm[((size_t)b * (kFrames + 2) + kFrames + 1) * kMels + tid] = from_f32<T>(0.f);
  }
  if (pad0 < kFrames) {
    if (feat) {                                            // rows are 12000 B and pad0 is a multiple of 32 frames: 16 B aligned
      const float4 v4 = make_float4(y_pad, y_pad, y_pad, y_pad);
      const int n4 = (kFrames - pad0) >> 2;
      for (int m = s * (kMels / ns) + warp; m < (s + 1) * (kMels / ns); m += 8) {
        float4* row = reinterpret_cast<float4*>(feat + ((size_t)b * kMels + m) * kFrames + pad0);
        for (int i = lane; i < n4; i += 32) row[i] = v4;
      }
    }
    if (feat_tm) {                                         // frames [pad0, 3000) x 128 mels are one contiguous block
      constexpr int kPer16 = 16 / (int)sizeof(T);
      __align__(16) T fill[kPer16];
#pragma unroll
      for (int i = 0; i < kPer16; ++i) fill[i] = from_f32<T>(y_pad);
      const uint4 v16 = *reinterpret_cast<const uint4*>(fill);
      uint4* blk = reinterpret_cast<uint4*>(feat_tm + ((size_t)b * (kFrames + 2) + 1 + pad0) * kMels);
      const int n16 = (kFrames - pad0) * kMels / kPer16;
      for (int i = s * 256 + tid; i < n16; i += ns * 256) blk[i] = v16;
    }
  }
  for (int tile = s; tile < active_tiles; tile += ns) {
    if (!(tile_min[(size_t)b * tiles_per_seg + tile] < fl)) continue;     // nothing below the floor in this tile: written once, done
    const int t0 = tile * kTileFrames;
    if (feat) {                                            // all 16 loads of a thread in flight, then the (rare) stores
      const int t = t0 + lane;
      float* p = feat + ((size_t)b * kMels + warp) * kFrames + t;
      float v[kMels / 8];
#pragma unroll
      for (int i = 0; i < kMels / 8; ++i) v[i] = (t < kFrames) ? p[(size_t)i * 8 * kFrames] : y_floor;
#pragma unroll
      for (int i = 0; i < kMels / 8; ++i)
        if (v[i] < y_floor) p[(size_t)i * 8 * kFrames] = y_floor;
    }
    if (feat_tm) {
      const int m = tid & (kMels - 1), r0 = tid >> 7;
      const T yf = from_f32<T>(y_floor);
      T* p = feat_tm + ((size_t)b * (kFrames + 2) + 1 + t0 + r0) * kMels + m;
      T v[kTileFrames / 2];
#pragma unroll
      for (int i = 0; i < kTileFrames / 2; ++i) v[i] = (t0 + r0 + 2 * i < kFrames) ? p[(size_t)i * 2 * kMels] : yf;
#pragma unroll
      for (int i = 0; i < kTileFrames / 2; ++i)
        if (to_f32(v[i]) < to_f32(yf)) p[(size_t)i * 2 * kMels] = yf;
    }
  }
}

__global__ void mel_init_kernel(unsigned* peak_bits, unsigned* gmax_bits, int nb) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb) { peak_bits[i] = 0u; gmax_bits[i] = f32_to_ordered(-10.0f); }
}

size_t mel_tables_bytes() { return sizeof(MelTables); }

void mel_build_tables(void* host_out, const int* tap_start, const int* tap_count, const float* tapw /*[128][12]*/) {
  MelTables* t = reinterpret_cast<MelTables*>(host_out);
  const double two_pi = 6.283185307179586476925286766559;
  for (int i = 0; i < kNfft; ++i) {
    // torch.hann_window(400) (periodic) is evaluated in fp32 by torch; the fp64->fp32 rounded value differs by <=1 ulp
    t->window[i] = (float)(0.5 - 0.5 * cos(two_pi * i / kNfft));
  }
  for (int k1 = 0; k1 < 16; ++k1)
    for (int n2 = 0; n2 < 25; ++n2) {
      const int e = (n2 * k1) % kNfft;
      t->tw[k1 * 25 + n2] = make_float2((float)cos(two_pi * e / kNfft), (float)(-sin(two_pi * e / kNfft)));
    }
  for (int m = 0; m < kMels; ++m) {
    t->tap_start[m] = tap_start[m];
    t->tap_count[m] = tap_count[m];
    for (int j = 0; j < kMaxTaps; ++j) t->tapw[m * kMaxTaps + j] = (j < tap_count[m]) ? tapw[m * kMaxTaps + j] : 0.f;
  }
  int n = 0;
  for (int g = 1; g <= kMaxTaps / 4; ++g) {
    for (int m = 0; m < kMels; ++m) {
      const int groups = tap_count[m] <= 0 ? 1 : (tap_count[m] + 3) / 4;
      if (groups == g) t->meta[n++] = m | (tap_start[m] << 8);
    }
    t->cls_end[g - 1] = n;
  }
  t->cls_end[3] = n;
}

static size_t mel_smem_bytes() {
  return sizeof(float) * (kSpan + 2 * kPairs * kNfft + kPFloats + kNfft + 2 * kNfft + kMels * kMaxTaps) + sizeof(int) * kMels;
}

cudaError_t mel_setup() {
  SONIC_CUDA_TRY(cudaFuncSetAttribute(mel_frames_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mel_smem_bytes()));
  return cudaFuncSetAttribute(mel_frames_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mel_smem_bytes());
}

int mel_tiles_per_segment() { return cdiv(kFrames, kTileFrames); }

template <typename T>
cudaError_t launch_mel(const float* pcm, const long long* offs, const int* lens, int batch, int max_len, int flags,
                       const void* tables, unsigned* peak_bits, unsigned* gmax_bits, float* tile_min, float* feat, T* feat_tm,
                       cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  mel_init_kernel<<<cdiv(batch, 128), 128, 0, st>>>(peak_bits, gmax_bits, batch);
  SONIC_LAUNCH_CHECK();
  const int n_eff = min(max_len, kWin);
  const int max_active = min(kFrames, (n_eff + 200 + kHop - 1) / kHop);
  const int tiles = cdiv(max_active, kTileFrames);                     // tiles that can hold transformed frames
  const int tps = mel_tiles_per_segment();
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // groups of <= 64 segments (82 MB of fp32 PCM): the frames kernel of a group re-reads from L2 what its peak pass just read
  // (measured at 1024 segments: 620 GB/s in groups of 64, 684 GB/s as one launch that reads the PCM twice from HBM; SONIC_MEL_GROUP)
  static const int kGroup = [] { const char* v = getenv("SONIC_MEL_GROUP"); const int g = v ? atoi(v) : 64; return g > 0 ? g : 64; }();
  static const bool mel_pdl = [] { const char* v = getenv("SONIC_NO_PDL"); return !(v && v[0] == '1'); }();
  for (int g0 = 0; g0 < batch; g0 += kGroup) {
    const int gb = min(kGroup, batch - g0);
    if (flags & SONIC_MEL_PEAK_NORM) {
      int gx = max(1, min(cdiv(max_len, 256 * 8), 64));
      // groups after the first: programmatic dependent launch behind the previous group's frames kernel (no data dependency)
      SONIC_CUDA_TRY(launch_ex(mel_peak_kernel, dim3(gx, gb), dim3(256), 0, st, g0 > 0 && mel_pdl, pcm, offs + g0, lens + g0, peak_bits + g0, flags));
    }
    const long long total = (long long)tiles * gb;
    const int grid = (int)(total < 2LL * sms ? total : 2LL * sms);      // two resident CTAs per SM, persistent
    mel_frames_kernel<T><<<grid, kMelThreads, mel_smem_bytes(), st>>>(
        pcm, offs + g0, lens + g0, peak_bits + g0, reinterpret_cast<const MelTables*>(tables), flags, gb, tiles, tps,
        feat ? feat + (size_t)g0 * kMels * kFrames : nullptr, feat_tm ? feat_tm + (size_t)g0 * (kFrames + 2) * kMels : nullptr,
        tile_min + (size_t)g0 * tps, gmax_bits + g0);
    SONIC_LAUNCH_CHECK();
  }
  mel_fixup_kernel<T><<<dim3(16, batch), 256, 0, st>>>(gmax_bits, lens, tile_min, tps, feat, feat_tm);
  SONIC_LAUNCH_CHECK();
  return cudaSuccess;
}

template cudaError_t launch_mel<float>(const float*, const long long*, const int*, int, int, int, const void*, unsigned*,
                                       unsigned*, float*, float*, float*, cudaStream_t);
template cudaError_t launch_mel<bf16>(const float*, const long long*, const int*, int, int, int, const void*, unsigned*,
                                      unsigned*, float*, float*, bf16*, cudaStream_t);

}  // namespace sonic
