// 16 kHz log-mel front end for sm_100a.
//
// Replaces (reference call chain): backend/asr.py:248-276 (peak-normalise + PCM_16 WAV round trip) and
// transformers/models/whisper/feature_extraction_whisper.py:135-164,296-337 (pad to 30 s, reflect-centred STFT
// n_fft=400 hop=160 periodic Hann, |.|^2, 128-bin slaney mel, log10 clamp, max-8 floor, (x+4)/4).
//
// Kernels
//   mel_peak_kernel     per-segment max|x| (the peak-normalise reduction).  Launched per group of <= 64 segments right before
//                       that group's frames kernel, so the second read of the PCM comes from the 126 MB L2, not from HBM.
//   mel_frames_kernel   pre-step on load (float4 loads on the aligned interior) -> smem framing/windowing -> 400-point FFT of
//                       frame PAIRS (two real frames as one complex transform, 400 = 16 x 25 Cooley-Tukey held in registers per
//                       thread, exchanged through shared memory) -> power -> sparse mel (<=9 taps) -> log10 -> (x+4)/4 written
//                       ONCE, unclamped, in the final layouts: fp32 [B,128,3000] (API/parity) and/or the time-major encoder
//                       input [B,3002,128]; per-tile minimum and per-segment maximum on the side.
//   mel_fixup_kernel    the max(x, gmax-8) clamp needs the segment maximum, known only after the last frame: this pass
//                       fills the frames that only see zero padding (never transformed: exactly log10(1e-10) = -10) and
//                       re-touches only the 32-frame tiles whose minimum is below gmax-8 (silence), instead of streaming a
//                       raw fp32 copy out and back in (round 1: 7.2 MB of traffic per segment against 2.8 MB algorithmic).
#include <stdlib.h>
#include "common.cuh"
#include "kernels.h"

namespace sonic {

static constexpr int kNfft = 400, kHop = 160, kBins = 201, kMels = 128, kFrames = 3000, kWin = 480000;
static constexpr int kPairs = 16;                 // complex transforms per CTA tile
static constexpr int kTileFrames = 2 * kPairs;    // 32 frames per tile
static constexpr int kSpan = (kTileFrames - 1) * kHop + kNfft;   // 5360 samples
static constexpr int kStride = 401;               // padded transform stride (bank-conflict-free mel reads)
static constexpr int kMelThreads = 256;
static constexpr int kMaxTaps = 12;

#include "mel_twiddles.inc"   // __constant__ float2 c_w16[16], c_w25[25]  (e^{-2 pi i k/n})

struct MelTables {             // built once on the host (api.cu) from the slaney filter bank, uploaded to global
  float window[kNfft];
  float2 tw[16 * 25];        // step-1 twiddles W400^(n2*k1) laid out [k1][n2]: consecutive lanes (n2) read consecutive entries
  float tapw[kMels * kMaxTaps];
  int tap_start[kMels];
  int tap_count[kMels];
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// forward 4-point DFT in place: (a0..a3) -> (X0..X3)
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
  float2 mi = make_float2(t3.y, -t3.x);          // -i * t3
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  a1 = cadd(t1, mi);
  a3 = csub(t1, mi);
}
// forward 5-point DFT in place
__device__ __forceinline__ void dft5(float2& a0, float2& a1, float2& a2, float2& a3, float2& a4) {
  const float c1 = 0.30901699437494745f, c2 = -0.8090169943749473f, s1 = 0.9510565162951535f, s2 = 0.5877852522924732f;
  float2 p = cadd(a1, a4), q = cadd(a2, a3), d1 = csub(a1, a4), d2 = csub(a2, a3);
  float2 x0 = make_float2(a0.x + p.x + q.x, a0.y + p.y + q.y);
  float2 p1 = make_float2(a0.x + c1 * p.x + c2 * q.x, a0.y + c1 * p.y + c2 * q.y);
  float2 p2 = make_float2(a0.x + c2 * p.x + c1 * q.x, a0.y + c2 * p.y + c1 * q.y);
  float2 q1 = make_float2(s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y);
  float2 q2 = make_float2(s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y);
  a0 = x0;
  a1 = make_float2(p1.x + q1.y, p1.y - q1.x);    // p1 - i q1
  a4 = make_float2(p1.x - q1.y, p1.y + q1.x);    // p1 + i q1
  a2 = make_float2(p2.x + q2.y, p2.y - q2.x);
  a3 = make_float2(p2.x - q2.y, p2.y + q2.x);
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
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(load_pcm(pcm, base, i, flags)));
  __shared__ float red[32];
  m = block_max(m, red);
  if (threadIdx.x == 0 && m > 0.f) atomicMax(peak_bits + b, __float_as_uint(m));   // non-negative floats order as uints
}

template <typename TM>
__global__ void __launch_bounds__(kMelThreads, 2)
mel_frames_kernel(const float* __restrict__ pcm, const long long* __restrict__ offs, const int* __restrict__ lens,
                  const unsigned* __restrict__ peak_bits, const MelTables* __restrict__ tab, int flags, int batch, int tiles_per_seg, int tile_min_stride,
                  float* __restrict__ feat /*[B][128][3000] or null*/, TM* __restrict__ feat_tm /*[B][3002][128] or null*/,
                  float* __restrict__ tile_min /*[B][94]*/, unsigned* __restrict__ gmax_bits /*[B]*/) {
  extern __shared__ float smem[];
  float* s_samp = smem;                              // kSpan
  float* s_re = s_samp + kSpan;                      // kPairs*kStride
  float* s_im = s_re + kPairs * kStride;             // kPairs*kStride   (base offset == 16 mod 32 banks)
  float* s_win = s_im + kPairs * kStride;            // 400
  float2* s_tw = reinterpret_cast<float2*>(s_win + kNfft);     // 16 x 25
  float* s_tapw = reinterpret_cast<float*>(s_tw + kNfft);      // 128*12
  int* s_tstart = reinterpret_cast<int*>(s_tapw + kMels * kMaxTaps);
  int* s_tcount = s_tstart + kMels;
  __shared__ float red[32];

  const int tid = threadIdx.x;
  for (int i = tid; i < kNfft; i += kMelThreads) { s_win[i] = tab->window[i]; s_tw[i] = tab->tw[i]; }
  for (int i = tid; i < kMels * kMaxTaps; i += kMelThreads) s_tapw[i] = tab->tapw[i];
  for (int i = tid; i < kMels; i += kMelThreads) { s_tstart[i] = tab->tap_start[i]; s_tcount[i] = tab->tap_count[i]; }

  // persistent CTAs: work item = (segment, tile of 32 frames); the tables above are loaded once per CTA
  for (int work = blockIdx.x; work < batch * tiles_per_seg; work += gridDim.x) {
    const int b = work / tiles_per_seg, tile = work - b * tiles_per_seg;
    const int n = min(lens[b], kWin);
    const int n_active = min(kFrames, (n + 200 + kHop - 1) / kHop);
    if (tile * kTileFrames >= n_active) continue;               // CTA-uniform
    const long long base = offs[b];
    const float peak = __uint_as_float(peak_bits[b]);
    const float gate = (peak > 1e-6f) ? 1.f : 0.f;
    float lmax = -10.0f;
    const int t0 = tile * kTileFrames;
    __syncthreads();                                  // previous item's readers are done with smem
    const int j0 = t0 * kHop - 200;
    const bool interior = j0 >= 0 && j0 + kSpan <= n && !(flags & SONIC_MEL_S16) && (((base + j0) & 3) == 0);
    if (interior) {
      // the whole span lies inside the segment and is 16 B aligned: float4 loads, all issued before the first value is used
      constexpr int kPer4 = (kSpan / 4 + kMelThreads - 1) / kMelThreads;                    // 1340 float4 -> 6 per thread
      const float4* x4 = reinterpret_cast<const float4*>(pcm + base + j0);
      float4 rawv[kPer4];
#pragma unroll
      for (int q = 0; q < kPer4; ++q) {
        const int i = tid + q * kMelThreads;
        rawv[q] = (i < kSpan / 4) ? __ldg(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int q = 0; q < kPer4; ++q) {
        const int i = tid + q * kMelThreads;
        float v[4] = {rawv[q].x, rawv[q].y, rawv[q].z, rawv[q].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if ((flags & SONIC_MEL_PEAK_NORM) && gate > 0.f) v[e] = v[e] / peak;              // asr.py:266-267 (true division)
          if (flags & SONIC_MEL_PCM16) v[e] = rintf(v[e] * 32767.0f) * (1.0f / 32768.0f);   // soundfile PCM_16 write + float read
        }
        if (i < kSpan / 4) *reinterpret_cast<float4*>(s_samp + 4 * i) = make_float4(v[0], v[1], v[2], v[3]);
      }
    } else {
      // edges (reflection, zero padding), int16 input or unaligned base: scalar loads, still all in flight together
      constexpr int kPer = (kSpan + kMelThreads - 1) / kMelThreads;
      float rawv[kPer];
#pragma unroll
      for (int q = 0; q < kPer; ++q) {
        const int i = tid + q * kMelThreads;
        int j = j0 + i;
        if (j < 0) j = -j;
        if (j >= kWin) j = 2 * (kWin - 1) - j;
        rawv[q] = (i < kSpan && j < n) ? load_pcm(pcm, base, j, flags) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < kPer; ++q) {
        const int i = tid + q * kMelThreads;
        float v = rawv[q];
        if ((flags & SONIC_MEL_PEAK_NORM) && gate > 0.f) v = v / peak;              // asr.py:266-267 (true division)
        if (flags & SONIC_MEL_PCM16) v = rintf(v * 32767.0f) * (1.0f / 32768.0f);   // soundfile PCM_16 write + float read
        if (i < kSpan) s_samp[i] = v;
      }
    }
    __syncthreads();

    // ---- step 1: for each (pair, n2): 16-point DFT over n1 of z[25*n1+n2], z = wA*fA + i*wB*fB; twiddle W400^{n2*k1}
    for (int it = tid; it < kPairs * 25; it += kMelThreads) {
      const int tr = it / 25, n2 = it - tr * 25;
      const float* fa = s_samp + (2 * tr) * kHop;
      float2 v[16];
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const int idx = 25 * n1 + n2;
        const float w = s_win[idx];
        v[n1] = make_float2(fa[idx] * w, fa[idx + kHop] * w);
      }
      dft16(v);
#pragma unroll
      for (int s = 0; s < 16; ++s) {
        const int k1 = (s >> 2) + 4 * (s & 3);
        float2 y = v[s];
        if (k1 != 0 && n2 != 0) y = cmul(y, s_tw[k1 * 25 + n2]);   // W400^(n2 k1); [k1][n2] layout: no bank conflicts across n2
        s_re[tr * kStride + k1 * 25 + n2] = y.x;
        s_im[tr * kStride + k1 * 25 + n2] = y.y;
      }
    }
    __syncthreads();
    // ---- step 2: for each (pair, k1): 25-point DFT over n2 -> Z[k1 + 16*k2], kept in place at k1*25 + k2
    for (int it = tid; it < kPairs * 16; it += kMelThreads) {
      const int tr = it >> 4, k1 = it & 15;
      float* pr = s_re + tr * kStride + k1 * 25;
      float* pi = s_im + tr * kStride + k1 * 25;
      float2 v[25];
#pragma unroll
      for (int i = 0; i < 25; ++i) v[i] = make_float2(pr[i], pi[i]);
      dft25(v);
#pragma unroll
      for (int s = 0; s < 25; ++s) {
        const int k2 = (s / 5) + 5 * (s % 5);
        pr[k2] = v[s].x;
        pi[k2] = v[s].y;
      }
    }
    __syncthreads();
    // ---- split the two real spectra and take |.|^2 in place: re <- P_A[k], im <- P_B[k], k = 0..200
    for (int it = tid; it < kPairs * kBins; it += kMelThreads) {
      const int tr = it / kBins, k = it - tr * kBins;
      const int kk = (k == 0) ? 0 : (kNfft - k);
      const int a0 = tr * kStride + (k & 15) * 25 + (k >> 4);
      const int a1 = tr * kStride + (kk & 15) * 25 + (kk >> 4);
      const float zr = s_re[a0], zi = s_im[a0], yr = s_re[a1], yi = s_im[a1];
      const float ar = zr + yr, ai = zi - yi;        // 2*X_A
      const float br = zr - yr, bi = zi + yi;        // 2*i*X_B (same modulus)
      // in-place is safe: a1 addresses bins >= 200, which no item writes (k = 200 maps to itself)
      s_re[a0] = 0.25f * (ar * ar + ai * ai);
      s_im[a0] = 0.25f * (br * br + bi * bi);
    }
    __syncthreads();
    // ---- sparse mel + log10: lane = frame of the tile, warp strides over mel bins; (l + 4) / 4 staged as a [128][33] tile
    float* s_out = s_samp;                               // the sample span is dead after step 1 (5360 >= 128 * 33 floats)
    float lmin = 0.f;
    {
      const int lane = tid & 31, warp = tid >> 5;
      const int t = t0 + lane;
      const float* P = ((lane & 1) ? s_im : s_re) + (lane >> 1) * kStride;
      for (int m = warp; m < kMels; m += kMelThreads / 32) {
        const int ks = s_tstart[m], kc = s_tcount[m];
        float acc = 0.f;
        for (int j = 0; j < kc; ++j) {
          const int k = ks + j;
          acc = fmaf(s_tapw[m * kMaxTaps + j], P[(k & 15) * 25 + (k >> 4)], acc);
        }
        // log10 via MUFU lg2 (relative error 2^-22: < 2e-7 in the log, far below the 1e-4 parity bar)
        float l = __log2f(fmaxf(acc, 1e-10f)) * 0.30102999566398120f;
        if (t < n_active) lmax = fmaxf(lmax, l);
        else l = -10.0f;                                   // frames of the tile that only see zero padding
        if (t < kFrames) lmin = fminf(lmin, l);
        s_out[m * 33 + lane] = (l + 4.0f) * 0.25f;
      }
    }
    __syncthreads();
    if (feat) {                                            // [m][t]: a warp writes 32 consecutive frames of one mel row
      const int lane = tid & 31, warp = tid >> 5;
      const int t = t0 + lane;
      if (t < kFrames)
        for (int m = warp; m < kMels; m += kMelThreads / 32) feat[((size_t)b * kMels + m) * kFrames + t] = s_out[m * 33 + lane];
    }
    if (feat_tm) {                                         // [t][m]: 128 consecutive mels of one frame
      const int m = tid & (kMels - 1);
      for (int r = tid >> 7; r < kTileFrames; r += kMelThreads / kMels) {
        const int t = t0 + r;
        if (t < kFrames) feat_tm[((size_t)b * (kFrames + 2) + 1 + t) * kMels + m] = from_f32<TM>(s_out[m * 33 + r]);
      }
    }
    lmax = block_max(lmax, red);
    lmin = -block_max(-lmin, red);
    if (tid == 0) {
      atomicMax(gmax_bits + b, f32_to_ordered(lmax));
      tile_min[(size_t)b * tile_min_stride + tile] = lmin;
    }
  }
}

// The clamp max(x, gmax - 8) in the written domain y = (x + 4) / 4 (a monotone map: clamping y at (gmax - 8 + 4) / 4 gives the
// same bits as clamping x first).  grid (94 tiles, segments).
template <typename T>
__global__ void __launch_bounds__(256) mel_fixup_kernel(const unsigned* __restrict__ gmax_bits, const int* __restrict__ lens,
                                                        const float* __restrict__ tile_min, int tiles_per_seg, float* __restrict__ feat, T* __restrict__ feat_tm) {
  const int b = blockIdx.y, tile = blockIdx.x, t0 = tile * kTileFrames;
  const int n = min(lens[b], kWin);
  const int n_active = min(kFrames, (n + 200 + kHop - 1) / kHop);
  const float fl = ordered_to_f32(gmax_bits[b]) - 8.0f;
  const float y_floor = (fl + 4.0f) * 0.25f;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (feat_tm && tile == 0 && tid < kMels) {               // conv padding rows of the time-major copy
    feat_tm[((size_t)b * (kFrames + 2)) * kMels + tid] = from_f32<T>(0.f);
    feat_tm[((size_t)b * (kFrames + 2) + kFrames + 1) * kMels + tid] = from_f32<T>(0.f);
  }
  if (t0 >= n_active) {                                    // never transformed: the constant value of silence after the clamp
    const float y_pad = (fmaxf(-10.0f, fl) + 4.0f) * 0.25f;
    if (feat) {
      const int t = t0 + lane;
      if (t < kFrames)
        for (int m = warp; m < kMels; m += 8) feat[((size_t)b * kMels + m) * kFrames + t] = y_pad;
    }
    if (feat_tm) {
      const int m = tid & (kMels - 1);
      for (int r = tid >> 7; r < kTileFrames; r += 2) {
        const int t = t0 + r;
        if (t < kFrames) feat_tm[((size_t)b * (kFrames + 2) + 1 + t) * kMels + m] = from_f32<T>(y_pad);
      }
    }
    return;
  }
  if (!(tile_min[(size_t)b * tiles_per_seg + tile] < fl)) return;     // nothing below the floor in this tile: written once, done
  if (feat) {
    const int t = t0 + lane;
    if (t < kFrames)
      for (int m = warp; m < kMels; m += 8) {
        float* p = feat + ((size_t)b * kMels + m) * kFrames + t;
        if (*p < y_floor) *p = y_floor;
      }
  }
  if (feat_tm) {
    const int m = tid & (kMels - 1);
    const T yf = from_f32<T>(y_floor);
    for (int r = tid >> 7; r < kTileFrames; r += 2) {
      const int t = t0 + r;
      if (t < kFrames) {
        T* p = feat_tm + ((size_t)b * (kFrames + 2) + 1 + t) * kMels + m;
        if (to_f32(*p) < to_f32(yf)) *p = yf;
      }
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
    for (int j = 0; j < kMaxTaps; ++j) t->tapw[m * kMaxTaps + j] = tapw[m * kMaxTaps + j];
  }
}

static size_t mel_smem_bytes() {
  return sizeof(float) * (kSpan + 2 * kPairs * kStride + kNfft + 2 * kNfft + kMels * kMaxTaps) + sizeof(int) * 2 * kMels;
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
  for (int g0 = 0; g0 < batch; g0 += kGroup) {
    const int gb = min(kGroup, batch - g0);
    if (flags & SONIC_MEL_PEAK_NORM) {
      int gx = max(1, min(cdiv(max_len, 256 * 8), 64));
      mel_peak_kernel<<<dim3(gx, gb), 256, 0, st>>>(pcm, offs + g0, lens + g0, peak_bits + g0, flags);
      SONIC_LAUNCH_CHECK();
    }
    const long long total = (long long)tiles * gb;
    const int grid = (int)(total < 2LL * sms ? total : 2LL * sms);      // two resident CTAs per SM, persistent
    mel_frames_kernel<T><<<grid, kMelThreads, mel_smem_bytes(), st>>>(
        pcm, offs + g0, lens + g0, peak_bits + g0, reinterpret_cast<const MelTables*>(tables), flags, gb, tiles, tps,
        feat ? feat + (size_t)g0 * kMels * kFrames : nullptr, feat_tm ? feat_tm + (size_t)g0 * (kFrames + 2) * kMels : nullptr,
        tile_min + (size_t)g0 * tps, gmax_bits + g0);
    SONIC_LAUNCH_CHECK();
  }
  mel_fixup_kernel<T><<<dim3(tps, batch), 256, 0, st>>>(gmax_bits, lens, tile_min, tps, feat, feat_tm);
  SONIC_LAUNCH_CHECK();
  return cudaSuccess;
}

template cudaError_t launch_mel<float>(const float*, const long long*, const int*, int, int, int, const void*, unsigned*,
                                       unsigned*, float*, float*, float*, cudaStream_t);
template cudaError_t launch_mel<bf16>(const float*, const long long*, const int*, int, int, int, const void*, unsigned*,
                                      unsigned*, float*, float*, bf16*, cudaStream_t);

}  // namespace sonic
