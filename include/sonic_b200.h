/* libsonic_b200 — C ABI of the B200-native SonicScribe transcription hot path.
 *
 * The reference has no FFI of its own: the boundary it exposes is the Python method
 *     ASRModel.transcribe(audio_tensor[1,N] float32 CPU, sampling_rate, max_new_tokens, hotwords) -> str
 * (/root/reference/backend/asr.py:335-342), which internally calls
 *     _prepare_audio_tempfile          asr.py:230-278   (peak-normalise + PCM_16 WAV round trip)
 *     processor.apply_chat_template    asr.py:393-399   (log-mel via WhisperFeatureExtractor + prompt ids)
 *     model.generate(do_sample=False)  asr.py:407-422   (encoder, adapter, prefill, greedy KV-cache decode)
 * Each entry point below names the reference call it replaces.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on failure; sonic_last_error(h) gives the message
 *     (h may be NULL for creation failures).  Nothing aborts the process.
 *   - one handle == one model replica on one GPU with its own CUDA stream; calls on a handle are serialised by an
 *     internal mutex, so N host threads may share a handle (asr.py is called from 3 executor threads + the event
 *     loop, backend/main.py:429-445, transcription_manager.py:58) and N handles drive N GPUs concurrently.
 *   - host pointers unless a flag says otherwise.  "segments" are independent <=30 s utterances processed together.
 */
#ifndef SONIC_B200_H
#define SONIC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SONIC_API __attribute__((visibility("default")))
#else
#define SONIC_API
#endif

typedef struct sonic_ctx* sonic_handle;

enum { SONIC_MODE_BF16 = 0, SONIC_MODE_FP32 = 1, SONIC_MODE_INT8 = 2 };
enum { SONIC_DTYPE_F32 = 0, SONIC_DTYPE_BF16 = 1 };
/* return codes: 0 ok, -1 failure (message in sonic_last_error), -2 sonic_load_tensor was given a name the model does not have */
#define SONIC_ERR_UNKNOWN_TENSOR (-2)

/* flags of sonic_mel / sonic_transcribe_batch */
#define SONIC_FLAG_PEAK_NORM   0x01  /* asr.py:265-267  wav / max|wav| when max > 1e-6                         */
#define SONIC_FLAG_PCM16       0x02  /* asr.py:276      soundfile PCM_16 write + float re-load                  */
#define SONIC_FLAG_PCM_S16     0x04  /* pcm holds int16 LE samples (the WebSocket wire format, pcm-processor.js:59-75): widened to
                                        float32 / 32768 on the device, as transcription_manager.py:45-51 does on the host        */
#define SONIC_FLAG_PCM_DEVICE  0x10  /* pcm pointer is device memory (bench: inputs resident in HBM)            */
#define SONIC_FLAG_OUT_DEVICE  0x20  /* features pointer of sonic_mel is device memory                          */
#define SONIC_FLAG_FEATURES_ONLY 0x40 /* sonic_mel writes `features` only, not the encoder's time-major copy (front-end
                                        bandwidth sweep, BASELINE.json configs[4]); a following sonic_encode is refused   */
#define SONIC_FLAG_SHORT_WINDOW 0x80  /* OPT-IN streaming encoder for interim calls (audio_manager.py:106-114 hands
                                        the last 1.28 s to transcribe, which the reference pads to 30 s): when every segment of the
                                        call has the same frame count F, the encoder runs over T = ceil(F/2) rounded up to 8
                                        positions instead of 1500.  NOT the reference's numbers — the reference attends over the
                                        padded positions too; parity is against the same HF classes fed the truncated features
                                        input_features[:, :, :2T].  Calls with mixed lengths use the full window.                  */
#define SONIC_FLAG_REFERENCE_PRESTEP (SONIC_FLAG_PEAK_NORM | SONIC_FLAG_PCM16)

typedef struct {
  int32_t device;        /* CUDA device ordinal                                                                */
  int32_t mode;          /* SONIC_MODE_*  — native bf16 (asr.py:61), fp32 parity arithmetic, int8 weight-only   */
  int32_t enc_layers;    /* 32 for GLM-ASR-Nano-2512; smaller values are for tests                              */
  int32_t dec_layers;    /* 28                                                                                  */
  int32_t max_batch;     /* segments per call                                                                   */
  int32_t max_prompt;    /* prompt tokens per segment (<= 8 + 375 + hotword text)                              */
  int32_t max_new;       /* upper bound of max_new_tokens                                                       */
  int32_t debug;         /* 1: keep probe copies of intermediate tensors for sonic_debug_read                   */
} sonic_config;

/* replaces AutoModel.from_pretrained(...) / ASRModel.__init__ (asr.py:25-87, 120-146) */
SONIC_API int sonic_create(const sonic_config* cfg, sonic_handle* out);
SONIC_API int sonic_destroy(sonic_handle h);
SONIC_API const char* sonic_last_error(sonic_handle h);
SONIC_API const char* sonic_version(void);

/* Weight upload: one call per tensor of the HF state dict (names as in SURVEY.md §8a row W, e.g.
 * "audio_tower.layers.0.self_attn.q_proj.weight").  data is host memory, row-major, dtype SONIC_DTYPE_*.
 * sonic_finalize_weights checks that every tensor arrived.  Replaces the checkpoint load of asr.py:137-140 and, for
 * SONIC_MODE_INT8, _quantize_model_int8 (asr.py:169-210) with per-output-row absmax weight-only int8. */
SONIC_API int sonic_load_tensor(sonic_handle h, const char* name, const void* data, int32_t dtype, const int64_t* shape, int32_t ndim);
SONIC_API int sonic_finalize_weights(sonic_handle h);

/* Pre-step + log-mel for `batch` segments.  Segment b is pcm[offsets[b] .. offsets[b]+lengths[b]) (float32, 16 kHz mono).
 * With SONIC_FLAG_PCM_S16 `pcm` points at int16 samples instead (offsets still count samples).
 * Writes input_features [batch,128,3000] float32 to `features` (may be NULL) and the valid-frame count
 * (input_features_mask.sum(), ceil(min(n,480000)/160)) to n_frames[b].  The time-major copy the encoder consumes stays
 * in the handle.  Replaces asr.py:230-278 + WhisperFeatureExtractor.__call__
 * (transformers/models/whisper/feature_extraction_whisper.py:135-164,296-337). */
SONIC_API int sonic_mel(sonic_handle h, const void* pcm, const int64_t* offsets, const int32_t* lengths, int32_t batch,
              int32_t flags, float* features, int32_t* n_frames);

/* Encoder + adapter on the features left in the handle by sonic_mel.  audio_embeds (may be NULL) receives
 * [batch,375,2048] float32; rows >= n_audio[b] are the adapter output of padded frames and are ignored downstream.
 * Replaces GlmAsrForConditionalGeneration.get_audio_features (transformers/models/glmasr/modeling_glmasr.py:394-426). */
SONIC_API int sonic_encode(sonic_handle h, int32_t batch, float* audio_embeds, int32_t* n_audio);

/* Prefill + greedy decode.  ids: concatenated prompt token ids; id_offsets[batch+1].  Positions holding the audio
 * placeholder 59260 receive the audio embeddings in order (modeling_glmasr.py:473-483).  out_ids [batch,max_new_tokens],
 * n_out[batch]; margins (may be NULL) [batch,max_new_tokens] top-1 minus top-2 logit of every step.
 * Replaces model.generate(**inputs, max_new_tokens, do_sample=False) (asr.py:411-422;
 * transformers/generation/utils.py:2658-2841). */
SONIC_API int sonic_generate(sonic_handle h, const int32_t* ids, const int32_t* id_offsets, int32_t batch, int32_t max_new_tokens,
                   int32_t* out_ids, int32_t* n_out, float* margins);

/* The whole of ASRModel.transcribe up to (not including) tokenizer decode, for a batch of segments. */
SONIC_API int sonic_transcribe_batch(sonic_handle h, const void* pcm, const int64_t* offsets, const int32_t* lengths, int32_t batch,
                           int32_t flags, const int32_t* ids, const int32_t* id_offsets, int32_t max_new_tokens,
                           int32_t* out_ids, int32_t* n_out, float* margins);

/* audio-token count of a segment of n samples: processing_glmasr.py:97-103 */
SONIC_API int32_t sonic_num_audio_tokens(int64_t n_samples);

/* Instrumentation (asr.py:357-366,431-443 times with CUDA events around the call) */
SONIC_API int sonic_sync(sonic_handle h);
SONIC_API int sonic_timer_begin(sonic_handle h);                 /* records an event on the handle's stream                    */
SONIC_API int sonic_timer_end(sonic_handle h, float* ms);        /* records, synchronises, returns elapsed device time        */
SONIC_API int sonic_stage_times(sonic_handle h, float* ms4);     /* device ms of the last transcribe: mel, encode, prefill, decode */
SONIC_API int64_t sonic_launch_count(sonic_handle h);            /* kernels launched through this handle so far                */
SONIC_API int64_t sonic_device_bytes(sonic_handle h);            /* device memory held by the handle                           */
/* Per-launch-class device-time breakdown: between begin and end every kernel launch of the handle is bracketed by an event
 * pair on its stream (the decode loop runs eagerly instead of as a CUDA graph).  Classes: sonic_profile_class_name(i). */
SONIC_API int sonic_profile_begin(sonic_handle h);
SONIC_API int sonic_profile_end(sonic_handle h, float* ms_per_class, int64_t* launches_per_class, int32_t n_classes);
SONIC_API int32_t sonic_profile_num_classes(void);
SONIC_API const char* sonic_profile_class_name(int32_t c);
/* Copy an intermediate tensor as float32 (handle created with debug=1):
 * "mel_tm", "conv_out", "enc_layer0", "enc_out", "audio_embeds", "first_logits", "dec_layer0", "rope_enc_cos", ... */
/* debug handles: keep the full fp32 logit rows of these greedy steps (step 0 = prefill); read them back as
 * "step_logits@<step>" [batch, 59264].  At most 8 steps; n = 0 clears.  Such handles run the decode loop eagerly. */
SONIC_API int sonic_debug_set_logit_steps(sonic_handle h, const int32_t* steps, int32_t n);
SONIC_API int sonic_debug_read(sonic_handle h, const char* name, float* out, size_t max_elems, size_t* n_elems);

/* Stand-alone kernel entry points for the unit tests and the roofline bench (device pointers, handle's stream).
 * gemm: C[M,N] = A[M,K] . W[N,K]^T (+bias[N]) in bf16 on tcgen05 (impl 0; swap=1 selects the decode orientation) or on
 * the CUDA-core cross-check path (impl 1).  All pointers are HOST float32; the call converts, runs and converts back. */
SONIC_API int sonic_test_gemm(sonic_handle h, int32_t impl, int32_t swap, const float* A, const float* W, const float* bias,
                    const float* resid, float* C, int32_t M, int32_t N, int32_t K, int32_t act);

/* same as sonic_test_gemm for the int8 weight-only operand path: W (host float32) is quantised per output row on the device
 * (s = max|row|, q = rint(127 w / s)) and expanded to bf16 inside the kernel; C = (A . q^T) * s/127 (+bias, act, resid). */
SONIC_API int sonic_test_gemm_int8(sonic_handle h, int32_t swap, const float* A, const float* W, const float* bias, const float* resid,
                                   float* C, int32_t M, int32_t N, int32_t K, int32_t act);
/* back-to-back launches of one tcgen05 GEMM shape (zero-filled bf16 operands, weights rotated through > 126 MB so L2
 * never holds them); returns the average device time per launch in microseconds (CUDA events on the handle's stream). */
SONIC_API int sonic_bench_gemm(sonic_handle h, int32_t swap, int32_t M, int32_t N, int32_t K, int32_t act, int32_t iters, float* avg_us);
/* tcgen05.mma rate microbenchmark (one CTA): clocks per MMA of shape m x ntok x 16 (bf16, operands in shared memory) cycling through
 * n_acc TMEM accumulators and n_tiles different 16 KB A tiles: issue-loop clocks and clocks until the last MMA has completed. */
SONIC_API int sonic_bench_mma(sonic_handle h, int32_t m, int32_t ntok, int32_t n_mma, int32_t n_acc, int32_t n_tiles, float* issue_clk, float* total_clk);
/* encoder attention alone (20 heads x 64, non-causal, scale 1/8) on a fused [segments*T, 3840] q|k|v buffer (host float32,
 * rounded to bf16 on the device): impl 0 = tcgen05 kernel, impl 1 = CUDA-core cross-check.  out: [segments*T, 1280]. */
SONIC_API int sonic_test_enc_attention(sonic_handle h, int32_t impl, const float* qkv, float* out, int32_t segments, int32_t T);

#ifdef __cplusplus
}
#endif
#endif /* SONIC_B200_H */
