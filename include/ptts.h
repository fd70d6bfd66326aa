/*
 * ptts.h — C ABI of libptts_cuda.so, the B200 (sm_100a) engine for the Pocket TTS
 * generation hot path: FlowLM autoregressive decode step, LSD flow head, streaming
 * Mimi decoder.
 *
 * The reference (ykevinc/pocket-tts) has no FFI; its seam is the public Rust API of
 * `TTSModel` (crates/pocket-tts/src/lib.rs:15-18).  Every entry point below names the
 * reference function it replaces.  A Rust `pocket-tts-cuda` crate binds exactly these
 * symbols; tests and bench.py bind them through ctypes.
 *
 * Conventions: all functions return 0 on success and a negative ptts_status otherwise;
 * ptts_last_error() returns a NUL-terminated message for the calling thread's last
 * failure.  Handles are opaque.  The library never frees caller memory; the caller
 * frees library objects only through the *_destroy / *_close calls.  One engine owns
 * one CUDA device and one stream set; an engine is not re-entrant (the Rust facade
 * wraps it in a Mutex, harnesses use one thread per GPU).  There is no CPU fallback:
 * without a CUDA device every compute entry point fails with PTTS_ERR_CUDA.
 */
#ifndef PTTS_H
#define PTTS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTTS_ABI_VERSION 2

typedef enum {
  PTTS_OK = 0,
  PTTS_ERR_INVALID = -1,   /* bad argument / unknown tensor / wrong shape */
  PTTS_ERR_CUDA = -2,      /* CUDA runtime or driver error, or no device */
  PTTS_ERR_CAPACITY = -3,  /* no free slot / KV capacity exceeded */
  PTTS_ERR_STATE = -4      /* call not valid in the current state */
} ptts_status;

typedef enum { PTTS_F32 = 0, PTTS_BF16 = 1, PTTS_F16 = 2 } ptts_dtype;

/* How GEMM weights are held in HBM.  PTTS_W_F16: 16-bit float operands, f32 accumulate
 * (checkpoint values are bf16-representable, tts_model.py:143-145, and convert exactly).
 * PTTS_W_INT8: per-tensor symmetric int8 with the reference's scheme and skip list
 * (crates/pocket-tts/src/quantize.rs:27-41,65-94), dequantised in registers. */
typedef enum { PTTS_W_F16 = 0, PTTS_W_INT8 = 1 } ptts_weight_mode;

/* One named tensor in host memory, safetensors key + PyTorch layout
 * (keys as built at crates/pocket-tts/src/tts_model.rs:296-409). */
typedef struct {
  const char* name;
  ptts_dtype dtype;
  int32_t ndim;
  int64_t shape[4];
  const void* data;
} ptts_tensor_desc;

typedef struct {
  int32_t device;          /* CUDA ordinal */
  int32_t max_slots;       /* concurrent streams held resident */
  int32_t max_batch;       /* streams advanced per ptts_step call (<= max_slots) */
  int32_t kv_capacity;     /* FlowLM KV rows per slot (text + generated frames) */
  int32_t weight_mode;     /* ptts_weight_mode */
  int32_t use_cuda_graph;  /* capture the decode step per batch bucket */
  int32_t debug_gemm;      /* 0 = tcgen05 path; 1 = SIMT cross-check kernel (tests only) */
  int32_t reserved[9];     /* zero in production; test switches documented in ptts_internal.h */
} ptts_engine_cfg;

typedef struct ptts_engine ptts_engine;
typedef struct ptts_voice ptts_voice;

/* Per-stream generation parameters: the public fields of TTSModel the reference lets
 * callers mutate (tts_model.rs:22-49) plus the per-segment values its host code derives
 * (tts_model.rs:968-969). */
typedef struct {
  int32_t max_gen_len;       /* (words(prepared)+2)*13, tts_model.rs:968 */
  int32_t frames_after_eos;  /* estimate_frames_after_eos, tts_model.rs:1230-1237 */
  float eos_threshold;       /* config.rs:123, default -4.0 */
  float temp;                /* config.rs:120, default 0.7; used only when noise == NULL */
  uint64_t seed;             /* seed of the device generator when noise == NULL: counter-based (hash of seed, frame,
                                lane -> two uniforms -> Box-Muller), one N(0, temp) draw per latent value per frame */
  const float* noise;        /* optional [max_gen_len, 32] injected x_0 per frame, already
                                scaled by sqrt(temp) (reference draws it at flow_lm.rs:148-153) */
} ptts_stream_params;

const char* ptts_last_error(void);
int32_t ptts_abi_version(void);

/* Replaces TTSModel::load / load_with_params (tts_model.rs:59-86) after the host side has
 * read the safetensors file: uploads and repacks every tensor into kernel layouts. */
int32_t ptts_engine_create(const ptts_engine_cfg* cfg, const ptts_tensor_desc* weights, int32_t n_weights,
                           ptts_engine** out);
void ptts_engine_destroy(ptts_engine* e);

/* Sets lsd_decode_steps (TTSModel field, tts_model.rs:28) and recomputes the hoisted time
 * embeddings (SimpleMLPAdaLN::compute_time_embeddings, modules/mlp.rs:296-319). */
int32_t ptts_engine_set_lsd_steps(ptts_engine* e, int32_t lsd_steps);

/* Frames per codec pass: 1 (default), 2 or 4.  The reference decodes every latent the moment it is generated
 * (mimi.decode_from_latent inside the frame loop, tts_model.rs:1033-1047); nothing in the language model waits for that
 * audio, and the Mimi decoder is streaming (conv.rs:90-136,219-267; attention.rs:167-264), so `frames` consecutive
 * latents of the same batch may be decoded by one pass with bit-identical PCM.  With frames > 1 a step's PCM leaves the
 * device when its group is complete -- or earlier, when ptts_step_pcm asks for it, the batch composition changes, a
 * stream is opened or closed, or ptts_sync is called (a partial group is decoded then).  Re-sizes the codec scratch:
 * call it with no step in flight. */
int32_t ptts_engine_set_codec_group(ptts_engine* e, int32_t frames);

/* Replaces TTSModel::get_voice_state_from_prompt_tensor (tts_model.rs:490-501, which runs
 * run_flow_lm_prompt :580-599): FlowLM prefill over audio_prompt [T,1024] f32 (host);
 * the resulting KV snapshot is immutable and shared by every stream opened with it. */
int32_t ptts_voice_from_prompt(ptts_engine* e, const float* audio_prompt, int32_t n_rows, ptts_voice** out);

/* Replaces TTSModel::get_voice_state / get_voice_state_from_tensor (tts_model.rs:449-556): voice cloning from 24 kHz
 * mono PCM (host f32; the WAV read and resample of tts_model.rs:449-466 stay host code).  The prompt is zero-padded to
 * whole 1920-sample frames and run through the Mimi encoder (SEANetEncoder -> encoder transformer -> ConvDownsample1d,
 * models/mimi.rs:113-141) and speaker_proj_weight; the FlowLM prefill of ptts_voice_from_prompt follows.  Needs the
 * encoder tensors in the checkpoint (PTTS_ERR_STATE otherwise); up to 1024 frames.  Prompts of more than 120 frames
 * follow the reference's chunked encoding (tts_model.rs:528-541,562-577): one carried state, the downsample's
 * replicate padding restarted at every chunk.
 * ptts_audio_prompt_from_pcm stops after the conditioning rows [n_rows,1024] (the tensor the reference stores under
 * `audio_prompt`); audio_prompt_out may be NULL to query n_rows only. */
int32_t ptts_voice_from_pcm(ptts_engine* e, const float* pcm24k, int32_t n_samples, ptts_voice** out);
int32_t ptts_audio_prompt_from_pcm(ptts_engine* e, const float* pcm24k, int32_t n_samples, float* audio_prompt_out,
                                   int32_t cap_rows, int32_t* n_rows_out);
void ptts_voice_destroy(ptts_engine* e, ptts_voice* v);
int32_t ptts_voice_len(const ptts_voice* v);

/* Voice-state safetensors.  The reference's voice files hold one tensor, `audio_prompt` f32 [1,T,1024]
 * (tts_model.rs:467-487, pocket-tts-cli voice.rs:122-131); ptts_voice_load reads exactly that (F32 / F16 / BF16) and
 * runs the prefill like ptts_voice_from_prompt.  ptts_voice_save writes `audio_prompt` and, with include_kv != 0, also
 * the prefilled FlowLM KV rows as `flow_lm_kv` f16 [6,2,16,T,64] -- the reference never serialises its ModelState; a
 * file with that tensor loads without running the prefill again (the file stays a valid voice file for the reference,
 * which ignores unknown keys). */
int32_t ptts_voice_save(ptts_engine* e, const ptts_voice* v, const char* path, int32_t include_kv);
int32_t ptts_voice_load(ptts_engine* e, const char* path, ptts_voice** out);

/* Replaces load_config + the dimension checks of TTSModel::from_config (config.rs:111, tts_model.rs:182-426): reads a
 * model YAML (crates/pocket-tts/config/b6369a24.yaml) and verifies every dimension this library is compiled for
 * (FlowLM 1024 x 16 heads x 6 layers, flow head 512 x 6, Mimi 512 x 8 heads x 2 layers, context 250, SEANet ratios
 * [6,5,4], 24 kHz, 12.5 fps, 4000 bins, latent 32).  PTTS_ERR_INVALID names the first mismatch.  No device needed. */
int32_t ptts_config_check(const char* yaml_path);

/* Replaces the head of TTSModel::generate_stream_segment (tts_model.rs:935-1004) for n
 * streams at once: clone voice state, embed tokens (conditioners/text.rs:289-303), text
 * prefill (tts_model.rs:958-964).  tokens are concatenated, token_offsets has n+1 entries.
 * voices and params have n entries.  Writes n slot ids. */
int32_t ptts_streams_open(ptts_engine* e, int32_t n, ptts_voice* const* voices, const int32_t* tokens,
                          const int32_t* token_offsets, const ptts_stream_params* params, int32_t* slots_out);

/* Replaces one iteration of the frame loop (tts_model.rs:1006-1070) for n slots:
 * FlowLMModel::forward (flow_lm.rs:98-164) -> latent de-norm + Quantizer (tts_model.rs:1033-1038)
 * -> MimiModel::decode_from_latent (mimi.rs:143-157) -> EOS bookkeeping (tts_model.rs:1055-1063).
 * pcm_out [n,1920] f32 host (pinned or pageable) or NULL; finished[n] is 1 when the frame just
 * produced is the stream's last (the frame itself is valid and emitted, D2 in SURVEY 8c).
 * latent_out [n,32] and eos_logit_out [n] are optional parity taps. */
int32_t ptts_step(ptts_engine* e, const int32_t* slots, int32_t n, float* pcm_out, uint8_t* finished,
                  float* latent_out, float* eos_logit_out);

/* The same step split so the host can overlap frames: the codec half of frame n (Mimi transformer + SEANet, which
 * feeds nothing back) runs on its own CUDA stream while the language-model half of frame n+1 runs.
 *   ticket = ptts_step_begin(e, slots, n, flags)           enqueue only, returns a ticket >= 0
 *   ptts_step_flags(e, ticket, finished, latent, logit)    waits for the language-model half (needed to choose the
 *                                                          next batch); in step order
 *   ptts_step_pcm(e, ticket, pcm_out)                      waits for the codec half; may follow the next begin
 * flags: PTTS_STEP_PCM copies the frame's PCM back to the host; PTTS_STEP_AHEAD lets this step be enqueued before the
 * flags of the previous one have been fetched, so the device never waits for the host between frames.  A stream that
 * turns out to have ended on the previous step then runs one frame past its end: ptts_step_flags reports
 * PTTS_FRAME_OVERRUN for that row, the caller drops the row's latent / PCM and closes the slot (the reference stops
 * at the last frame, tts_model.rs:1055-1069; nothing of the overrun frame is ever emitted).  Needs one spare KV row
 * (kv_capacity > tokens + max_gen_len), else PTTS_ERR_CAPACITY.  At most one step ahead of unfetched flags, three
 * tickets in flight.  ptts_step == begin + flags + pcm. */
#define PTTS_STEP_PCM 1
#define PTTS_STEP_AHEAD 2
#define PTTS_STEP_PCM_I16 4  /* read the frame back as i16 instead (ptts_step_pcm_i16): the reference's wire format,
                                audio.rs:129-146 pcm_i16_le_bytes = clamp to [-1, 1], * 32767, truncating cast, packed by
                                the last SEANet conv on the device; half the bytes of the f32 frame */
#define PTTS_FRAME_OVERRUN 2 /* finished[] value: this frame lies past the stream's last one */
int64_t ptts_step_begin(ptts_engine* e, const int32_t* slots, int32_t n, int32_t flags);
int32_t ptts_step_flags(ptts_engine* e, int64_t ticket, uint8_t* finished, float* latent_out, float* eos_logit_out);
int32_t ptts_step_pcm(ptts_engine* e, int64_t ticket, float* pcm_out);
int32_t ptts_step_pcm_i16(ptts_engine* e, int64_t ticket, int16_t* pcm_out);

/* Same step with every buffer resident on the device (no host copies, no sync): used to time the
 * kernels alone.  pcm_dev may be NULL to keep the PCM in the engine's own buffer. */
int32_t ptts_step_device(ptts_engine* e, const int32_t* slots, int32_t n);
int32_t ptts_sync(ptts_engine* e);

/* Teacher forcing for parity tests: overrides the latent fed back into FlowLM at the next
 * step of `slot` (the reference feeds next_latent back at tts_model.rs:1065). */
int32_t ptts_stream_set_feedback(ptts_engine* e, int32_t slot, const float* latent32);
/* Closing is host bookkeeping only (no device synchronisation): work already enqueued for the slot completes, and a
 * later open of the same slot orders itself behind it on the device.  A slot listed by a step whose flags have not been
 * fetched cannot be closed (PTTS_ERR_STATE). */
int32_t ptts_stream_close(ptts_engine* e, int32_t slot);
int32_t ptts_streams_close(ptts_engine* e, const int32_t* slots, int32_t n);
int32_t ptts_stream_frames(ptts_engine* e, int32_t slot, int32_t* frames_out, int32_t* eos_step_out);

/* Continuous batching of many long-form requests inside the library (BASELINE configs[4]: thousands of concurrent 60 s
 * requests with [pause:Xms]).  Replaces the loop of TTSModel::generate_stream_long (tts_model.rs:1074-1127) over
 * generate_stream (tts_model.rs:894-913) for a whole population of requests: a request is an ordered list of segments,
 * PTTS_SEG_TEXT (one <= 50-token chunk: tokens + per-segment parameters, exactly what ptts_streams_open takes) or
 * PTTS_SEG_PAUSE (host silence of pause_ms, pause.rs:183-185 = ms * 24 samples).  The chunks of one request run one after
 * the other, each restarting from the voice state, while different requests fill the batch; whenever streams finish the
 * next chunks are opened together (one packed open, one batched prefill) and join the next step; the device is kept one
 * step ahead of the host (PTTS_STEP_AHEAD).  ptts_sched_run returns when every submitted request is complete; the PCM
 * of a request (all its segments concatenated, f32 or i16 as chosen at run) is then read with ptts_sched_result (copy) or ptts_sched_result_view. */
typedef struct ptts_sched ptts_sched;
#define PTTS_SEG_TEXT 0
#define PTTS_SEG_PAUSE 1
typedef struct {
  int32_t kind;               /* PTTS_SEG_TEXT | PTTS_SEG_PAUSE */
  int32_t n_tokens;           /* text */
  const int32_t* tokens;      /* text: n_tokens ids (copied at submit) */
  ptts_stream_params params;  /* text (params.noise is copied at submit when not NULL) */
  int32_t pause_ms;           /* pause */
  int32_t reserved;
} ptts_segment;
int32_t ptts_sched_create(ptts_engine* e, ptts_voice* voice, int32_t max_batch, ptts_sched** out);
void ptts_sched_destroy(ptts_sched* s);
int64_t ptts_sched_submit(ptts_sched* s, const ptts_segment* segments, int32_t n_segments); /* -> request id >= 0 */
int32_t ptts_sched_run(ptts_sched* s, int32_t pcm_i16);
int64_t ptts_sched_result_samples(const ptts_sched* s, int64_t request);
int32_t ptts_sched_result(const ptts_sched* s, int64_t request, void* pcm_out, int64_t cap_samples);
/* The same PCM without the copy: *data points at the request's samples in the scheduler's own host memory (f32 or i16 as
 * chosen at run), *n_samples is their count; valid until the next ptts_sched_run or ptts_sched_destroy.  A 60 s request is
 * 2.9 MB of i16: for 512 requests the copy above costs as much wall time as 40% of the generation itself. */
int32_t ptts_sched_result_view(const ptts_sched* s, int64_t request, const void** data, int64_t* n_samples);
int64_t ptts_sched_steps(const ptts_sched* s); /* decode steps the last run took */

#ifdef __cplusplus
}
#endif
#endif /* PTTS_H */
