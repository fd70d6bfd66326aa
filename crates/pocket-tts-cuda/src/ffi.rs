//! One declaration per symbol of include/ptts.h (the product ABI; ptts_internal.h is for tests and is not bound here).
//! Field order and widths follow the C structs exactly (`tests/c_abi/abi_check.c` pins the sizes: cfg 64 B, tensor
//! descriptor 56 B, stream params 32 B, segment 56 B).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub const PTTS_ABI_VERSION: i32 = 2;
pub const PTTS_OK: i32 = 0;
pub const PTTS_ERR_INVALID: i32 = -1;
pub const PTTS_ERR_CUDA: i32 = -2;
pub const PTTS_ERR_CAPACITY: i32 = -3;
pub const PTTS_ERR_STATE: i32 = -4;
pub const PTTS_F32: i32 = 0;
pub const PTTS_BF16: i32 = 1;
pub const PTTS_F16: i32 = 2;
pub const PTTS_W_F16: i32 = 0;
pub const PTTS_W_INT8: i32 = 1;
pub const PTTS_STEP_PCM: i32 = 1;
pub const PTTS_STEP_AHEAD: i32 = 2;
pub const PTTS_STEP_PCM_I16: i32 = 4;
pub const PTTS_FRAME_OVERRUN: u8 = 2;
pub const PTTS_SEG_TEXT: i32 = 0;
pub const PTTS_SEG_PAUSE: i32 = 1;

#[repr(C)]
pub struct ptts_tensor_desc {
    pub name: *const c_char,
    pub dtype: i32,
    pub ndim: i32,
    pub shape: [i64; 4],
    pub data: *const c_void,
}

#[repr(C)]
#[derive(Default)]
pub struct ptts_engine_cfg {
    pub device: i32,
    pub max_slots: i32,
    pub max_batch: i32,
    pub kv_capacity: i32,
    pub weight_mode: i32,
    pub use_cuda_graph: i32,
    pub debug_gemm: i32,
    pub reserved: [i32; 9],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ptts_stream_params {
    pub max_gen_len: i32,
    pub frames_after_eos: i32,
    pub eos_threshold: f32,
    pub temp: f32,
    pub seed: u64,
    pub noise: *const f32,
}

#[repr(C)]
pub struct ptts_segment {
    pub kind: i32,
    pub n_tokens: i32,
    pub tokens: *const i32,
    pub params: ptts_stream_params,
    pub pause_ms: i32,
    pub reserved: i32,
}

pub enum ptts_engine {}
pub enum ptts_voice {}
pub enum ptts_sched {}

extern "C" {
    pub fn ptts_last_error() -> *const c_char;
    pub fn ptts_abi_version() -> i32;

    pub fn ptts_engine_create(cfg: *const ptts_engine_cfg, weights: *const ptts_tensor_desc, n_weights: i32, out: *mut *mut ptts_engine) -> i32;
    pub fn ptts_engine_destroy(e: *mut ptts_engine);
    pub fn ptts_engine_set_lsd_steps(e: *mut ptts_engine, lsd_steps: i32) -> i32;
    pub fn ptts_engine_set_codec_group(e: *mut ptts_engine, frames: i32) -> i32;
    pub fn ptts_config_check(yaml_path: *const c_char) -> i32;

    pub fn ptts_voice_from_prompt(e: *mut ptts_engine, audio_prompt: *const f32, n_rows: i32, out: *mut *mut ptts_voice) -> i32;
    pub fn ptts_voice_from_pcm(e: *mut ptts_engine, pcm24k: *const f32, n_samples: i32, out: *mut *mut ptts_voice) -> i32;
    pub fn ptts_audio_prompt_from_pcm(e: *mut ptts_engine, pcm24k: *const f32, n_samples: i32, audio_prompt_out: *mut f32, cap_rows: i32,
                                      n_rows_out: *mut i32) -> i32;
    pub fn ptts_voice_destroy(e: *mut ptts_engine, v: *mut ptts_voice);
    pub fn ptts_voice_len(v: *const ptts_voice) -> i32;
    pub fn ptts_voice_save(e: *mut ptts_engine, v: *const ptts_voice, path: *const c_char, include_kv: i32) -> i32;
    pub fn ptts_voice_load(e: *mut ptts_engine, path: *const c_char, out: *mut *mut ptts_voice) -> i32;

    pub fn ptts_streams_open(e: *mut ptts_engine, n: i32, voices: *const *mut ptts_voice, tokens: *const i32, token_offsets: *const i32,
                             params: *const ptts_stream_params, slots_out: *mut i32) -> i32;
    pub fn ptts_step(e: *mut ptts_engine, slots: *const i32, n: i32, pcm_out: *mut f32, finished: *mut u8, latent_out: *mut f32,
                     eos_logit_out: *mut f32) -> i32;
    pub fn ptts_step_begin(e: *mut ptts_engine, slots: *const i32, n: i32, flags: i32) -> i64;
    pub fn ptts_step_flags(e: *mut ptts_engine, ticket: i64, finished: *mut u8, latent_out: *mut f32, eos_logit_out: *mut f32) -> i32;
    pub fn ptts_step_pcm(e: *mut ptts_engine, ticket: i64, pcm_out: *mut f32) -> i32;
    pub fn ptts_step_pcm_i16(e: *mut ptts_engine, ticket: i64, pcm_out: *mut i16) -> i32;
    pub fn ptts_step_device(e: *mut ptts_engine, slots: *const i32, n: i32) -> i32;
    pub fn ptts_sync(e: *mut ptts_engine) -> i32;
    pub fn ptts_stream_set_feedback(e: *mut ptts_engine, slot: i32, latent32: *const f32) -> i32;
    pub fn ptts_stream_close(e: *mut ptts_engine, slot: i32) -> i32;
    pub fn ptts_streams_close(e: *mut ptts_engine, slots: *const i32, n: i32) -> i32;
    pub fn ptts_stream_frames(e: *mut ptts_engine, slot: i32, frames_out: *mut i32, eos_step_out: *mut i32) -> i32;

    pub fn ptts_sched_create(e: *mut ptts_engine, voice: *mut ptts_voice, max_batch: i32, out: *mut *mut ptts_sched) -> i32;
    pub fn ptts_sched_destroy(s: *mut ptts_sched);
    pub fn ptts_sched_submit(s: *mut ptts_sched, segments: *const ptts_segment, n_segments: i32) -> i64;
    pub fn ptts_sched_run(s: *mut ptts_sched, pcm_i16: i32) -> i32;
    pub fn ptts_sched_result_samples(s: *const ptts_sched, request: i64) -> i64;
    pub fn ptts_sched_result(s: *const ptts_sched, request: i64, pcm_out: *mut c_void, cap_samples: i64) -> i32;
    pub fn ptts_sched_result_view(s: *const ptts_sched, request: i64, data: *mut *const c_void, n_samples: *mut i64) -> i32;
    pub fn ptts_sched_steps(s: *const ptts_sched) -> i64;
}
