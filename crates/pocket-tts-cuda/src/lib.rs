//! `pocket-tts-cuda`: the public surface of the reference's `TTSModel` (crates/pocket-tts/src/tts_model.rs:59-1131, re-exported
//! at lib.rs:15-18) with the generation hot path on a B200 through libptts_cuda.so.
//!
//! What stays host code, taken from the `pocket-tts` crate unchanged: text preparation (`prepare_text_prompt`,
//! `split_into_best_sentences`, `estimate_frames_after_eos`, tts_model.rs:604-684,1194-1237), the tokenizer
//! (conditioners/text.rs), pause parsing (pause.rs), WAV read / resample (audio.rs).  What moves to the device: voice
//! prefill, text prefill, the frame loop (FlowLM step, flow head, Mimi decoder), EOS bookkeeping, the Mimi encoder for
//! voice cloning, i16 packing.
pub mod ffi;

use anyhow::{anyhow, bail, Result};
use candle_core::{Device, Tensor};
use std::ffi::{CStr, CString};
use std::path::Path;
use std::sync::{Arc, Mutex};

pub use pocket_tts::pause::{parse_text_with_pauses, silence_samples};
use pocket_tts::tts_model::{estimate_frames_after_eos, prepare_text_prompt};

const FRAME: usize = 1920;

fn check(status: i32) -> Result<()> {
    if status >= 0 {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(ffi::ptts_last_error()) }.to_string_lossy().into_owned();
    Err(anyhow!("ptts error {status}: {msg}"))
}

/// One engine = one GPU.  The engine is not re-entrant: every call goes through this mutex (the reference serialises
/// generation behind one lock as well, pocket-tts-cli server/state.rs:67-69).
struct EngineHandle(*mut ffi::ptts_engine);
unsafe impl Send for EngineHandle {}
impl Drop for EngineHandle {
    fn drop(&mut self) {
        unsafe { ffi::ptts_engine_destroy(self.0) }
    }
}

/// The reference's `ModelState` (voice_state.rs:7) for a voice: an immutable KV snapshot on the device, shared by clones
/// (the reference's `voice_state.clone()` at tts_model.rs:940 is shallow too).
pub struct VoiceHandle {
    engine: Arc<Mutex<EngineHandle>>,
    ptr: *mut ffi::ptts_voice,
}
unsafe impl Send for VoiceHandle {}
unsafe impl Sync for VoiceHandle {}
impl Drop for VoiceHandle {
    fn drop(&mut self) {
        let e = self.engine.lock().unwrap();
        unsafe { ffi::ptts_voice_destroy(e.0, self.ptr) }
    }
}
pub type ModelState = Arc<VoiceHandle>;

#[derive(Clone)]
pub struct TTSModel {
    engine: Arc<Mutex<EngineHandle>>,
    conditioner: Arc<pocket_tts::conditioners::text::LUTConditioner>,
    // public fields callers mutate, exactly as in the reference (tts_model.rs:22-49)
    pub temp: f32,
    pub lsd_decode_steps: usize,
    pub eos_threshold: f32,
    pub noise_clamp: Option<f32>,
    pub sample_rate: usize,
    pub dim: usize,
    pub ldim: usize,
    pub device: Device,
}

impl TTSModel {
    /// tts_model.rs:59
    pub fn load(variant: &str) -> Result<Self> {
        Self::load_with_params(variant, 0.7, 1, -4.0) // config.rs:118-124
    }

    /// tts_model.rs:69: reads config/{variant}.yaml and the safetensors file like `from_config` (:182-203), verifies the
    /// dimensions against the compiled kernels, hands every tensor to the engine (keys of :296-409) which repacks them.
    pub fn load_with_params(variant: &str, temp: f32, lsd_decode_steps: usize, eos_threshold: f32) -> Result<Self> {
        let config_path = pocket_tts::tts_model::find_config_path(variant)?;
        check(unsafe { ffi::ptts_config_check(CString::new(config_path.to_string_lossy().as_bytes())?.as_ptr()) })?;
        let config = pocket_tts::config::load_config(&config_path)?;
        let weights_path = pocket_tts::weights::download_if_necessary(&config.weights_path)?;
        let file = std::fs::File::open(&weights_path)?;
        let mmap = unsafe { memmap2::Mmap::map(&file)? };
        let st = safetensors::SafeTensors::deserialize(&mmap)?;
        let names: Vec<CString> = st.names().iter().map(|n| CString::new(n.as_str()).unwrap()).collect();
        let mut descs = Vec::with_capacity(names.len());
        for (name, cname) in st.names().iter().zip(&names) {
            let t = st.tensor(name)?;
            let dtype = match t.dtype() {
                safetensors::Dtype::F32 => ffi::PTTS_F32,
                safetensors::Dtype::BF16 => ffi::PTTS_BF16,
                safetensors::Dtype::F16 => ffi::PTTS_F16,
                other => bail!("tensor {name}: unsupported dtype {other:?}"),
            };
            let mut shape = [0i64; 4];
            for (d, s) in shape.iter_mut().zip(t.shape()) {
                *d = *s as i64;
            }
            descs.push(ffi::ptts_tensor_desc { name: cname.as_ptr(), dtype, ndim: t.shape().len() as i32, shape, data: t.data().as_ptr().cast() });
        }
        let cfg = ffi::ptts_engine_cfg { device: 0, max_slots: 64, max_batch: 64, kv_capacity: 1024, weight_mode: ffi::PTTS_W_F16,
                                         use_cuda_graph: 1, ..Default::default() };
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::ptts_engine_create(&cfg, descs.as_ptr(), descs.len() as i32, &mut raw) })?;
        let model = Self {
            engine: Arc::new(Mutex::new(EngineHandle(raw))),
            conditioner: Arc::new(pocket_tts::conditioners::text::LUTConditioner::tokenizer_only(&config)?),
            temp, lsd_decode_steps, eos_threshold, noise_clamp: None,
            sample_rate: 24000, dim: 1024, ldim: 32, device: Device::Cpu,
        };
        Ok(model)
    }

    // ---- voice state -----------------------------------------------------------------------------------------------
    /// tts_model.rs:467: a voice-state safetensors file, key `audio_prompt` f32 [1,T,1024] (+ optional KV snapshot).
    pub fn get_voice_state_from_prompt_file<P: AsRef<Path>>(&self, path: P) -> Result<ModelState> {
        let c = CString::new(path.as_ref().to_string_lossy().as_bytes())?;
        let e = self.engine.lock().unwrap();
        let mut v = std::ptr::null_mut();
        check(unsafe { ffi::ptts_voice_load(e.0, c.as_ptr(), &mut v) })?;
        Ok(Arc::new(VoiceHandle { engine: self.engine.clone(), ptr: v }))
    }

    /// tts_model.rs:490
    pub fn get_voice_state_from_prompt_tensor(&self, prompt: &Tensor) -> Result<ModelState> {
        let rows: Vec<f32> = prompt.flatten_all()?.to_vec1()?;
        let e = self.engine.lock().unwrap();
        let mut v = std::ptr::null_mut();
        check(unsafe { ffi::ptts_voice_from_prompt(e.0, rows.as_ptr(), (rows.len() / self.dim) as i32, &mut v) })?;
        Ok(Arc::new(VoiceHandle { engine: self.engine.clone(), ptr: v }))
    }

    /// tts_model.rs:504: audio [1, 1, T] at the model's sample rate -> Mimi encoder + speaker projection + prefill on the GPU.
    pub fn get_voice_state_from_tensor(&self, audio: &Tensor) -> Result<ModelState> {
        let pcm: Vec<f32> = audio.flatten_all()?.to_vec1()?;
        let e = self.engine.lock().unwrap();
        let mut v = std::ptr::null_mut();
        check(unsafe { ffi::ptts_voice_from_pcm(e.0, pcm.as_ptr(), pcm.len() as i32, &mut v) })?;
        Ok(Arc::new(VoiceHandle { engine: self.engine.clone(), ptr: v }))
    }

    /// tts_model.rs:449: WAV read and resampling stay host code (audio.rs)
    pub fn get_voice_state<P: AsRef<Path>>(&self, audio_path: P) -> Result<ModelState> {
        let (audio, sr) = pocket_tts::audio::read_wav(audio_path)?;
        let audio = if sr != self.sample_rate as u32 { pocket_tts::audio::resample(&audio, sr, self.sample_rate as u32)? } else { audio };
        self.get_voice_state_from_tensor(&audio.unsqueeze(0)?)
    }

    /// New with this crate: writes `audio_prompt` (and, with `include_kv`, the prefilled KV rows) as a voice-state file.
    pub fn save_voice_state<P: AsRef<Path>>(&self, voice: &ModelState, path: P, include_kv: bool) -> Result<()> {
        let c = CString::new(path.as_ref().to_string_lossy().as_bytes())?;
        let e = self.engine.lock().unwrap();
        check(unsafe { ffi::ptts_voice_save(e.0, voice.ptr, c.as_ptr(), include_kv as i32) })
    }

    // ---- generation ------------------------------------------------------------------------------------------------
    /// tts_model.rs:687
    pub fn generate(&self, text: &str, voice_state: &ModelState) -> Result<Tensor> {
        let frames: Vec<Tensor> = self.generate_stream(text, voice_state).collect::<Result<_>>()?;
        if frames.is_empty() {
            bail!("No audio generated"); // tts_model.rs:695-697
        }
        Ok(Tensor::cat(&frames, 2)?.squeeze(0)?)
    }

    /// tts_model.rs:894: chunks of <= 50 tokens, each restarting from the voice state, frames in order, errors in-band.
    pub fn generate_stream<'a>(&'a self, text: &str, voice_state: &ModelState) -> Box<dyn Iterator<Item = Result<Tensor>> + 'a> {
        let chunks = self.split_into_best_sentences(text);
        let voice = voice_state.clone();
        Box::new(chunks.into_iter().flat_map(move |chunk| SegmentIter::open(self, chunk, voice.clone())))
    }

    /// tts_model.rs:1074: `[pause:Xms]` markers and natural pauses become host zeros between text segments.
    pub fn generate_stream_long<'a>(&'a self, text: &str, voice_state: &ModelState) -> Box<dyn Iterator<Item = Result<Tensor>> + 'a> {
        let voice = voice_state.clone();
        let sr = self.sample_rate as u32;
        Box::new(parse_text_with_pauses(text).into_iter().flat_map(move |seg| -> Box<dyn Iterator<Item = Result<Tensor>> + 'a> {
            match seg {
                pocket_tts::pause::Segment::Text(t) => self.generate_stream(&t, &voice),
                pocket_tts::pause::Segment::Pause(ms) => {
                    let n = silence_samples(ms, sr);
                    Box::new(std::iter::once(Tensor::zeros((1, 1, n), candle_core::DType::F32, &Device::Cpu).map_err(Into::into)))
                }
            }
        }))
    }

    pub fn split_into_best_sentences(&self, text: &str) -> Vec<String> {
        pocket_tts::tts_model::split_into_best_sentences_with(text, |s| self.conditioner.count_tokens(s)) // tts_model.rs:604-684
    }

    /// Frames per Mimi-decoder pass (1, 2 or 4; `ptts_engine_set_codec_group`).  The reference decodes every latent inside the
    /// frame loop (tts_model.rs:1033-1047); the streaming decoder gives bit-identical PCM for a group, the PCM of a frame then
    /// leaves the device with its group.  No counterpart in the reference; call with no iterator alive.
    pub fn set_codec_group(&self, frames: usize) -> Result<()> {
        let e = self.engine.lock().unwrap();
        check(unsafe { ffi::ptts_engine_set_codec_group(e.0, frames as i32) })
    }

    fn sync_params(&self) -> Result<()> {
        let e = self.engine.lock().unwrap();
        check(unsafe { ffi::ptts_engine_set_lsd_steps(e.0, self.lsd_decode_steps as i32) })
    }
}

/// `generate_stream_segment` (tts_model.rs:935-1071) as an iterator over the C ABI: open (text prefill), then per frame
/// `ptts_step_begin` of frame n+1 ahead of the flags of frame n (`PTTS_STEP_AHEAD`), `ptts_step_flags` of frame n,
/// `ptts_step_pcm` of frame n.  Stops after the frame whose `finished` flag is 1; the frame enqueued ahead of an EOS
/// ending is retired unseen (`PTTS_FRAME_OVERRUN`).  Dropping the iterator retires what is in flight and closes the slot.
pub struct SegmentIter<'a> {
    model: &'a TTSModel,
    _voice: ModelState,
    slot: i32,
    max_gen_len: usize,
    issued: usize,
    current: Option<i64>, // ticket whose frame is delivered next
    ahead: Option<i64>,   // ticket enqueued ahead of it
    done: bool,
    failed: Option<anyhow::Error>,
}

impl<'a> SegmentIter<'a> {
    pub fn open(model: &'a TTSModel, text: String, voice: ModelState) -> Self {
        let mut it = Self { model, _voice: voice.clone(), slot: -1, max_gen_len: 0, issued: 0, current: None, ahead: None, done: false, failed: None };
        if let Err(e) = it.start(&text, &voice) {
            it.failed = Some(e);
        }
        it
    }

    fn start(&mut self, text: &str, voice: &ModelState) -> Result<()> {
        self.model.sync_params()?;
        let prepared = prepare_text_prompt(text); // tts_model.rs:944,1194-1227
        let tokens: Vec<i32> = self.model.conditioner.token_ids(&prepared)?.into_iter().map(|t| t as i32).collect();
        self.max_gen_len = (prepared.split_whitespace().count() + 2) * 13; // tts_model.rs:968
        let noise: Option<Vec<f32>> = self.model.noise_clamp.map(|limit| clamped_noise(self.max_gen_len * 32, self.model.temp, limit));
        let params = ffi::ptts_stream_params {
            max_gen_len: self.max_gen_len as i32,
            frames_after_eos: estimate_frames_after_eos(text) as i32, // tts_model.rs:969,1230-1237
            eos_threshold: self.model.eos_threshold,
            temp: self.model.temp,
            seed: rand::random(),
            noise: noise.as_ref().map_or(std::ptr::null(), |n| n.as_ptr()),
        };
        let offsets = [0i32, tokens.len() as i32];
        let voices = [voice.ptr];
        let e = self.model.engine.lock().unwrap();
        check(unsafe { ffi::ptts_streams_open(e.0, 1, voices.as_ptr(), tokens.as_ptr(), offsets.as_ptr(), &params, &mut self.slot) })?;
        self.current = Some(self.begin(e.0, false)?);
        Ok(())
    }

    fn begin(&mut self, e: *mut ffi::ptts_engine, ahead: bool) -> Result<i64> {
        let flags = ffi::PTTS_STEP_PCM | if ahead { ffi::PTTS_STEP_AHEAD } else { 0 };
        let t = unsafe { ffi::ptts_step_begin(e, &self.slot, 1, flags) };
        check(t.min(0) as i32)?;
        self.issued += 1;
        Ok(t)
    }

    fn frame(&mut self) -> Result<Option<Tensor>> {
        let Some(ticket) = self.current else { return Ok(None) };
        let guard = self.model.engine.lock().unwrap();
        let e = guard.0;
        // frame n+1 goes out before frame n's flags come back, unless frame n is the max_gen_len-th (then it is the last)
        if self.ahead.is_none() && self.issued < self.max_gen_len {
            match self.begin(e, true) {
                Ok(t) => self.ahead = Some(t),
                Err(err) if err.to_string().starts_with("ptts error -3") => {} // no spare KV row: begin after the flags instead
                Err(err) => return Err(err),
            }
        }
        let mut fin = 0u8;
        check(unsafe { ffi::ptts_step_flags(e, ticket, &mut fin, std::ptr::null_mut(), std::ptr::null_mut()) })?;
        if self.ahead.is_none() && fin == 0 && self.issued < self.max_gen_len {
            self.ahead = Some(self.begin(e, false)?);
        }
        let mut pcm = vec![0f32; FRAME];
        check(unsafe { ffi::ptts_step_pcm(e, ticket, pcm.as_mut_ptr()) })?;
        self.current = if fin != 0 { None } else { self.ahead.take() };
        if fin != 0 {
            self.done = true;
        }
        Ok(Some(Tensor::from_vec(pcm, (1, 1, FRAME), &Device::Cpu)?))
    }

    fn retire(&mut self) {
        let Ok(guard) = self.model.engine.lock() else { return };
        for t in [self.current.take(), self.ahead.take()].into_iter().flatten() {
            unsafe {
                ffi::ptts_step_flags(guard.0, t, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut());
                ffi::ptts_step_pcm(guard.0, t, std::ptr::null_mut());
            }
        }
        if self.slot >= 0 {
            unsafe { ffi::ptts_stream_close(guard.0, self.slot) };
            self.slot = -1;
        }
    }
}

impl Iterator for SegmentIter<'_> {
    type Item = Result<Tensor>;
    fn next(&mut self) -> Option<Self::Item> {
        if let Some(e) = self.failed.take() {
            self.done = true;
            return Some(Err(e)); // errors in-band, like tts_model.rs:1028-1030
        }
        if self.done && self.current.is_none() {
            return None;
        }
        match self.frame() {
            Ok(Some(t)) => Some(Ok(t)),
            Ok(None) => None,
            Err(e) => {
                self.done = true;
                self.current = None;
                Some(Err(e))
            }
        }
    }
}

impl Drop for SegmentIter<'_> {
    fn drop(&mut self) {
        self.retire();
    }
}

/// models/flow_lm.rs:39-65 with `noise_clamp = Some(limit)`: N(0, temp) by rejection sampling, handed to the engine as the
/// stream's injected noise (the device generator serves the unclamped case).
fn clamped_noise(n: usize, temp: f32, limit: f32) -> Vec<f32> {
    use rand::Rng;
    let std = temp.sqrt();
    let mut rng = rand::thread_rng();
    let mut out = Vec::with_capacity(n);
    while out.len() < n {
        // Box-Muller
        let (u1, u2): (f32, f32) = (rng.gen_range(f32::MIN_POSITIVE..1.0), rng.gen());
        let v = (-2.0 * u1.ln()).sqrt() * (2.0 * std::f32::consts::PI * u2).cos() * std;
        if v.abs() <= limit {
            out.push(v);
        }
    }
    out
}

/// BASELINE configs[4]: a population of long-form requests through the library's own continuous-batching scheduler
/// (`ptts_sched_*`): chunk every request like `generate_stream_long`, submit, run, collect i16 PCM.
pub fn generate_many_long(model: &TTSModel, texts: &[&str], voice: &ModelState) -> Result<Vec<Vec<i16>>> {
    model.sync_params()?;
    let e = model.engine.lock().unwrap();
    let mut sched = std::ptr::null_mut();
    check(unsafe { ffi::ptts_sched_create(e.0, voice.ptr, 0, &mut sched) })?;
    let run = || -> Result<Vec<Vec<i16>>> {
        for text in texts {
            let mut keep: Vec<Vec<i32>> = Vec::new();
            let mut segs: Vec<ffi::ptts_segment> = Vec::new();
            let no_params = ffi::ptts_stream_params { max_gen_len: 0, frames_after_eos: 0, eos_threshold: 0.0, temp: 0.0, seed: 0, noise: std::ptr::null() };
            for seg in parse_text_with_pauses(text) {
                match seg {
                    pocket_tts::pause::Segment::Pause(ms) => segs.push(ffi::ptts_segment { kind: ffi::PTTS_SEG_PAUSE, n_tokens: 0, tokens: std::ptr::null(),
                                                                                      params: no_params, pause_ms: ms as i32, reserved: 0 }),
                    pocket_tts::pause::Segment::Text(t) => {
                        for chunk in model.split_into_best_sentences(&t) {
                            let prepared = prepare_text_prompt(&chunk);
                            keep.push(model.conditioner.token_ids(&prepared)?.into_iter().map(|x| x as i32).collect());
                            let tok = keep.last().unwrap();
                            let params = ffi::ptts_stream_params { max_gen_len: ((prepared.split_whitespace().count() + 2) * 13) as i32,
                                frames_after_eos: estimate_frames_after_eos(&chunk) as i32, eos_threshold: model.eos_threshold, temp: model.temp,
                                seed: rand::random(), noise: std::ptr::null() };
                            segs.push(ffi::ptts_segment { kind: ffi::PTTS_SEG_TEXT, n_tokens: tok.len() as i32, tokens: tok.as_ptr(), params, pause_ms: 0, reserved: 0 });
                        }
                    }
                }
            }
            let id = unsafe { ffi::ptts_sched_submit(sched, segs.as_ptr(), segs.len() as i32) }; // tokens are copied at submit
            check(id.min(0) as i32)?;
        }
        check(unsafe { ffi::ptts_sched_run(sched, 1) })?;
        (0..texts.len() as i64).map(|r| {
            // the samples sit in the scheduler's own host buffer: one copy into the Vec the caller owns, no zero-fill first
            let (mut data, mut n) = (std::ptr::null::<std::ffi::c_void>(), 0i64);
            check(unsafe { ffi::ptts_sched_result_view(sched, r, &mut data, &mut n) })?;
            if n <= 0 || data.is_null() { return Ok(Vec::new()); }
            Ok(unsafe { std::slice::from_raw_parts(data.cast::<i16>(), n as usize) }.to_vec())
        }).collect()
    };
    let out = run();
    unsafe { ffi::ptts_sched_destroy(sched) };
    out
}
