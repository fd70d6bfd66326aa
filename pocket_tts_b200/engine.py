"""Python view of the C-ABI engine (numpy in / numpy out).  Used by tests and bench.py; a Rust or
C++ host binds the same symbols."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import EngineCfg, Segment, StreamParams, TensorDesc, check

FRAME = 1920
LDIM = 32


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class StreamSpec:
    tokens: np.ndarray
    max_gen_len: int
    frames_after_eos: int = 3
    eos_threshold: float = -4.0
    temp: float = 0.7
    seed: int = 0
    noise: np.ndarray | None = None  # [max_gen_len, 32], already scaled by sqrt(temp)


class Voice:
    def __init__(self, engine: "Engine", handle):
        self._engine, self._h = engine, handle

    def __len__(self):
        return int(_lib.lib().ptts_voice_len(self._h))

    def close(self):
        if self._h:
            _lib.lib().ptts_voice_destroy(self._engine._h, self._h)
            self._h = None

    def save(self, path, include_kv: bool = False):
        """Voice-state safetensors (ptts_voice_save): `audio_prompt` f32 [1,T,1024] as the reference reads it
        (tts_model.rs:467-487); include_kv adds the prefilled FlowLM KV rows so that a later load skips the prefill."""
        check(_lib.lib().ptts_voice_save(self._engine._h, self._h, str(path).encode(), int(include_kv)))


class Engine:
    def __init__(self, weights: dict[str, np.ndarray], device: int = 0, max_slots: int = 64, max_batch: int | None = None,
                 kv_capacity: int = 1024, debug_gemm: int = 0, gemm_mode: int = 0, cuda_graph: bool = True, int8_weights: bool = False,
                 int8_storage: bool = True, lm_step_kernel: bool | None = None, codec_group: int | None = None,
                 gemv_off: bool = False):
        L = _lib.lib()
        self._keep = []
        descs = (TensorDesc * len(weights))()
        for i, (name, arr) in enumerate(weights.items()):
            a = np.ascontiguousarray(arr, dtype=np.float32)
            self._keep.append(a)
            descs[i].name = name.encode()
            descs[i].dtype = 0
            descs[i].ndim = a.ndim
            for j, s in enumerate(a.shape):
                descs[i].shape[j] = s
            descs[i].data = a.ctypes.data
        cfg = EngineCfg()
        cfg.device, cfg.max_slots = device, max_slots
        cfg.max_batch = max_batch or max_slots
        cfg.kv_capacity, cfg.weight_mode, cfg.use_cuda_graph, cfg.debug_gemm = kv_capacity, int(int8_weights), int(cuda_graph), debug_gemm
        cfg.reserved[0] = gemm_mode
        cfg.reserved[1] = 1 if gemv_off else 0      # test hook: Linear layers of 1-4 rows on the tensor-core path instead of csrc/gemv.cuh
        cfg.reserved[7] = 0 if int8_storage else 1  # test hook: int8 mode streaming f16 copies of the codes instead of bytes
        # the persistent FlowLM step kernel (csrc/lm_step.cuh): None = the library's default, True / False = force
        cfg.reserved[8] = 0 if lm_step_kernel is None else (2 if lm_step_kernel else 1)
        h = C.c_void_p()
        check(L.ptts_engine_create(C.byref(cfg), descs, len(weights), C.byref(h)))
        self._h = h
        self._keep = []
        self.max_batch = cfg.max_batch
        if codec_group is not None:
            self.set_codec_group(codec_group)

    def close(self):
        if self._h:
            _lib.lib().ptts_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_lsd_steps(self, n: int):
        check(_lib.lib().ptts_engine_set_lsd_steps(self._h, n))

    def set_codec_group(self, frames: int):
        """Frames per codec pass (ptts_engine_set_codec_group): 1, 2 or 4; same PCM, the codec's launches paid once per group."""
        check(_lib.lib().ptts_engine_set_codec_group(self._h, frames))

    def voice_from_prompt(self, audio_prompt: np.ndarray) -> Voice:
        a = np.ascontiguousarray(audio_prompt, dtype=np.float32).reshape(-1, 1024)
        h = C.c_void_p()
        check(_lib.lib().ptts_voice_from_prompt(self._h, _ptr(a), a.shape[0], C.byref(h)))
        return Voice(self, h)

    def voice_from_pcm(self, pcm24k: np.ndarray) -> Voice:
        """Voice cloning from 24 kHz mono PCM (ptts_voice_from_pcm): Mimi encoder + speaker projection + FlowLM prefill."""
        a = np.ascontiguousarray(pcm24k, dtype=np.float32).reshape(-1)
        h = C.c_void_p()
        check(_lib.lib().ptts_voice_from_pcm(self._h, _ptr(a), a.shape[0], C.byref(h)))
        return Voice(self, h)

    def voice_load(self, path) -> Voice:
        """ptts_voice_load: `audio_prompt` from a voice-state safetensors file (+ the KV snapshot when the file has one)."""
        h = C.c_void_p()
        check(_lib.lib().ptts_voice_load(self._h, str(path).encode(), C.byref(h)))
        return Voice(self, h)

    def audio_prompt_from_pcm(self, pcm24k: np.ndarray) -> np.ndarray:
        """The conditioning rows [frames, 1024] of a PCM prompt (what the reference stores as `audio_prompt`)."""
        a = np.ascontiguousarray(pcm24k, dtype=np.float32).reshape(-1)
        rows = (a.shape[0] + FRAME - 1) // FRAME   # the prompt is zero-padded to whole frames (tts_model.rs:514-527)
        out = np.empty((max(rows, 1), 1024), np.float32)
        n = C.c_int32()
        check(_lib.lib().ptts_audio_prompt_from_pcm(self._h, _ptr(a), a.shape[0], _ptr(out), out.shape[0], C.byref(n)))
        return out[:n.value]

    def open_streams(self, voices: list[Voice], specs: list[StreamSpec]) -> np.ndarray:
        n = len(specs)
        toks = np.concatenate([np.asarray(s.tokens, dtype=np.int32) for s in specs]) if n else np.zeros(0, np.int32)
        toks = np.ascontiguousarray(toks, dtype=np.int32)
        offs = np.zeros(n + 1, np.int32)
        offs[1:] = np.cumsum([len(s.tokens) for s in specs])
        params = (StreamParams * n)()
        keep = []
        for i, s in enumerate(specs):
            params[i].max_gen_len, params[i].frames_after_eos = s.max_gen_len, s.frames_after_eos
            params[i].eos_threshold, params[i].temp, params[i].seed = s.eos_threshold, s.temp, s.seed
            if s.noise is not None:
                nz = np.ascontiguousarray(s.noise, dtype=np.float32)
                assert nz.shape == (s.max_gen_len, LDIM), nz.shape
                keep.append(nz)
                params[i].noise = nz.ctypes.data
        vh = (C.c_void_p * n)(*[v._h for v in voices])
        slots = np.zeros(n, np.int32)
        check(_lib.lib().ptts_streams_open(self._h, n, vh, _ptr(toks), _ptr(offs), params, _ptr(slots)))
        return slots

    def step(self, slots: np.ndarray, want_pcm: bool = True):
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        n = len(slots)
        pcm = np.empty((n, FRAME), np.float32) if want_pcm else None
        fin = np.zeros(n, np.uint8)
        lat = np.empty((n, LDIM), np.float32)
        logit = np.empty(n, np.float32)
        check(_lib.lib().ptts_step(self._h, _ptr(slots), n, _ptr(pcm), _ptr(fin), _ptr(lat), _ptr(logit)))
        return pcm, fin.astype(bool), lat, logit

    # ---- pipelined form: begin -> flags -> (next begin) -> pcm
    def step_begin(self, slots: np.ndarray, want_pcm: bool = True, ahead: bool = False, i16: bool = False) -> int:
        """Enqueue one frame.  ahead=True (PTTS_STEP_AHEAD): the previous step's flags need not have been fetched yet;
        rows whose stream ended on that previous step come back from step_flags with fin == 2 (frame past the end).
        i16=True (PTTS_STEP_PCM_I16): the frame is read back as i16 with step_pcm_i16."""
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        t = int(_lib.lib().ptts_step_begin(self._h, _ptr(slots), len(slots), (4 if i16 else (1 if want_pcm else 0)) | (2 if ahead else 0)))
        check(t)
        self._pending_n = getattr(self, "_pending_n", {})
        self._pending_n[t] = len(slots)
        return t

    def step_flags(self, ticket: int):
        n = self._pending_n[ticket]
        fin = np.zeros(n, np.uint8)
        lat = np.empty((n, LDIM), np.float32)
        logit = np.empty(n, np.float32)
        check(_lib.lib().ptts_step_flags(self._h, ticket, _ptr(fin), _ptr(lat), _ptr(logit)))
        self.last_overrun = fin == 2  # rows of a step enqueued ahead whose stream had already ended: drop their frame
        return fin.astype(bool), lat, logit

    def step_pcm(self, ticket: int, want: bool = True):
        n = self._pending_n.pop(ticket)
        pcm = np.empty((n, FRAME), np.float32) if want else None
        check(_lib.lib().ptts_step_pcm(self._h, ticket, _ptr(pcm)))
        return pcm

    def step_pcm_i16(self, ticket: int) -> np.ndarray:
        n = self._pending_n.pop(ticket)
        pcm = np.empty((n, FRAME), np.int16)
        check(_lib.lib().ptts_step_pcm_i16(self._h, ticket, _ptr(pcm)))
        return pcm

    def step_device(self, slots: np.ndarray):
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        check(_lib.lib().ptts_step_device(self._h, _ptr(slots), len(slots)))

    def step_timed(self, slots: np.ndarray) -> np.ndarray:
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        ms = np.zeros(8, np.float32)
        check(_lib.lib().ptts_step_timed(self._h, _ptr(slots), len(slots), _ptr(ms)))
        return ms

    def sync(self):
        check(_lib.lib().ptts_sync(self._h))

    def set_feedback(self, slot: int, latent: np.ndarray):
        a = np.ascontiguousarray(latent, dtype=np.float32).reshape(LDIM)
        check(_lib.lib().ptts_stream_set_feedback(self._h, int(slot), _ptr(a)))

    def close_stream(self, slot: int):
        check(_lib.lib().ptts_stream_close(self._h, int(slot)))

    def close_streams(self, slots):
        a = np.ascontiguousarray(slots, dtype=np.int32)
        check(_lib.lib().ptts_streams_close(self._h, _ptr(a), len(a)))

    def f16_overflow_count(self) -> int:
        v = C.c_int64()
        check(_lib.lib().ptts_debug_f16_overflow(self._h, C.byref(v)))
        return int(v.value)

    def stream_frames(self, slot: int) -> tuple[int, int]:
        f, e = C.c_int32(), C.c_int32()
        check(_lib.lib().ptts_stream_frames(self._h, int(slot), C.byref(f), C.byref(e)))
        return f.value, e.value

    def debug_read(self, name: str, row: int, cap: int = 1920 * 64) -> np.ndarray:
        out = np.empty(cap, np.float32)
        n = _lib.lib().ptts_debug_read(self._h, name.encode(), row, _ptr(out), cap)
        check(int(n))
        return out[:n].copy()

    def profile(self, on: bool):
        check(_lib.lib().ptts_profile_enable(self._h, int(on)))

    def profile_report(self) -> dict[str, dict]:
        buf = C.create_string_buffer(1 << 16)
        n = _lib.lib().ptts_profile_report(self._h, buf, len(buf))
        check(int(n))
        out = {}
        for line in buf.value.decode().splitlines():
            tag, fn, cnt, ms, by, fl = line.split()
            out[tag] = dict(fn=fn, launches=int(cnt), ms=float(ms), bytes=float(by), flops=float(fl))
        return out

    def gemm_replay(self, rows: int, iters: int = 20) -> dict[str, dict]:
        """ptts_profile_gemm_replay: {kind: {us, bytes}} for the four FlowLM decode GEMMs, graph-replayed."""
        us = np.zeros(4, np.float32)
        by = np.zeros(4, np.float64)
        check(_lib.lib().ptts_profile_gemm_replay(self._h, rows, iters, _ptr(us), _ptr(by)))
        return {k: {"us": float(u), "bytes": float(b)} for k, u, b in zip(("in_proj", "out_proj", "linear1", "linear2"), us, by)}

    def profile_overhead_us(self) -> float:
        v = C.c_float()
        check(_lib.lib().ptts_profile_overhead(self._h, C.byref(v)))
        return 1000.0 * v.value

    @property
    def cuda_stream(self) -> int:
        return int(_lib.lib().ptts_cuda_stream(self._h) or 0)

    def launch_count(self, reset: bool = False) -> int:
        return int(_lib.lib().ptts_launch_count(self._h, int(reset)))


class NativeScheduler:
    """ptts_sched_*: continuous batching of long-form requests inside the library (C++), the host only submits segment
    lists and collects PCM.  A request is [("text", StreamSpec) | ("pause", ms)], like tts_model.BatchScheduler."""

    def __init__(self, engine: Engine, voice: Voice, max_batch: int | None = None):
        self._engine = engine
        h = C.c_void_p()
        check(_lib.lib().ptts_sched_create(engine._h, voice._h, int(max_batch or 0), C.byref(h)))
        self._h = h
        self.n_requests = 0

    def submit(self, request: list[tuple]) -> int:
        segs = (Segment * len(request))()
        keep = []
        for i, (kind, val) in enumerate(request):
            if kind == "pause":
                segs[i].kind, segs[i].pause_ms = 1, int(val)
                continue
            tok = np.ascontiguousarray(val.tokens, dtype=np.int32)
            keep.append(tok)
            segs[i].kind, segs[i].n_tokens, segs[i].tokens = 0, len(tok), tok.ctypes.data
            p = segs[i].params
            p.max_gen_len, p.frames_after_eos, p.eos_threshold, p.temp, p.seed = val.max_gen_len, val.frames_after_eos, val.eos_threshold, val.temp, val.seed
            if val.noise is not None:
                nz = np.ascontiguousarray(val.noise, dtype=np.float32)
                assert nz.shape == (val.max_gen_len, LDIM), nz.shape
                keep.append(nz)
                p.noise = nz.ctypes.data
        r = int(_lib.lib().ptts_sched_submit(self._h, segs, len(request)))
        check(r)
        self.n_requests = r + 1
        return r

    def run(self, requests: list[list[tuple]] | None = None, i16: bool = False, view: bool = False) -> list[np.ndarray]:
        """-> one PCM array per request.  view=True: arrays are read-only windows onto the scheduler's own host buffers
        (ptts_sched_result_view, no copy; valid until the next run() / close()), else private copies."""
        import time
        t0 = time.perf_counter()
        for r in requests or []:
            self.submit(r)
        t1 = time.perf_counter()
        check(_lib.lib().ptts_sched_run(self._h, int(i16)))
        t2 = time.perf_counter()
        out = []
        dt = np.int16 if i16 else np.float32
        for r in range(self.n_requests):
            if view:
                ptr, n = C.c_void_p(), C.c_int64()
                check(_lib.lib().ptts_sched_result_view(self._h, r, C.byref(ptr), C.byref(n)))
                if n.value > 0:
                    a = np.frombuffer((C.c_char * (n.value * np.dtype(dt).itemsize)).from_address(ptr.value), dtype=dt)
                    a.flags.writeable = False
                else:
                    a = np.empty(0, dt)
                out.append(a)
                continue
            n = int(_lib.lib().ptts_sched_result_samples(self._h, r))
            a = np.empty(max(n, 0), dt)
            if n > 0:
                check(_lib.lib().ptts_sched_result(self._h, r, _ptr(a), n))
            out.append(a)
        self.last_times = {"submit_s": t1 - t0, "run_s": t2 - t1, "collect_s": time.perf_counter() - t2}
        return out

    @property
    def steps(self) -> int:
        return int(_lib.lib().ptts_sched_steps(self._h))

    def close(self):
        if self._h:
            _lib.lib().ptts_sched_destroy(self._h)
            self._h = None


def config_check(yaml_path) -> None:
    """ptts_config_check: the model YAML against the dimensions the library is compiled for (raises PttsError)."""
    check(_lib.lib().ptts_config_check(str(yaml_path).encode()))


def device_noise(seed: int, frames: int, device: int = 0) -> np.ndarray:
    """The N(0, 1) draws of the device generator for one stream (ptts_test_noise)."""
    out = np.empty((frames, LDIM), np.float32)
    check(_lib.lib().ptts_test_noise(device, seed, frames, _ptr(out)))
    return out


def test_gemm(a, w, bias=None, mode=0, split_k=1, act=0, use_simt=0, device=0):
    a = np.ascontiguousarray(a, np.float32); w = np.ascontiguousarray(w, np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    d = np.zeros((a.shape[0], w.shape[0]), np.float32)
    check(_lib.lib().ptts_test_gemm(device, _ptr(a), _ptr(w), _ptr(b), _ptr(d), a.shape[0], w.shape[0], a.shape[1], mode,
                                    split_k, act, use_simt))
    return d


def test_gemm_int8(a, w, split_k=1, storage=1, device=0):
    """-> (D [rows, feats], per-tensor scale): the int8 weight path of the decode GEMM (ptts.h)."""
    a = np.ascontiguousarray(a, np.float32); w = np.ascontiguousarray(w, np.float32)
    d = np.zeros((a.shape[0], w.shape[0]), np.float32)
    scale = C.c_float()
    check(_lib.lib().ptts_test_gemm_int8(device, _ptr(a), _ptr(w), _ptr(d), a.shape[0], w.shape[0], a.shape[1], split_k,
                                         storage, C.byref(scale)))
    return d, float(scale.value)


def gemv_launches() -> int:
    """Launches of the small-batch GEMV (csrc/gemv.cuh) by this process so far."""
    return int(_lib.lib().ptts_test_gemv_launches())


def test_conv1d(x, prev, w, bias, device=0):
    """x [n,t,cin], prev [n,k-1,cin] or None, w [cout,cin,k] -> y [n,t,cout]"""
    x = np.ascontiguousarray(x, np.float32); w = np.ascontiguousarray(w, np.float32)
    bias = np.ascontiguousarray(bias, np.float32)
    p = None if prev is None else np.ascontiguousarray(prev, np.float32)
    n, t, cin = x.shape
    cout, _, k = w.shape
    y = np.zeros((n, t, cout), np.float32)
    check(_lib.lib().ptts_test_conv1d(device, _ptr(x), _ptr(p), _ptr(w), _ptr(bias), _ptr(y), n, t, cin, cout, k))
    return y


def test_convtr1d(x, prev_row, w, bias, stride, device=0):
    """x [n,t,cin], prev_row [n,cin] or None, w [cin,cout,2*stride] -> y [n,t*stride,cout]"""
    x = np.ascontiguousarray(x, np.float32); w = np.ascontiguousarray(w, np.float32)
    bias = np.ascontiguousarray(bias, np.float32)
    p = None if prev_row is None else np.ascontiguousarray(prev_row, np.float32)
    n, t, cin = x.shape
    cout = w.shape[1]
    y = np.zeros((n, t * stride, cout), np.float32)
    check(_lib.lib().ptts_test_convtr1d(device, _ptr(x), _ptr(p), _ptr(w), _ptr(bias), _ptr(y), n, t, cin, cout, stride))
    return y
