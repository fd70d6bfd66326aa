"""Host-side mirror of the reference's `TTSModel` surface for the generation path
(crates/pocket-tts/src/tts_model.rs), on top of the C-ABI engine.

Same names, argument meaning and error behaviour as the Rust API the `pocket-tts-cuda` crate keeps
(INTEGRATION.md): `load` / `load_with_params`, `get_voice_state_from_prompt_file|tensor`,
`generate`, `generate_stream`, plus the public fields `temp`, `lsd_decode_steps`, `eos_threshold`.
Host text preparation (pause parsing, prompt normalisation, sentence packing, Unigram tokenizer) lives in `text.py`
(SURVEY 8f N2), output formats in `audio.py` (N4).  Pass `tokenizer=` (a `text.UnigramTokenizer`, or any callable
str -> list[int]) or call the `*_tokens` methods with ids.
"""
from __future__ import annotations

import json
import struct
from pathlib import Path
from typing import Callable, Iterator

import numpy as np

from .engine import FRAME, LDIM, Engine, StreamSpec, Voice

SAMPLE_RATE = 24000
# defaults: crates/pocket-tts/src/config.rs:118-124
DEFAULT_TEMPERATURE = 0.7
DEFAULT_LSD_DECODE_STEPS = 1
DEFAULT_EOS_THRESHOLD = -4.0
DEFAULT_VARIANT = "b6369a24"

from .text import (UnigramTokenizer, estimate_frames_after_eos, estimate_generation_steps, long_form_segments,  # noqa: F401
                   max_gen_len, parse_text_with_pauses, prepare_text_prompt, silence_samples, split_into_best_sentences,
                   strip_pause_markers)


def parse_pauses(text: str) -> list[tuple[str, int]]:
    """Explicit `[pause:Xms|Xs]` markers only: [(text_segment, pause_ms_after)], pause 0 for the last segment.
    (generate_stream_long itself also inserts natural pauses: text.long_form_segments.)"""
    from .text import _EXPLICIT, _duration_ms
    out, last = [], 0
    for m in _EXPLICIT.finditer(text):
        out.append((text[last:m.start()], _duration_ms(m.group(1), m.group(2))))
        last = m.end()
    out.append((text[last:], 0))
    return out


def read_safetensors(path: str | Path) -> dict[str, np.ndarray]:
    """Minimal safetensors reader (F32 / BF16 / F16), enough for the checkpoint and `audio_prompt` voice files
    (tts_model.rs:467-487)."""
    data = Path(path).read_bytes()
    (hlen,) = struct.unpack("<Q", data[:8])
    header = json.loads(data[8:8 + hlen])
    base = 8 + hlen
    out = {}
    for name, meta in header.items():
        if name == "__metadata__":
            continue
        b, e = meta["data_offsets"]
        raw = np.frombuffer(data, dtype=np.uint8, count=e - b, offset=base + b)
        dt = meta["dtype"]
        if dt == "F32":
            arr = raw.view(np.float32)
        elif dt == "F16":
            arr = raw.view(np.float16).astype(np.float32)
        elif dt == "BF16":
            arr = (raw.view(np.uint16).astype(np.uint32) << 16).view(np.float32)
        else:
            raise ValueError(f"{name}: unsupported safetensors dtype {dt}")
        out[name] = arr.reshape(meta["shape"]).copy()
    return out


def sample_clamped_noise(frames: int, temp: float, limit: float, seed: int = 0, ldim: int = LDIM) -> np.ndarray:
    """models/flow_lm.rs:39-65 with `noise_clamp = Some(limit)`: i.i.d. N(0, temp) (std = sqrt(temp)) by rejection
    sampling, every value kept only if |v| <= limit.  The reference draws from an unseedable thread RNG; here the draw
    is seeded and handed to the engine as the stream's injected noise [frames, ldim]."""
    if temp <= 0.0:
        return np.zeros((frames, ldim), np.float32)
    if limit <= 0.0:
        raise ValueError("noise_clamp must be positive")
    rng = np.random.default_rng(seed)
    std = np.float32(np.sqrt(np.float32(temp)))
    out = np.empty(frames * ldim, np.float32)
    have = 0
    while have < out.size:
        v = (rng.standard_normal(max(out.size - have, 64), dtype=np.float32) * std).astype(np.float32)
        v = v[np.abs(v) <= limit][: out.size - have]
        out[have:have + v.size] = v
        have += v.size
    return out.reshape(frames, ldim)


class TTSModel:
    def __init__(self, weights: dict[str, np.ndarray], temp: float = DEFAULT_TEMPERATURE,
                 lsd_decode_steps: int = DEFAULT_LSD_DECODE_STEPS, eos_threshold: float = DEFAULT_EOS_THRESHOLD,
                 device: int = 0, max_slots: int = 64, kv_capacity: int = 1024,
                 tokenizer: Callable[[str], list[int]] | None = None):
        self.temp, self.lsd_decode_steps, self.eos_threshold = temp, lsd_decode_steps, eos_threshold
        self.noise_clamp = None  # reference field (tts_model.rs:34): host-side rejection sampling, injected as the stream's noise
        self.sample_rate, self.dim, self.ldim = SAMPLE_RATE, 1024, LDIM
        self.tokenizer = tokenizer
        self.engine = Engine(weights, device=device, max_slots=max_slots, kv_capacity=kv_capacity)
        self._lsd_on_device = 1
        # steps in flight of every live generate_stream iterator: ticket -> {"owner", "flags", "pcm"}.  The engine's ticket
        # ring is shared and its flags come back in step order, while the reference's iterators are independent
        # (tts_model.rs:894-913: each owns a clone of the state): an iterator that needs the engine while another one has
        # a step outstanding fetches that step's results on the other's behalf and parks them here.
        self._tickets: dict[int, dict] = {}

    # ---- loading (tts_model.rs:59-86)
    @classmethod
    def load(cls, weights_path: str | Path, **kw) -> "TTSModel":
        return cls.load_with_params(weights_path, DEFAULT_TEMPERATURE, DEFAULT_LSD_DECODE_STEPS, DEFAULT_EOS_THRESHOLD, **kw)

    @classmethod
    def load_with_params(cls, weights_path: str | Path, temp: float, lsd_decode_steps: int, eos_threshold: float,
                         config: str | Path | None = None, **kw) -> "TTSModel":
        """tts_model.rs:69-86.  `config` = the variant's model YAML (config/{variant}.yaml, config.rs:111): its dimensions are
        checked against the ones the library is compiled for before any weight is read (ptts_config_check); when it is
        None, a `<variant>.yaml` next to the weights file is used if present."""
        from .engine import config_check
        if config is None:
            cand = Path(weights_path).with_suffix(".yaml")
            config = cand if cand.exists() else None
        if config is not None:
            config_check(config)
        return cls(read_safetensors(weights_path), temp, lsd_decode_steps, eos_threshold, **kw)

    # ---- voice state (tts_model.rs:467-501)
    def get_voice_state_from_prompt_file(self, path: str | Path) -> Voice:
        t = read_safetensors(path)
        if "audio_prompt" not in t:
            raise KeyError("'audio_prompt' not found in safetensors file")
        return self.get_voice_state_from_prompt_tensor(t["audio_prompt"])

    def get_voice_state_from_prompt_tensor(self, prompt: np.ndarray) -> Voice:
        return self.engine.voice_from_prompt(np.asarray(prompt, np.float32).reshape(-1, self.dim))

    def get_voice_state_from_tensor(self, audio: np.ndarray) -> Voice:
        """tts_model.rs:504-556: audio f32 [1, 1, T] (or [T]) at the model's sample rate -> voice (Mimi encoder on the GPU)."""
        return self.engine.voice_from_pcm(np.asarray(audio, np.float32).reshape(-1))

    def get_voice_state(self, path: str | Path) -> Voice:
        """tts_model.rs:449-466 for a mono 16-bit PCM WAV already at 24 kHz (the WAV reader variants and the resampler of
        audio.rs are product-shell code and not built; Mimi takes one channel, mimi.channels = 1)."""
        import wave
        with wave.open(str(path), "rb") as w:
            if w.getsampwidth() != 2 or w.getnchannels() != 1:
                raise ValueError("only mono 16-bit PCM WAV is supported")
            if w.getframerate() != self.sample_rate:
                raise ValueError(f"WAV sample rate {w.getframerate()} != {self.sample_rate}: resample first")
            pcm = np.frombuffer(w.readframes(w.getnframes()), "<i2").astype(np.float32) / 32768.0
        return self.get_voice_state_from_tensor(pcm)

    # ---- generation (tts_model.rs:687-703, 894-1071)
    def _sync_params(self):
        if self._lsd_on_device != self.lsd_decode_steps:
            self.engine.set_lsd_steps(self.lsd_decode_steps)
            self._lsd_on_device = self.lsd_decode_steps

    def _tokens(self, prepared: str) -> np.ndarray:
        if self.tokenizer is None:
            raise RuntimeError("no tokenizer attached: pass tokenizer= or use generate_stream_tokens")
        return np.asarray(self.tokenizer(prepared), np.int32)

    # ---- steps in flight, shared by every live iterator of this model
    def _park(self, t: int):
        st = self._tickets[t]
        if st["flags"] is None:
            st["flags"] = self.engine.step_flags(t)
        if st["pcm"] is None:
            st["pcm"] = self.engine.step_pcm(t)

    def _begin(self, owner, ids, ahead=False) -> int:
        for t in sorted(self._tickets):          # another iterator's step: finish it for them first
            if self._tickets[t]["owner"] is not owner:
                self._park(t)
        t = self.engine.step_begin(ids, ahead=ahead)
        self._tickets[t] = {"owner": owner, "flags": None, "pcm": None}
        return t

    def _flags(self, t: int):
        for u in sorted(self._tickets):          # flags come back in step order
            if u < t and self._tickets[u]["flags"] is None:
                self._park(u)
        st = self._tickets[t]
        if st["flags"] is None:
            st["flags"] = self.engine.step_flags(t)
        return st["flags"]

    def _pcm(self, t: int, want=True):
        st = self._tickets.pop(t)
        if st["flags"] is None:
            self.engine.step_flags(t)
        return st["pcm"] if st["pcm"] is not None else self.engine.step_pcm(t, want)

    def generate_stream_tokens(self, tokens, voice: Voice, max_gen_len: int, frames_after_eos: int, noise=None,
                               seed: int = 0) -> Iterator[np.ndarray]:
        """One segment (generate_stream_segment, tts_model.rs:935-1071): yields f32 [1,1,1920] per frame.  Several
        iterators of one model may be consumed in any interleaving, like the reference's (each owns its state)."""
        self._sync_params()
        if noise is None and self.noise_clamp is not None:
            noise = sample_clamped_noise(max_gen_len, self.temp, float(self.noise_clamp), seed, self.ldim)
        spec = StreamSpec(np.asarray(tokens, np.int32), max_gen_len, frames_after_eos, self.eos_threshold, self.temp, seed, noise)
        (slot,) = self.engine.open_streams([voice], [spec])
        ids = np.array([slot], np.int32)
        eng = self.engine
        owner = object()
        mine: list[int] = []

        def begin(ahead=False):
            t = self._begin(owner, ids, ahead=ahead)
            mine.append(t)
            return t

        try:
            # Frame n's codec half overlaps frame n+1's language-model half, and frame n+1 is enqueued before frame n's
            # flags reach the host (PTTS_STEP_AHEAD) unless frame n is known to be the last (max_gen_len); if frame n
            # turns out to end the stream at EOS, the frame enqueued ahead is retired unseen.
            can_ahead = True
            ticket = begin()
            issued = 1
            while True:
                nxt = None
                # (another iterator may already have fetched this step's flags on our behalf: if they say the stream ends
                # here, there is nothing to enqueue ahead)
                parked = self._tickets[ticket]["flags"]
                ends_here = parked is not None and bool(parked[0][0])
                if issued < max_gen_len and can_ahead and not ends_here:
                    try:
                        nxt = begin(ahead=True)
                        issued += 1
                    except Exception as e:  # no spare KV row: fall back to begin-after-flags
                        if getattr(e, "code", 0) != -3:
                            raise
                        can_ahead = False
                fin, _, _ = self._flags(ticket)
                if nxt is None and not fin[0] and issued < max_gen_len:
                    nxt = begin()
                    issued += 1
                frame = self._pcm(ticket).reshape(1, 1, FRAME)
                mine.remove(ticket)
                yield frame
                if fin[0] or nxt is None:
                    break
                ticket = nxt
        finally:
            # retire whatever is still in flight: the frame enqueued ahead of an EOS finish, or everything when the
            # consumer drops the iterator early (the reference's iterator simply stops being polled)
            for t in sorted(mine):
                if t in self._tickets:
                    self._flags(t)
                    self._pcm(t, want=False)
            mine.clear()
            eng.close_stream(int(slot))

    def split_into_best_sentences(self, text: str) -> list[str]:
        """tts_model.rs:603-684: chunks of at most 50 tokens on sentence boundaries."""
        return split_into_best_sentences(text, lambda s: len(self._tokens(s)))

    def generate_stream(self, text: str, voice: Voice, seed: int = 0) -> Iterator[np.ndarray]:
        """tts_model.rs:894-913: every chunk restarts from the (immutable) voice state, frames come out in order."""
        for chunk in self.split_into_best_sentences(text):
            # generate_stream_segment re-prepares its chunk (tts_model.rs:941-969)
            prepared = prepare_text_prompt(chunk)
            yield from self.generate_stream_tokens(self._tokens(prepared), voice, max_gen_len(prepared),
                                                   estimate_frames_after_eos(chunk), seed=seed)

    def generate(self, text: str, voice: Voice, seed: int = 0) -> np.ndarray:
        chunks = list(self.generate_stream(text, voice, seed))
        if not chunks:
            raise RuntimeError("No audio generated")  # tts_model.rs:695-697
        return np.concatenate(chunks, axis=2)[0]

    def generate_stream_long(self, text: str, voice: Voice, seed: int = 0) -> Iterator[np.ndarray]:
        """tts_model.rs:1074-1127: text segments interleaved with host zeros for explicit `[pause:..]` markers and
        natural pauses (ellipsis 500 ms, comma 200 ms)."""
        for kind, val in long_form_segments(text):
            if kind == "text":
                yield from self.generate_stream(val, voice, seed)
            else:
                yield np.zeros((1, 1, silence_samples(val, self.sample_rate)), np.float32)

    def long_form_request(self, text: str, seed: int = 0, noise_fn=None) -> list[tuple]:
        """The same segmentation as a BatchScheduler request: [("text", StreamSpec) | ("pause", ms)]."""
        req: list[tuple] = []
        for kind, val in long_form_segments(text):
            if kind == "pause":
                req.append(("pause", val))
                continue
            for chunk in self.split_into_best_sentences(val):
                prepared = prepare_text_prompt(chunk)
                req.append(("text", StreamSpec(self._tokens(prepared), max_gen_len(prepared), estimate_frames_after_eos(chunk),
                                               self.eos_threshold, self.temp, seed,
                                               noise_fn(max_gen_len(prepared)) if noise_fn else None)))
        return req

    def close(self):
        self.engine.close()


def _run_ahead(sched: "BatchScheduler", requests: list[list[tuple]]) -> list[np.ndarray]:
    """BatchScheduler.run with the device kept one step ahead of the host (PTTS_STEP_AHEAD): while the host unpacks step
    n (flags, PCM rows, bookkeeping) the device already runs step n+1 on the same rows.  A step is only enqueued ahead
    when no row can reach its max_gen_len on the current step (the host knows that); a row that ends at EOS instead
    shows up in the step enqueued ahead as an overrun row (dropped) and its slot is closed once that step is drained."""
    eng = sched.engine
    out: list[list[np.ndarray]] = [[] for _ in requests]
    cursor = [0] * len(requests)
    waiting = list(range(len(requests)))
    active: dict[int, int] = {}      # slot -> request
    budget: dict[int, int] = {}      # slot -> frames the stream may still produce (max_gen_len - frames begun)

    def admit():
        specs, owners, still = [], [], []
        for r in waiting:
            while cursor[r] < len(requests[r]) and requests[r][cursor[r]][0] == "pause":
                out[r].append(np.zeros(silence_samples(requests[r][cursor[r]][1]), np.float32))
                cursor[r] += 1
            if cursor[r] >= len(requests[r]):
                continue
            if len(active) + len(specs) < sched.max_batch:
                specs.append(requests[r][cursor[r]][1])
                owners.append(r)
                cursor[r] += 1
            else:
                still.append(r)
        waiting[:] = still
        if specs:
            for slot, r, sp in zip(eng.open_streams([sched.voice] * len(specs), specs), owners, specs):
                active[int(slot)] = r
                budget[int(slot)] = int(sp.max_gen_len)

    def begin(slots, ahead):
        t = eng.step_begin(slots, ahead=ahead)
        for s in slots:
            budget[int(s)] -= 1
        return t

    inflight: list[tuple] = []        # [(ticket, slots, owners)] in step order, at most two
    try:
        admit()
        while active or inflight:
            if not inflight:
                slots = np.fromiter(active.keys(), np.int32, len(active))
                inflight.append((begin(slots, False), slots, [active[int(s)] for s in slots]))
            ticket, slots, owners = inflight[0]
            # the same rows again, ahead of this step's flags, unless one of them is on its last possible frame
            if len(inflight) == 1 and all(budget[int(s)] > 0 for s in slots):
                try:
                    inflight.append((begin(slots, True), slots, owners))
                except Exception as e:
                    if getattr(e, "code", 0) != -3:   # no spare KV row: stay in lock step
                        raise
            fin, _, _ = eng.step_flags(ticket)
            over = eng.last_overrun.copy() if hasattr(eng, "last_overrun") else np.zeros(len(slots), bool)
            pcm = eng.step_pcm(ticket)
            inflight.pop(0)
            for row, r in enumerate(owners):
                if not over[row]:
                    out[r].append(pcm[row])
            done = [int(s) for s, f, o in zip(slots, fin, over) if f and not o]
            if done:
                if inflight:   # the step enqueued ahead still reads the finished slots: drain it first
                    t2, s2, o2 = inflight.pop(0)
                    fin2, _, _ = eng.step_flags(t2)
                    over2 = eng.last_overrun.copy() if hasattr(eng, "last_overrun") else np.zeros(len(s2), bool)
                    pcm2 = eng.step_pcm(t2)
                    for row, r in enumerate(o2):
                        if not over2[row]:
                            out[r].append(pcm2[row])
                    done += [int(s) for s, f, o in zip(s2, fin2, over2) if f and not o and int(s) not in done]
                for s in done:
                    eng.close_stream(s)
                    waiting.append(active.pop(s))
                    budget.pop(s, None)
                admit()
    finally:
        for t, _, _ in inflight:
            try:
                eng.step_flags(t)
            except Exception:
                pass
            try:
                eng.step_pcm(t, want=False)
            except Exception:
                pass
        for s in list(active):
            try:
                eng.close_stream(s)
            except Exception:
                pass
            active.pop(s, None)
    return [np.concatenate(o) if o else np.zeros(0, np.float32) for o in out]


def shard_requests(n_requests: int, world_size: int, rank: int) -> range:
    """Request sharding across the GPUs of one box: contiguous, disjoint, covering; streams are
    independent so no data-path collective exists."""
    base, rem = divmod(n_requests, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


class BatchScheduler:
    """Continuous batching of many long-form requests on one engine (BASELINE configs[4]: concurrent 60 s requests
    with `[pause:Xms]`).  A request is an ordered list of segments, ("text", StreamSpec) or ("pause", ms); the chunks
    of one request run one after the other (ordered emission, each restarting from the voice state exactly like
    generate_stream's flat_map, tts_model.rs:899-910) while different requests fill the batch.  Whenever streams
    finish, the next chunks of their requests are opened together (one batched text prefill) and join the next step;
    pauses are host zeros (tts_model.rs:1115-1124).  Frames are fetched one step behind the launch
    (ptts_step_begin / flags / pcm), so the codec of frame n overlaps the language model of frame n+1."""

    def __init__(self, engine: Engine, voice: Voice, max_batch: int | None = None):
        self.engine, self.voice = engine, voice
        self.max_batch = max_batch or engine.max_batch

    def run(self, requests: list[list[tuple]], ahead: bool = False, native: bool = False, i16: bool = False) -> list[np.ndarray]:
        """ahead=True keeps the device one step ahead of the host (see _run_ahead); the default is the lock-step loop.
        native=True runs the same policy inside the library (ptts_sched_*, C++): the host only submits the segment lists
        and collects each request's PCM (f32, or i16 packed on the device with i16=True)."""
        if native:
            from .engine import NativeScheduler
            ns = NativeScheduler(self.engine, self.voice, self.max_batch)
            try:
                return ns.run(requests, i16=i16)
            finally:
                ns.close()
        if ahead:
            return _run_ahead(self, requests)
        eng = self.engine
        out: list[list[np.ndarray]] = [[] for _ in requests]
        cursor = [0] * len(requests)          # next segment of each request
        waiting = list(range(len(requests)))  # requests with no stream in flight
        active: dict[int, int] = {}           # slot -> request
        pending = None                        # (ticket, [requests in row order]) whose PCM is still on the device

        def admit():
            specs, owners = [], []
            still = []
            for r in waiting:
                # host-side pauses are consumed immediately; the first text segment claims a slot
                while cursor[r] < len(requests[r]) and requests[r][cursor[r]][0] == "pause":
                    out[r].append(np.zeros(silence_samples(requests[r][cursor[r]][1]), np.float32))
                    cursor[r] += 1
                if cursor[r] >= len(requests[r]):
                    continue
                if len(active) + len(specs) < self.max_batch:
                    specs.append(requests[r][cursor[r]][1])
                    owners.append(r)
                    cursor[r] += 1
                else:
                    still.append(r)
            waiting[:] = still
            if specs:
                for slot, r in zip(eng.open_streams([self.voice] * len(specs), specs), owners):
                    active[int(slot)] = r

        def collect(p):
            ticket, owners = p
            pcm = eng.step_pcm(ticket)
            for row, r in enumerate(owners):
                out[r].append(pcm[row])

        try:
            admit()
            while active:
                slots = np.fromiter(active.keys(), np.int32, len(active))
                owners = [active[int(s)] for s in slots]
                ticket = eng.step_begin(slots)
                fin, _, _ = eng.step_flags(ticket)
                if pending is not None:
                    collect(pending)
                pending = (ticket, owners)
                done = [int(s) for s, f in zip(slots, fin) if f]
                if done:
                    collect(pending)  # a finished slot is closed below: drain the step that still reads it
                    pending = None
                    for s in done:
                        eng.close_stream(s)
                        waiting.append(active.pop(s))
                    admit()
            if pending is not None:
                collect(pending)
                pending = None
        finally:
            # an error part-way (a bad spec, capacity) must not leave steps in flight or slots open on the shared engine
            if pending is not None:
                try:
                    eng.step_pcm(pending[0], want=False)
                except Exception:
                    pass
            for s in list(active):
                try:
                    eng.close_stream(s)
                except Exception:
                    pass
                active.pop(s, None)
        return [np.concatenate(o) if o else np.zeros(0, np.float32) for o in out]
