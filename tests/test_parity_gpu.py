"""Parity of the CUDA engine (through the C ABI) against the CPU oracle and against the golden
vectors produced by the unmodified reference PyTorch package.

Tolerances are the ones BASELINE.json's north_star states for 16-bit operands:
  per-frame FlowLM latents  max-abs <= 1e-2   (teacher-forced: the reference latent is fed back, so the
                                               comparison is per frame and not a chaotic AR divergence)
  Mimi PCM                  SNR >= 40 dB
  EOS frame index / frame counts bit-exact
"""
import numpy as np
import pytest
import torch

from pocket_tts_b200 import synth

pytestmark = pytest.mark.gpu

LAT_TOL = 1e-2
SNR_MIN = 40.0


def snr(ref, x):
    return 10 * np.log10((ref ** 2).sum() / max(((ref - x) ** 2).sum(), 1e-30))


_cache = {}


def engine_for(seed, ls, **kw):
    from pocket_tts_b200.engine import Engine
    key = (int(seed), float(ls), tuple(sorted(kw.items())))
    if key not in _cache:
        for e in _cache.values():
            e[0].close()
        _cache.clear()
        w = synth.make_weights(key[0], layer_scale=key[1])
        _cache[key] = (Engine(w, max_slots=kw.get("max_slots", 8), kv_capacity=kw.get("kv_capacity", 512)), w)
    return _cache[key]


def run_engine(eng, voice, g, kind="tanh", teacher=True, frames=None):
    from pocket_tts_b200.engine import StreamSpec
    frames = frames or g[f"{kind}_latents"].shape[0]
    slots = eng.open_streams([voice], [StreamSpec(g["tokens"], frames, 0, 1e30, noise=g["noise"][:frames])])
    lat, pcm, logit = [], [], []
    for f in range(frames):
        if teacher and f > 0:
            eng.set_feedback(int(slots[0]), g[f"{kind}_latents"][f - 1])
        p, fin, l, lg = eng.step(slots)
        lat.append(l[0]); pcm.append(p[0]); logit.append(lg[0])
        assert bool(fin[0]) == (f == frames - 1)
    eng.close_stream(int(slots[0]))
    return np.stack(lat), np.stack(pcm), np.array(logit)


@pytest.mark.parametrize("case", ["cfg1_lsd1", "cfg3_lsd4", "stress_ls05"])
def test_golden_teacher_forced(golden_dir, case):
    g = np.load(golden_dir / f"{case}.npz")
    eng, _ = engine_for(g["weight_seed"], g["layer_scale"])
    eng.set_lsd_steps(int(g["lsd_steps"]))
    voice = eng.voice_from_prompt(synth.make_voice_prompt(int(g["voice_rows"]), seed=7))
    lat, pcm, logit = run_engine(eng, voice, g)
    voice.close()
    eng.set_lsd_steps(1)
    err = np.abs(lat - g["tanh_latents"]).max()
    assert err <= LAT_TOL, f"latent max-abs {err}"
    # first frame uses no fed-back latent at all; later PCM frames depend on the engine's own Mimi state
    s = snr(g["tanh_pcm"], pcm)
    assert s >= SNR_MIN, f"PCM SNR {s} dB"
    assert np.abs(logit - g["tanh_eos_logits"]).max() < 2e-2


def test_golden_free_running_prefix(golden_dir):
    """Free-running AR (engine feeds its own latents back) stays inside the tolerance over the golden window."""
    g = np.load(golden_dir / "cfg1_lsd1.npz")
    eng, _ = engine_for(g["weight_seed"], g["layer_scale"])
    voice = eng.voice_from_prompt(synth.make_voice_prompt(int(g["voice_rows"]), seed=7))
    lat, pcm, _ = run_engine(eng, voice, g, teacher=False, frames=6)
    voice.close()
    assert np.abs(lat - g["tanh_latents"][:6]).max() <= LAT_TOL
    assert snr(g["tanh_pcm"][:6], pcm) >= SNR_MIN


def test_eos_frame_index_bit_exact(golden_dir):
    """EOS bookkeeping (tts_model.rs:1055-1069): with a threshold placed in the widest gap of the reference's
    own logit trace, the engine must stop on exactly the reference's frame."""
    from pocket_tts_b200.engine import StreamSpec
    g = np.load(golden_dir / "stress_ls05.npz")
    eng, _ = engine_for(g["weight_seed"], g["layer_scale"])
    voice = eng.voice_from_prompt(synth.make_voice_prompt(int(g["voice_rows"]), seed=7))
    logits = g["tanh_eos_logits"]
    srt = np.sort(logits)
    gaps = np.diff(srt)
    # thresholds in the three widest gaps -> different eos steps; margin is half the gap
    for gi in np.argsort(gaps)[-3:]:
        thr = float((srt[gi] + srt[gi + 1]) / 2)
        margin = float(gaps[gi] / 2)
        assert margin > 0.02, margin
        for fae in (3, 5):
            above = np.nonzero(logits > thr)[0]
            eos_step = int(above[0]) if len(above) else -1
            want = min(len(logits), eos_step + fae + 1) if eos_step >= 0 else len(logits)
            slots = eng.open_streams([voice], [StreamSpec(g["tokens"], len(logits), fae, thr, noise=g["noise"])])
            n = 0
            while True:
                if n > 0:
                    eng.set_feedback(int(slots[0]), g["tanh_latents"][n - 1])
                _, fin, _, _ = eng.step(slots)
                n += 1
                if fin[0]:
                    break
            frames, es = eng.stream_frames(int(slots[0]))
            eng.close_stream(int(slots[0]))
            assert (n, frames, es) == (want, want, eos_step), (thr, fae, n, frames, es, want, eos_step)
    voice.close()


def test_ragged_batch_matches_oracle():
    """B independent streams = B independent batch-1 reference runs: different token counts,
    noise and lifetimes in one batch, slots recycled, each stream compared with its own oracle run."""
    from oracle import ptts_oracle as O
    from pocket_tts_b200.engine import StreamSpec
    eng, wnp = engine_for(1234, 0.01)
    W = O.to_torch(wnp)
    prompt = synth.make_voice_prompt(33, seed=21)
    voice = eng.voice_from_prompt(prompt)
    ov = O.voice_state_from_prompt(W, prompt)
    specs, refs = [], []
    for i, (ntok, frames) in enumerate([(5, 3), (17, 5), (1, 2), (9, 4), (30, 3)]):
        tok = synth.make_tokens(ntok, seed=100 + i)
        noise = synth.make_noise(frames, seed=200 + i)
        specs.append(StreamSpec(tok, frames, 0, 1e30, noise=noise))
        refs.append(O.generate_segment(W, ov, tok, noise, frames, 0, float("inf")))
    slots = eng.open_streams([voice] * len(specs), specs)
    active = list(range(len(specs)))
    got_lat = [[] for _ in specs]
    got_pcm = [[] for _ in specs]
    step = 0
    while active:
        for i in active:  # teacher forcing per stream
            if step > 0:
                eng.set_feedback(int(slots[i]), refs[i]["latents"][step - 1])
        pcm, fin, lat, _ = eng.step(slots[active])
        nxt = []
        for j, i in enumerate(active):
            got_lat[i].append(lat[j]); got_pcm[i].append(pcm[j])
            if fin[j]:
                eng.close_stream(int(slots[i]))
            else:
                nxt.append(i)
        active = nxt
        step += 1
    for i, r in enumerate(refs):
        assert len(got_lat[i]) == r["frames"]
        assert np.abs(np.stack(got_lat[i]) - r["latents"]).max() <= LAT_TOL
        assert snr(r["pcm"], np.stack(got_pcm[i])) >= SNR_MIN
    # recycled slot must start from clean streaming state
    tok = synth.make_tokens(5, seed=100)
    noise = synth.make_noise(3, seed=200)
    s2 = eng.open_streams([voice], [StreamSpec(tok, 3, 0, 1e30, noise=noise)])
    out = []
    for f in range(3):
        if f:
            eng.set_feedback(int(s2[0]), refs[0]["latents"][f - 1])
        pcm, _, _, _ = eng.step(s2)
        out.append(pcm[0])
    eng.close_stream(int(s2[0]))
    assert snr(refs[0]["pcm"], np.stack(out)) >= SNR_MIN
    voice.close()


def test_temp_zero_is_deterministic_and_device_noise_runs():
    """temp = 0 -> x_0 = 0 exactly (flow_lm.rs:39-48 with std 0), the reference's own determinism test
    (tests/streaming_tests.rs:21-70); temp > 0 without injected noise uses the device generator."""
    from pocket_tts_b200.engine import StreamSpec
    eng, _ = engine_for(1234, 0.01)
    voice = eng.voice_from_prompt(synth.make_voice_prompt(20, seed=3))
    tok = synth.make_tokens(8, seed=1)
    outs = []
    for _ in range(2):
        s = eng.open_streams([voice], [StreamSpec(tok, 3, 0, 1e30, temp=0.0)])
        outs.append(np.stack([eng.step(s)[0][0] for _ in range(3)]))
        eng.close_stream(int(s[0]))
    np.testing.assert_array_equal(outs[0], outs[1])  # ordered split-K: bit-reproducible
    s = eng.open_streams([voice, voice], [StreamSpec(tok, 2, 0, 1e30, temp=0.7, seed=1), StreamSpec(tok, 2, 0, 1e30, temp=0.7, seed=2)])
    pcm, _, lat, _ = eng.step(s)
    assert np.isfinite(pcm).all() and np.abs(lat[0] - lat[1]).max() > 1e-3
    for x in s:
        eng.close_stream(int(x))
    voice.close()


def test_pipelined_step_matches_synchronous_step():
    """begin/flags/pcm with two frames in flight (codec of frame n overlapping the LM of frame n+1) must give
    bit-identical frames to the synchronous ptts_step."""
    from pocket_tts_b200.engine import StreamSpec
    eng, _ = engine_for(1234, 0.01)
    voice = eng.voice_from_prompt(synth.make_voice_prompt(30, seed=3))
    specs = [StreamSpec(synth.make_tokens(7 + i, seed=i), 6, 0, 1e30, temp=0.7, seed=i) for i in range(3)]
    s = eng.open_streams([voice] * 3, specs)
    ref = [eng.step(s) for _ in range(6)]
    for x in s:
        eng.close_stream(int(x))
    s = eng.open_streams([voice] * 3, specs)
    got, prev = [], None
    for f in range(6):
        t = eng.step_begin(s)
        fin, lat, logit = eng.step_flags(t)
        np.testing.assert_array_equal(lat, ref[f][2])
        assert (fin == ref[f][1]).all()
        if prev is not None:
            got.append(eng.step_pcm(prev))
        prev = t
    got.append(eng.step_pcm(prev))
    for f in range(6):
        np.testing.assert_array_equal(got[f], ref[f][0])
    for x in s:
        eng.close_stream(int(x))
    voice.close()


def test_api_errors():
    from pocket_tts_b200._lib import PttsError
    from pocket_tts_b200.engine import StreamSpec
    eng, _ = engine_for(1234, 0.01)
    voice = eng.voice_from_prompt(synth.make_voice_prompt(4, seed=3))
    with pytest.raises(PttsError) as ei:  # KV capacity
        eng.open_streams([voice], [StreamSpec(synth.make_tokens(10, 1), 100000)])
    assert ei.value.code == -3
    with pytest.raises(PttsError):        # bad token id
        eng.open_streams([voice], [StreamSpec(np.array([4001], np.int32), 4)])
    s = eng.open_streams([voice], [StreamSpec(synth.make_tokens(3, 1), 1, temp=0.0)])
    _, fin, _, _ = eng.step(s)
    assert fin[0]
    with pytest.raises(PttsError) as ei:  # stepping a finished stream
        eng.step(s)
    assert ei.value.code == -4
    eng.close_stream(int(s[0]))
    with pytest.raises(PttsError):
        eng.close_stream(int(s[0]))
    voice.close()


def test_host_mirror_generate_stream_tokens():
    """TTSModel facade: yields [1,1,1920] frames until the engine reports the last one (tts_model.rs:894-1071)."""
    from pocket_tts_b200.tts_model import TTSModel
    for e in _cache.values():
        e[0].close()
    _cache.clear()
    m = TTSModel(synth.make_weights(1234), temp=0.0, max_slots=2, kv_capacity=256)
    voice = m.get_voice_state_from_prompt_tensor(synth.make_voice_prompt(10, seed=1)[None])
    m.eos_threshold = 1e30
    frames = list(m.generate_stream_tokens(synth.make_tokens(6, 2), voice, max_gen_len=4, frames_after_eos=3))
    assert len(frames) == 4 and frames[0].shape == (1, 1, 1920)
    m.eos_threshold = -1e30  # EOS at step 0 -> 0 + 3 + 1 frames (D2)
    frames = list(m.generate_stream_tokens(synth.make_tokens(6, 2), voice, max_gen_len=20, frames_after_eos=3))
    assert len(frames) == 4
    voice.close()
    m.close()


def _encoder_engine(g, **kw):
    from pocket_tts_b200.engine import Engine
    for e in _cache.values():
        e[0].close()
    _cache.clear()
    w = dict(synth.make_weights(1234, layer_scale=0.01))
    w.update(synth.make_encoder_weights(int(g["enc_seed"]), layer_scale=float(g["enc_layer_scale"])))
    return Engine(w, max_slots=2, kv_capacity=256, **kw), w


def test_voice_cloning_from_pcm_matches_reference(golden_dir):
    """BASELINE configs[2] / SURVEY 8f N1: PCM -> SEANetEncoder -> encoder transformer -> ConvDownsample1d -> speaker_proj
    on the GPU against the unmodified reference package's `_encode_audio` (tanh-GELU golden) and the oracle.  The
    reference's own tolerance for this tensor is 2e-2 max-abs (crates/pocket-tts/tests/parity_tests.rs:60-142);
    f16 operands stay well inside it.  Then the voice built from PCM must behave like the voice built from the same
    conditioning rows."""
    from oracle import ptts_oracle as O
    from pocket_tts_b200.engine import StreamSpec
    g = np.load(golden_dir / "enc_pcm22.npz")
    eng, w = _encoder_engine(g)
    pcm = synth.make_pcm(int(g["n_samples"]), seed=int(g["pcm_seed"]))
    got = eng.audio_prompt_from_pcm(pcm)
    want = g["tanh_audio_prompt"]
    assert got.shape == want.shape == (22, 1024)
    err = np.abs(got - want).max()
    assert err <= 2e-2, f"audio_prompt max-abs {err}"
    rel = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert rel < 5e-3, rel
    # ragged lengths: one sample, one frame minus one, exactly one frame
    W = O.to_torch(w)
    for n in (1, 1919, 1920, 1921):
        p = synth.make_pcm(n, seed=n)
        a = eng.audio_prompt_from_pcm(p)
        b = O.audio_prompt_from_pcm(W, p, "tanh").numpy()
        assert a.shape == b.shape == ((n + 1919) // 1920, 1024)
        assert np.abs(a - b).max() <= 2e-2
    # voice from PCM == voice from its conditioning rows (same prefill either way)
    tok = synth.make_tokens(7, seed=3)
    noise = synth.make_noise(3, seed=4)
    outs = []
    for voice in (eng.voice_from_pcm(pcm), eng.voice_from_prompt(got)):
        s = eng.open_streams([voice], [StreamSpec(tok, 3, 0, 1e30, noise=noise)])
        frames = [eng.step(s) for _ in range(3)]
        outs.append((np.stack([f[0][0] for f in frames]), np.stack([f[2][0] for f in frames])))
        eng.close_stream(int(s[0]))
        voice.close()
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    # a prompt of more than 120 frames: the reference's chunked encoding (one carried state, the downsample's replicate
    # padding restarted at frame 120, tts_model.rs:528-541)
    p = synth.make_pcm(123 * 1920 - 5, seed=9)
    a = eng.audio_prompt_from_pcm(p)
    b = O.audio_prompt_from_pcm(W, p, "tanh").numpy()
    assert a.shape == b.shape == (123, 1024)
    assert np.abs(a - b).max() <= 2e-2
    unchunked = O.audio_prompt_from_pcm(W, p, "tanh", chunk_frames=123).numpy()
    assert np.abs(b[120] - unchunked[120]).max() > 10 * np.abs(a[120] - b[120]).max()  # the boundary frame really is the restarted one
    from pocket_tts_b200 import _lib
    with pytest.raises(_lib.PttsError) as ei:
        eng.audio_prompt_from_pcm(np.zeros(1025 * 1920, np.float32))
    assert ei.value.code == -3
    eng.close()


def test_cfg3_voice_cloning_batch32_lsd4(golden_dir):
    """BASELINE configs[2] at its own shape: the voice comes from an 87-frame PCM prompt (the length of the reference's
    assets/ref.wav) through the GPU Mimi encoder, lsd_decode_steps = 4, 32 concurrent streams.  Two of the streams are
    teacher-forced against the oracle run end to end from the same PCM (oracle encoder -> oracle prefill -> oracle
    frames); the others only have to finish with finite audio and the right frame counts."""
    from oracle import ptts_oracle as O
    from pocket_tts_b200.engine import Engine, StreamSpec
    g = np.load(golden_dir / "enc_pcm22.npz")
    for e in _cache.values():
        e[0].close()
    _cache.clear()
    w = dict(synth.make_weights(1234, layer_scale=0.01))
    w.update(synth.make_encoder_weights(int(g["enc_seed"]), layer_scale=float(g["enc_layer_scale"])))
    eng = Engine(w, max_slots=32, kv_capacity=128)
    eng.set_lsd_steps(4)
    W = O.to_torch(w)
    pcm = synth.make_pcm(87 * 1920 - 123, seed=17)
    voice = eng.voice_from_pcm(pcm)
    ov = O.voice_state_from_prompt(W, O.audio_prompt_from_pcm(W, pcm, "tanh"))
    n, frames = 32, 3
    checked = (0, 31)
    specs, refs = [], {}
    for i in range(n):
        tok = synth.make_tokens(4 + (i * 7) % 30, seed=300 + i)
        noise = synth.make_noise(frames, seed=400 + i)
        specs.append(StreamSpec(tok, frames, 0, 1e30, noise=noise))
        if i in checked:
            refs[i] = O.generate_segment(W, ov, tok, noise, frames, 0, float("inf"), lsd_steps=4)
    slots = eng.open_streams([voice] * n, specs)
    lat = {i: [] for i in checked}
    pcm_out = {i: [] for i in checked}
    for f in range(frames):
        if f:
            for i in checked:
                eng.set_feedback(int(slots[i]), refs[i]["latents"][f - 1])
        p, fin, l, _ = eng.step(slots)
        assert np.isfinite(p).all() and bool(fin.all()) == (f == frames - 1)
        for i in checked:
            lat[i].append(l[i]); pcm_out[i].append(p[i])
    for i in checked:
        assert np.abs(np.stack(lat[i]) - refs[i]["latents"]).max() <= LAT_TOL
        assert snr(refs[i]["pcm"], np.stack(pcm_out[i])) >= SNR_MIN
    assert all(eng.stream_frames(int(s))[0] == frames for s in slots)
    for s in slots:
        eng.close_stream(int(s))
    voice.close()
    eng.close()


def test_facade_text_to_audio_end_to_end(tmp_path):
    """BASELINE configs[0] through the host mirror of the reference surface: WAV voice prompt -> get_voice_state,
    text -> prepare / sentence split / Unigram tokenizer -> generate / generate_stream_long -> WAV, with the frame
    arithmetic of tts_model.rs:968,1055-1069 and the pause arithmetic of pause.rs:183-185."""
    from pocket_tts_b200 import audio as A
    from pocket_tts_b200 import text as T
    from pocket_tts_b200.tts_model import TTSModel
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("_text_fixtures", Path(__file__).with_name("test_text.py"))
    fixtures = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fixtures)
    _synthetic_vocab = fixtures._synthetic_vocab
    for e in _cache.values():
        e[0].close()
    _cache.clear()
    w = dict(synth.make_weights(1234, layer_scale=0.01))
    w.update(synth.make_encoder_weights(4321))
    tok = T.UnigramTokenizer(_synthetic_vocab(), 0, True, "always", ())
    m = TTSModel(w, temp=0.7, max_slots=2, kv_capacity=256, tokenizer=tok)
    wav = tmp_path / "prompt.wav"
    prompt_pcm = synth.make_pcm(5 * 1920 + 77, seed=2)[None]
    A.write_wav(wav, prompt_pcm, 24000)
    voice = m.get_voice_state(wav)
    assert len(voice) == 6                                    # 5 frames + 77 samples -> padded to 6 frames
    m.eos_threshold = 1e30                                    # never EOS: runs to max_gen_len = (words + 2) * 13
    pcm = m.generate("Hello, world!", voice)
    assert pcm.shape == (1, (2 + 2) * 13 * 1920) and np.isfinite(pcm).all() and np.abs(pcm).max() > 0
    m.eos_threshold = -1e30                                   # EOS at step 0: 0 + frames_after_eos (5 for <= 4 words) + 1
    pcm = m.generate("Hello, world!", voice)
    assert pcm.shape == (1, 6 * 1920)
    chunks = list(m.generate_stream_long("Hello there [pause:250ms] world", voice))
    sizes = [c.shape[-1] for c in chunks]
    assert sizes.count(6000) == 1 and all(c.shape[:2] == (1, 1) for c in chunks)   # 250 ms at 24 kHz
    assert not chunks[sizes.index(6000)].any() and sum(sizes) == 6000 + 2 * 6 * 1920
    out = tmp_path / "out.wav"
    A.write_wav(out, np.concatenate(chunks, axis=2)[0], m.sample_rate)
    import wave
    with wave.open(str(out)) as f:
        assert f.getnframes() == sum(sizes) and f.getframerate() == 24000
    # a consumer that drops the iterator after one frame leaves nothing in flight behind it
    m.eos_threshold = 1e30
    it = m.generate_stream("Hello, world!", voice)
    assert next(it).shape == (1, 1, 1920)
    it.close()
    m.eos_threshold = -1e30
    assert m.generate("Hello, world!", voice).shape == (1, 6 * 1920)
    m.noise_clamp = 1.0                                       # host-side rejection sampling feeds the stream's noise
    assert m.generate("Hello, world!", voice).shape == (1, 6 * 1920)
    voice.close()
    m.close()


def test_voice_cloning_needs_encoder_tensors():
    from pocket_tts_b200 import _lib
    eng, _ = engine_for(1234, 0.01)
    with pytest.raises(_lib.PttsError) as ei:
        eng.voice_from_pcm(np.zeros(1920, np.float32))
    assert ei.value.code == -4


def test_step_ahead_matches_lockstep_and_never_emits_past_the_end():
    """PTTS_STEP_AHEAD: frame n+1 enqueued before frame n's flags are fetched.  Same frames bit for bit as the lock-step
    calls; a stream that ends at EOS (D2: frames = eos_step + frames_after_eos + 1, tts_model.rs:1055-1069) reports
    its one overrun frame as such, its frame counters stay at the last real frame, and a batch keeps going for the
    streams that did not end."""
    from pocket_tts_b200.engine import StreamSpec
    from pocket_tts_b200 import _lib
    eng, _ = engine_for(1234, 0.01)
    voice = eng.voice_from_prompt(synth.make_voice_prompt(12, seed=3))
    toks = [synth.make_tokens(7, seed=21), synth.make_tokens(5, seed=22)]
    noise = [synth.make_noise(8, seed=31), synth.make_noise(8, seed=32)]

    def specs(thr0):
        # stream 0 ends by EOS when thr0 is -inf-like (eos at step 0, 3 frames after -> 4 frames); stream 1 runs 8 frames
        return [StreamSpec(toks[0], 8, 3, thr0, noise=noise[0]), StreamSpec(toks[1], 8, 3, 1e30, noise=noise[1])]

    # lock-step reference run
    slots = eng.open_streams([voice, voice], specs(-1e30))
    want = []
    live = list(slots)
    while live:
        p, fin, l, _ = eng.step(np.array(live, np.int32))
        want.append({int(s): (p[i].copy(), l[i].copy(), bool(fin[i])) for i, s in enumerate(live)})
        live = [s for i, s in enumerate(live) if not fin[i]]
    assert [len(w) for w in want] == [2, 2, 2, 2, 1, 1, 1, 1]
    frames_lock = [eng.stream_frames(int(s)) for s in slots]
    for s in slots:
        eng.close_stream(int(s))

    # the same job one step ahead of the host
    slots = eng.open_streams([voice, voice], specs(-1e30))
    live = [int(s) for s in slots]
    got = []
    ticket, rows = eng.step_begin(np.array(live, np.int32)), list(live)
    issued = 1
    while ticket is not None:
        nxt = nxt_rows = None
        if issued < 8 and live:
            nxt_rows = list(live)           # chosen before this step's flags are known
            nxt = eng.step_begin(np.array(nxt_rows, np.int32), ahead=True)
            issued += 1
        fin, lat, _ = eng.step_flags(ticket)
        over = eng.last_overrun.copy()
        pcm = eng.step_pcm(ticket)
        got.append({s: (pcm[i].copy(), lat[i].copy(), bool(fin[i])) for i, s in enumerate(rows) if not over[i]})
        live = [s for s in live if not any(s == r and fin[i] for i, r in enumerate(rows))]
        ticket, rows = nxt, nxt_rows
    assert len(got) == len(want)
    for step, (g, w) in enumerate(zip(got, want)):
        assert g.keys() == w.keys()
        for s_g, s_w in zip(sorted(g), sorted(w)):
            if step <= 4:
                # same batch geometry as the lock-step run (step 4 carries the overrun row of stream 0 next to stream 1,
                # the lock-step run has stream 1 alone: different split-K tiling, so only steps 0-3 are bit-identical)
                cmp = np.testing.assert_array_equal if step < 4 else (lambda a, b: np.testing.assert_allclose(a, b, atol=5e-3))
            else:
                cmp = lambda a, b: np.testing.assert_allclose(a, b, atol=2e-2)  # free-running from step 4's rounding
            cmp(g[s_g][0], w[s_w][0])
            cmp(g[s_g][1], w[s_w][1])
            assert g[s_g][2] == w[s_w][2]
    assert [eng.stream_frames(int(s)) for s in slots] == frames_lock == [(4, 0), (8, -1)]
    # rules: two steps ahead of unfetched flags is refused; a finished slot cannot be stepped
    with pytest.raises(_lib.PttsError):
        eng.step_begin(np.array([int(slots[0])], np.int32))
    for s in slots:
        eng.close_stream(int(s))
    s2 = eng.open_streams([voice], [StreamSpec(toks[0], 8, 3, 1e30, noise=noise[0])])
    t0 = eng.step_begin(s2)
    t1 = eng.step_begin(s2, ahead=True)
    with pytest.raises(_lib.PttsError):
        eng.step_begin(s2, ahead=True)
    for t in (t0, t1):
        eng.step_flags(t); eng.step_pcm(t)
    eng.close_stream(int(s2[0]))
    voice.close()


def test_int8_weight_mode_matches_reference_quantisation():
    """BASELINE configs[3]: per-tensor symmetric int8 with the reference's scheme and skip list
    (crates/pocket-tts/src/quantize.rs:27-41,65-94,117-154).  The oracle runs f32 math on the fake-quantised
    tensors exactly like the reference's QuantizedTensor; the engine holds the integer codes as the GEMM operand and
    applies the scale in the epilogue."""
    from oracle import ptts_oracle as O
    from pocket_tts_b200.engine import Engine, StreamSpec
    for e in _cache.values():
        e[0].close()
    _cache.clear()
    wnp = synth.make_weights(1234)
    wq = {}
    n_quant = 0
    for name, w in wnp.items():
        if O.should_quantize(name, w.size):
            q, scale = O.quantize_per_tensor(w)
            wq[name] = (q.astype(np.float32) * scale).reshape(w.shape)
            n_quant += 1
        else:
            wq[name] = w
    assert n_quant > 60
    assert not np.array_equal(wq["flow_lm.transformer.layers.0.linear1.weight"], wnp["flow_lm.transformer.layers.0.linear1.weight"])
    np.testing.assert_array_equal(wq["flow_lm.transformer.layers.0.self_attn.out_proj.weight"],
                                  wnp["flow_lm.transformer.layers.0.self_attn.out_proj.weight"])  # skip list
    W = O.to_torch(wq)
    prompt = synth.make_voice_prompt(24, seed=9)
    tok = synth.make_tokens(9, seed=4)
    frames = 4
    noise = synth.make_noise(frames, seed=6)
    ref = O.generate_segment(W, O.voice_state_from_prompt(W, prompt), tok, noise, frames, 0, float("inf"))
    eng = Engine(wnp, max_slots=2, kv_capacity=128, int8_weights=True)  # the engine quantises the raw weights itself
    voice = eng.voice_from_prompt(prompt)
    s = eng.open_streams([voice], [StreamSpec(tok, frames, 0, 1e30, noise=noise)])
    lat, pcm = [], []
    for f in range(frames):
        if f:
            eng.set_feedback(int(s[0]), ref["latents"][f - 1])
        p, _, l, _ = eng.step(s)
        lat.append(l[0]); pcm.append(p[0])
    eng.close_stream(int(s[0]))
    voice.close()
    eng.close()
    assert np.abs(np.stack(lat) - ref["latents"]).max() <= LAT_TOL
    assert snr(ref["pcm"], np.stack(pcm)) >= SNR_MIN
    # one-byte storage (production) vs f16 copies of the same codes: the whole path must be bit-identical.  One utterance runs
    # its Linear layers on the small-batch GEMV (csrc/gemv.cuh), whose byte and f16 variants deal k to the lanes differently
    # (16 vs 8 codes per chunk: same exact products, another summation order), so bit-identity is asserted on the
    # tensor-core path (gemv_off) and the GEMV run above is held to the oracle like everything else.
    runs = []
    for storage in (True, False):
        eng2 = Engine(wnp, max_slots=2, kv_capacity=128, int8_weights=True, int8_storage=storage, gemv_off=True)
        voice2 = eng2.voice_from_prompt(prompt)
        s2 = eng2.open_streams([voice2], [StreamSpec(tok, frames, 0, 1e30, noise=noise)])
        out = []
        for f in range(frames):
            if f:
                eng2.set_feedback(int(s2[0]), ref["latents"][f - 1])
            p, _, l, _ = eng2.step(s2)
            out.append((l[0].copy(), p[0].copy()))
        runs.append(out)
        if storage:
            eng2.close_stream(int(s2[0]))
            voice2.close()
            eng2.close()
    for (l_a, p_a), (l_b, p_b) in zip(*runs):
        np.testing.assert_array_equal(l_a, l_b)
        np.testing.assert_array_equal(p_a, p_b)
    assert np.abs(np.stack([l for l, _ in runs[0]]) - np.stack(lat)).max() < 1e-2   # GEMV vs tensor-core int8 step
    eng2.close_stream(int(s2[0]))
    voice2.close()
    eng2.close()


def test_continuous_batching_long_form():
    """configs[4] in miniature: requests made of several chunks and pauses share a small batch; chunks of a request
    are emitted in order, pauses add exactly ms*24 zero samples (pause.rs:183-185), and every chunk equals the same
    chunk generated alone."""
    from pocket_tts_b200.engine import StreamSpec
    from pocket_tts_b200.tts_model import BatchScheduler
    eng, _ = engine_for(1234, 0.01)
    voice = eng.voice_from_prompt(synth.make_voice_prompt(16, seed=5))

    def chunk(seed, frames):
        return StreamSpec(synth.make_tokens(5 + seed % 7, seed=seed), frames, 0, 1e30, temp=0.0)

    requests = [
        [("text", chunk(1, 3)), ("pause", 300), ("text", chunk(2, 2))],
        [("text", chunk(3, 5))],
        [("pause", 100), ("text", chunk(4, 2)), ("text", chunk(5, 4)), ("pause", 50)],
        [("text", chunk(6, 1)), ("text", chunk(7, 1)), ("text", chunk(8, 3))],
        [("text", chunk(9, 4)), ("pause", 1000)],
    ]
    got = BatchScheduler(eng, voice, max_batch=3).run(requests)
    # the same job with the device kept one step ahead of the host (PTTS_STEP_AHEAD) and with EOS endings, which the host
    # only learns one step late: same audio, nothing of an overrun frame emitted
    got_ahead = BatchScheduler(eng, voice, max_batch=3).run(requests, ahead=True)
    for a, b in zip(got_ahead, got):
        assert a.shape == b.shape and (snr(b, a) >= 60.0 or np.abs(b).max() == 0)

    def eos_chunk(seed, frames):   # ends at EOS on step 0 + frames_after_eos 2 -> 3 frames, well before max_gen_len
        return StreamSpec(synth.make_tokens(5 + seed % 7, seed=seed), frames, 2, -1e30, temp=0.0)

    eos_requests = [[("text", eos_chunk(11, 9)), ("pause", 100), ("text", chunk(12, 4))], [("text", chunk(13, 6))],
                    [("text", eos_chunk(14, 7)), ("text", eos_chunk(15, 8))]]
    lock = BatchScheduler(eng, voice, max_batch=3).run(eos_requests)
    ahead = BatchScheduler(eng, voice, max_batch=3).run(eos_requests, ahead=True)
    assert [len(x) for x in lock] == [3 * 1920 + 2400 + 4 * 1920, 6 * 1920, 6 * 1920]
    for a, b in zip(ahead, lock):
        assert a.shape == b.shape and snr(b, a) >= 60.0

    def alone(spec):
        s = eng.open_streams([voice], [spec])
        frames = []
        while True:
            pcm, fin, _, _ = eng.step(s)
            frames.append(pcm[0])
            if fin[0]:
                break
        eng.close_stream(int(s[0]))
        return np.concatenate(frames)

    for req, pcm in zip(requests, got):
        want = []
        for kind, arg in req:
            want.append(np.zeros(arg * 24, np.float32) if kind == "pause" else alone(arg))
        want = np.concatenate(want)
        assert pcm.shape == want.shape
        # identical streams in a different batch: only the GEMM tiling (hence f32 summation order) may differ
        assert snr(want, pcm) >= 60.0 or np.abs(want).max() == 0
        zero_mask = want == 0
        assert (pcm[zero_mask] == 0).all()
    voice.close()


@pytest.mark.parametrize("n_streams", [130, 300])
def test_large_ragged_batches_match_oracle(n_streams):
    """Batch sizes beyond one MMA-N tile (130: two feature-side tiles with a ragged tail; 300: the FlowLM GEMMs move
    to the activation-as-M tiling): a few probe streams inside the batch must still equal their own oracle runs."""
    from oracle import ptts_oracle as O
    from pocket_tts_b200.engine import Engine, StreamSpec
    for e in _cache.values():
        e[0].close()
    _cache.clear()
    wnp = synth.make_weights(1234)
    W = O.to_torch(wnp)
    eng = Engine(wnp, max_slots=n_streams, kv_capacity=64)
    prompt = synth.make_voice_prompt(12, seed=2)
    voice = eng.voice_from_prompt(prompt)
    ov = O.voice_state_from_prompt(W, prompt)
    frames = 2
    specs = [StreamSpec(synth.make_tokens(3 + i % 5, seed=i), frames, 0, 1e30, noise=synth.make_noise(frames, seed=1000 + i))
             for i in range(n_streams)]
    probes = [0, n_streams // 2, n_streams - 1]
    refs = {i: O.generate_segment(W, ov, specs[i].tokens, specs[i].noise, frames, 0, float("inf")) for i in probes}
    slots = eng.open_streams([voice] * n_streams, specs)
    lat, pcm = [], []
    for f in range(frames):
        if f:
            for i in probes:
                eng.set_feedback(int(slots[i]), refs[i]["latents"][f - 1])
        p, fin, l, _ = eng.step(slots)
        lat.append(l); pcm.append(p)
    assert fin.all()
    for i in probes:
        got_lat = np.stack([lat[f][i] for f in range(frames)])
        got_pcm = np.stack([pcm[f][i] for f in range(frames)])
        assert np.abs(got_lat[0] - refs[i]["latents"][0]).max() <= LAT_TOL
        assert np.abs(got_lat - refs[i]["latents"]).max() <= LAT_TOL
        assert snr(refs[i]["pcm"], got_pcm) >= SNR_MIN
    voice.close()
    eng.close()


@pytest.mark.parametrize("lsd", [2, 3, 6])
def test_flow_head_euler_steps_vs_oracle(lsd):
    """lsd_decode_steps other than the goldens' 1 and 4 (reference flow_lm.rs:7-22): 2 and 3 run as one flow-head launch with the
    modulations of all steps from one Linear (modules/mlp.rs:322-368), 6 exceeds the modulation scratch and takes one launch
    per step.  Teacher-forced, two streams (the second row sits in a different half of the 16-row chunk's registers)."""
    from oracle import ptts_oracle as O
    from pocket_tts_b200.engine import StreamSpec
    eng, w = engine_for(1234, 1.0, max_slots=16)
    W = O.to_torch(w)
    prompt = synth.make_voice_prompt(9, seed=21)
    voice = eng.voice_from_prompt(prompt)
    ov = O.voice_state_from_prompt(W, prompt)
    frames = 3
    eng.set_lsd_steps(lsd)
    try:
        toks = [synth.make_tokens(6 + i, seed=60 + i) for i in range(9)]
        noises = [synth.make_noise(frames, seed=80 + i) for i in range(9)]
        refs = {i: O.generate_segment(W, ov, toks[i], noises[i], frames, 0, float("inf"), lsd_steps=lsd, decode_audio=False) for i in (0, 8)}
        slots = eng.open_streams([voice] * 9, [StreamSpec(toks[i], frames, 0, 1e30, noise=noises[i]) for i in range(9)])
        for f in range(frames):
            if f:
                for i in (0, 8):
                    eng.set_feedback(int(slots[i]), refs[i]["latents"][f - 1])
            _, _, lat, _ = eng.step(slots)
            for i in (0, 8):
                err = float(np.abs(lat[i] - refs[i]["latents"][f]).max())
                assert err <= LAT_TOL, f"lsd {lsd} stream {i} frame {f}: latent max-abs {err:.3e}"
        eng.close_streams(slots)
    finally:
        eng.set_lsd_steps(1)
        voice.close()
