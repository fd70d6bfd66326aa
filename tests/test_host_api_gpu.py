"""Host-facing behaviour of the C ABI added in round 2: packed open / sync-free close, i16 PCM packed on the device,
voice-state safetensors save / load, the native (C++) continuous-batching scheduler, interleaved stream iterators, the
LSD-step switch under captured graphs, the device noise generator's distribution, and the f16 overflow counter."""
import numpy as np
import pytest

from pocket_tts_b200 import synth

pytestmark = pytest.mark.gpu

LAT_TOL = 1e-2


@pytest.fixture(scope="module")
def rig():
    from pocket_tts_b200.engine import Engine
    w = synth.make_weights(1234)
    eng = Engine(w, max_slots=16, kv_capacity=256)
    voice = eng.voice_from_prompt(synth.make_voice_prompt(23, seed=7))
    yield eng, voice, w
    voice.close()
    eng.close()


def _spec(i, frames, ntok=7, noise=True):
    from pocket_tts_b200.engine import StreamSpec
    return StreamSpec(synth.make_tokens(ntok, seed=100 + i), frames, 0, 1e30, temp=0.7, seed=i,
                      noise=synth.make_noise(frames, seed=200 + i) if noise else None)


def test_lsd_switch_with_cached_graphs_keeps_parity(rig):
    """the graphs captured for (batch, lsd) bake the address of the time embeddings; lsd 1 -> 4 -> (an open with
    injected noise, i.e. an allocation in between) -> 1 at the same batch size must replay against live memory."""
    eng, voice, _ = rig

    def run(frames=3):
        s = eng.open_streams([voice], [_spec(0, frames)])
        out = [eng.step(s)[2][0].copy() for _ in range(frames)]
        eng.close_stream(int(s[0]))
        return np.stack(out)

    a1 = run()
    eng.set_lsd_steps(4)
    a4 = run()
    extra = eng.open_streams([voice], [_spec(5, 40, ntok=3)])   # a noise buffer is allocated here
    eng.set_lsd_steps(1)
    b1 = run()
    eng.set_lsd_steps(4)
    b4 = run()
    eng.set_lsd_steps(1)
    eng.close_stream(int(extra[0]))
    np.testing.assert_array_equal(a1, b1)
    np.testing.assert_array_equal(a4, b4)
    assert np.abs(a1 - a4).max() > 1e-3


def test_close_and_step_state_rules(rig):
    """A slot listed by a step whose flags are unfetched cannot be closed; a slot may not appear twice in one batch."""
    from pocket_tts_b200 import _lib
    eng, voice, _ = rig
    s = eng.open_streams([voice] * 2, [_spec(0, 6), _spec(1, 6)])
    with pytest.raises(_lib.PttsError) as ei:
        eng.step(np.array([s[0], s[0]], np.int32))
    assert ei.value.code == -1
    t = eng.step_begin(s)
    with pytest.raises(_lib.PttsError) as ei:
        eng.close_stream(int(s[0]))
    assert ei.value.code == -4
    eng.step_flags(t)
    eng.step_pcm(t)
    eng.close_streams(s)           # batch close, host bookkeeping only
    s2 = eng.open_streams([voice], [_spec(0, 2)])   # the recycled slot works at once (ordered on the device)
    assert np.isfinite(eng.step(s2)[0]).all()
    eng.close_stream(int(s2[0]))


def test_i16_pcm_is_the_reference_packing(rig):
    """audio.rs:129-146: clamp to [-1, 1], * 32767, truncating cast -- done by the last SEANet conv on the device."""
    eng, voice, _ = rig
    frames = 3
    outs = {}
    for i16 in (False, True):
        s = eng.open_streams([voice] * 3, [_spec(i, frames) for i in range(3)])
        got = []
        for _ in range(frames):
            t = eng.step_begin(s, i16=i16)
            eng.step_flags(t)
            got.append(eng.step_pcm_i16(t) if i16 else eng.step_pcm(t))
        outs[i16] = np.stack(got)
        eng.close_streams(s)
    want = (np.clip(outs[False] * 40.0, -1.0, 1.0) * 32767.0).astype(np.int16)  # what the packing of a louder frame would be
    assert outs[True].dtype == np.int16 and outs[True].shape == outs[False].shape
    np.testing.assert_array_equal(outs[True], (np.clip(outs[False], -1.0, 1.0) * np.float32(32767.0)).astype(np.int16))
    assert np.abs(outs[True]).max() > 0 and want.shape == outs[True].shape


def test_voice_state_safetensors_round_trip(rig, tmp_path):
    """tts_model.rs:467-487: the voice file is `audio_prompt` f32 [1,T,1024]; the optional KV snapshot skips the prefill."""
    from pocket_tts_b200.tts_model import read_safetensors
    eng, voice, _ = rig
    prompt = synth.make_voice_prompt(23, seed=7)

    def gen(v):
        s = eng.open_streams([v], [_spec(3, 3)])
        out = [eng.step(s) for _ in range(3)]
        eng.close_stream(int(s[0]))
        return np.stack([o[2][0] for o in out]), np.stack([o[0][0] for o in out])

    base = gen(voice)
    for include_kv in (False, True):
        path = tmp_path / f"voice_{int(include_kv)}.safetensors"
        voice.save(path, include_kv=include_kv)
        t = read_safetensors(path)
        assert t["audio_prompt"].shape == (1, 23, 1024)
        np.testing.assert_array_equal(t["audio_prompt"][0], prompt)
        assert ("flow_lm_kv" in t) == include_kv
        v2 = eng.voice_load(path)
        assert len(v2) == 23
        got = gen(v2)
        np.testing.assert_array_equal(got[0], base[0])
        np.testing.assert_array_equal(got[1], base[1])
        v2.close()
    (tmp_path / "bad.safetensors").write_bytes(b"\x10\x00\x00\x00\x00\x00\x00\x00" + b'{"x":{"dtype":"F32","shape":[1],"data_offsets":[0,4]}}'[:16])
    from pocket_tts_b200 import _lib
    with pytest.raises(_lib.PttsError):
        eng.voice_load(tmp_path / "bad.safetensors")


def test_native_scheduler_matches_python_scheduler(rig):
    """ptts_sched_* (C++) == tts_model.BatchScheduler (Python) on a long-form workload with pauses, more requests than
    batch rows, ragged chunk lengths and EOS endings; and its i16 output is the packing of its f32 output."""
    from pocket_tts_b200.engine import NativeScheduler, StreamSpec
    from pocket_tts_b200.tts_model import BatchScheduler
    eng, voice, _ = rig
    reqs = []
    for r in range(9):
        segs = []
        for c in range(1 + r % 3):
            frames = 3 + (r + c) % 4
            segs.append(("text", StreamSpec(synth.make_tokens(4 + (r * 3 + c) % 9, seed=10 * r + c), frames, 1 + c % 2, -0.5 if (r + c) % 2 else 1e30,
                                            temp=0.7, seed=r, noise=synth.make_noise(frames, seed=50 * r + c))))
            if c % 2 == 0:
                segs.append(("pause", 40 + 10 * r))
        reqs.append(segs)
    py = BatchScheduler(eng, voice, max_batch=4).run(reqs)
    py_ahead = BatchScheduler(eng, voice, max_batch=4).run(reqs, ahead=True)
    ns = NativeScheduler(eng, voice, max_batch=4)
    nat = ns.run(reqs)
    steps = ns.steps
    ns.close()
    ns2 = NativeScheduler(eng, voice, max_batch=4)
    views = ns2.run(reqs, view=True)      # ptts_sched_result_view: the same samples read in place, no copy
    for a, v in zip(nat, views):
        assert not v.flags.writeable and not v.flags.owndata
        np.testing.assert_array_equal(a, v)
    del views
    ns2.close()
    assert steps > 0
    for a, b, c in zip(py, py_ahead, nat):
        # the same policy (one step ahead of the host) batches the same rows at every step: bit-identical
        np.testing.assert_array_equal(b, c)
        # lock step batches differently around EOS endings (a finished row rides one more step as an overrun row), and a
        # GEMM's split-K summation order depends on the batch size: same frames, same counts, equal to rounding
        assert a.shape == c.shape
        assert np.abs(a - c).max() < 2e-2
    nat16 = BatchScheduler(eng, voice, max_batch=4).run(reqs, native=True, i16=True)
    for a, b in zip(nat, nat16):
        np.testing.assert_array_equal((np.clip(a, -1.0, 1.0) * np.float32(32767.0)).astype(np.int16), b)


def test_interleaved_stream_iterators(rig):
    """The reference's generate_stream iterators are independent (each owns a clone of the state, tts_model.rs:894-913): two
    iterators of one model consumed alternately give what each gives alone."""
    from pocket_tts_b200.tts_model import TTSModel
    _, _, w = rig
    m = TTSModel(w, max_slots=4, kv_capacity=128)
    voice = m.get_voice_state_from_prompt_tensor(synth.make_voice_prompt(12, seed=3))
    args = [(synth.make_tokens(5, seed=1), 6, 3, synth.make_noise(6, seed=2)), (synth.make_tokens(8, seed=4), 5, 3, synth.make_noise(5, seed=5))]
    alone = [np.concatenate(list(m.generate_stream_tokens(t, voice, n, fae, noise=nz)), axis=2) for t, n, fae, nz in args]
    its = [m.generate_stream_tokens(t, voice, n, fae, noise=nz) for t, n, fae, nz in args]
    got = [[], []]
    live = [0, 1]
    while live:
        for i in list(live):
            try:
                got[i].append(next(its[i]))
            except StopIteration:
                live.remove(i)
    for i in range(2):
        np.testing.assert_array_equal(np.concatenate(got[i], axis=2), alone[i])
    # an iterator dropped half way releases its slot and its steps in flight
    it = m.generate_stream_tokens(args[0][0], voice, args[0][1], args[0][2], noise=args[0][3])
    next(it); next(it)
    it.close()
    again = np.concatenate(list(m.generate_stream_tokens(args[1][0], voice, args[1][1], args[1][2], noise=args[1][3])), axis=2)
    np.testing.assert_array_equal(again, alone[1])
    voice.close()
    m.close()


def test_device_noise_generator_distribution():
    """flow_lm.rs:39-65 draws N(0, temp) i.i.d.; the device generator (hash of seed / frame / lane -> Box-Muller) must be
    standard normal, uncorrelated across lanes, frames and seeds."""
    from scipy import stats
    from pocket_tts_b200.engine import device_noise
    z = device_noise(seed=12345, frames=8192)
    assert z.shape == (8192, 32) and np.isfinite(z).all()
    flat = z.reshape(-1).astype(np.float64)
    n = flat.size
    assert abs(flat.mean()) < 4 / np.sqrt(n)
    assert abs(flat.var() - 1) < 4 * np.sqrt(2 / n)
    assert abs(stats.skew(flat)) < 0.03 and abs(stats.kurtosis(flat)) < 0.06
    assert stats.kstest(flat[:50000], "norm").pvalue > 1e-3
    c = np.corrcoef(z.T)
    assert np.abs(c - np.eye(32)).max() < 0.06                       # lanes of a frame
    assert abs(np.corrcoef(z[:-1].reshape(-1), z[1:].reshape(-1))[0, 1]) < 0.01   # consecutive frames
    z2 = device_noise(seed=12346, frames=8192)
    assert abs(np.corrcoef(flat, z2.reshape(-1))[0, 1]) < 0.01        # neighbouring seeds
    np.testing.assert_array_equal(z, device_noise(seed=12345, frames=8192))
    assert np.abs(flat).max() < 6.5


def test_f16_overflow_counter_is_zero_on_a_normal_run(rig):
    eng, voice, _ = rig
    s = eng.open_streams([voice] * 2, [_spec(0, 2), _spec(1, 2)])
    eng.step(s); eng.step(s)
    assert eng.f16_overflow_count() == 0
    eng.close_streams(s)
