"""Codec groups (ptts_engine_set_codec_group): the Mimi decoder + SEANet run once per `frames` latents of the same batch.
The reference decodes every latent inside the frame loop (tts_model.rs:1033-1047); its decoder is streaming
(conv.rs:90-136,219-267; attention.rs:167-264), so a group is only a longer chunk of each stream: the PCM must be
BIT-identical to the per-frame passes, whatever cuts the groups (full groups, PCM fetched early, a batch that changes,
streams that end at different frames)."""
import numpy as np
import pytest

from pocket_tts_b200 import synth
from pocket_tts_b200.engine import Engine, StreamSpec, NativeScheduler

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def weights():
    return synth.make_weights(1234)


def _specs(n, frames, seed0=0):
    return [StreamSpec(synth.make_tokens(5 + (i % 5), seed=300 + i), frames, 0, 1e30, noise=synth.make_noise(frames, seed=seed0 + i)) for i in range(n)]


def _run_ahead(eng, voice, n, frames, fetch_lag):
    """begin(i, ahead) ... flags(i-1) ... pcm(i-fetch_lag): with lag >= group the groups fill, with lag 1 every fetch cuts one."""
    slots = eng.open_streams([voice] * n, _specs(n, frames))
    tickets, pcm, lat = [], {}, []
    for i in range(frames):
        tickets.append(eng.step_begin(slots, ahead=i > 0))
        if i >= 1:
            lat.append(eng.step_flags(tickets[i - 1])[1])
        if i >= fetch_lag:
            pcm[i - fetch_lag] = eng.step_pcm(tickets[i - fetch_lag])
    lat.append(eng.step_flags(tickets[-1])[1])
    for i in range(max(0, frames - fetch_lag), frames):
        pcm[i] = eng.step_pcm(tickets[i])
    eng.close_streams(slots)
    return np.stack([pcm[i] for i in range(frames)]), np.stack(lat)


@pytest.mark.parametrize("group", [2, 4])
def test_group_pcm_bit_identical(weights, group):
    n, frames = 5, 23   # crosses the 250-position Mimi window (16 positions per frame); 23 is not a multiple of the group
    eng = Engine(weights, max_slots=8, kv_capacity=64)
    voice = eng.voice_from_prompt(synth.make_voice_prompt(11, seed=7))
    ref_pcm, ref_lat = _run_ahead(eng, voice, n, frames, 1)
    eng.set_codec_group(group)
    for lag in (group, 1, 2):   # full groups; every fetch cuts the group short; in between
        pcm, lat = _run_ahead(eng, voice, n, frames, lag)
        assert np.array_equal(lat, ref_lat)
        assert np.array_equal(pcm, ref_pcm), f"group {group}, fetch lag {lag}: max diff {np.abs(pcm - ref_pcm).max()}"
    eng.set_codec_group(1)
    pcm, lat = _run_ahead(eng, voice, n, frames, 1)
    assert np.array_equal(pcm, ref_pcm)
    voice.close(); eng.close()


def test_group_with_changing_batch(weights):
    """Streams of different lengths leave the batch one by one: every change of composition flushes the partial group."""
    lens = [3, 6, 7, 12]
    out = {}
    for group in (1, 2, 4):
        eng = Engine(weights, max_slots=4, kv_capacity=64, codec_group=group)
        voice = eng.voice_from_prompt(synth.make_voice_prompt(9, seed=3))
        specs = [StreamSpec(synth.make_tokens(6, seed=50 + i), L, 0, 1e30, noise=synth.make_noise(L, seed=70 + i)) for i, L in enumerate(lens)]
        slots = list(eng.open_streams([voice] * len(lens), specs))
        frames = [[] for _ in lens]
        active = list(range(len(lens)))
        pend = []   # (ticket, rows) whose PCM has not been fetched
        while active:
            t = eng.step_begin(np.array([slots[i] for i in active], np.int32))
            fin, _, _ = eng.step_flags(t)
            pend.append((t, list(active)))
            if len(pend) > 2 or fin.any():
                for tk, rows in pend:
                    p = eng.step_pcm(tk)
                    for r, i in enumerate(rows):
                        frames[i].append(p[r])
                pend = []
            for r in [r for r, f in enumerate(fin) if f][::-1]:
                eng.close_stream(int(slots[active[r]]))
                active.pop(r)
        for tk, rows in pend:
            p = eng.step_pcm(tk)
            for r, i in enumerate(rows):
                frames[i].append(p[r])
        out[group] = [np.concatenate(f) for f in frames]
        assert [len(f) // 1920 for f in out[group]] == lens
        voice.close(); eng.close()
    for group in (2, 4):
        for a, b in zip(out[1], out[group]):
            assert np.array_equal(a, b)


def test_group_device_steps_and_scheduler(weights):
    """ptts_step_device fills groups on its own (ptts_sync decodes the rest); the native scheduler gives the same PCM."""
    eng = Engine(weights, max_slots=8, kv_capacity=64)
    voice = eng.voice_from_prompt(synth.make_voice_prompt(9, seed=3))
    reqs = [[("text", StreamSpec(synth.make_tokens(6, seed=10 + i), 5 + i, 0, 1e30, noise=synth.make_noise(5 + i, seed=20 + i))), ("pause", 40),
             ("text", StreamSpec(synth.make_tokens(4, seed=30 + i), 4, 0, 1e30, noise=synth.make_noise(4, seed=40 + i)))] for i in range(6)]
    ref = NativeScheduler(eng, voice, 4).run(reqs)
    eng.set_codec_group(2)
    got = NativeScheduler(eng, voice, 4).run(reqs)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    # device-resident steps: 5 frames in groups of 2, the last frame decoded by sync; the final PCM frame is the 5th
    slots = eng.open_streams([voice] * 3, _specs(3, 5, seed0=900))
    for _ in range(5):
        eng.step_device(slots)
    eng.sync()
    last = np.stack([eng.debug_read("pcm", r, 1920) for r in range(3)])
    eng.close_streams(slots)
    eng.set_codec_group(1)
    slots = eng.open_streams([voice] * 3, _specs(3, 5, seed0=900))
    for _ in range(5):
        pcm = eng.step(slots)[0]
    eng.close_streams(slots)
    assert np.array_equal(last, pcm)
    voice.close(); eng.close()


def test_set_codec_group_rejects_bad_values_and_steps_in_flight(weights):
    from pocket_tts_b200._lib import PttsError
    eng = Engine(weights, max_slots=4, kv_capacity=64)
    voice = eng.voice_from_prompt(synth.make_voice_prompt(9, seed=3))
    with pytest.raises(PttsError):
        eng.set_codec_group(3)
    slots = eng.open_streams([voice], _specs(1, 4))
    t = eng.step_begin(slots)
    with pytest.raises(PttsError):          # a ticket with unfetched results: the codec scratch may not be re-sized now
        eng.set_codec_group(2)
    eng.step_flags(t)
    eng.step_pcm(t)
    eng.set_codec_group(2)                  # fine between steps, with the stream still open
    p = eng.step(slots)[0]
    assert np.isfinite(p).all()
    eng.close_streams(slots)
    voice.close(); eng.close()
