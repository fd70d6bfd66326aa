"""The fused SEANet tail (csrc/seanet_tail.cuh: last ResBlock + final 64 -> 1 conv + i16 packing in one kernel, reference
models/seanet.rs:82-88,379-392) against the three launches it replaces: same MMAs in the same order, same order of additions
in the final conv, so the PCM must be bit-identical -- across frames (the two rows of left context travel through the slot
state), for streams in any batch row, and in codec groups (tiles then straddle frame boundaries)."""
import os

import numpy as np
import pytest

from pocket_tts_b200 import synth
from pocket_tts_b200.engine import Engine, StreamSpec

pytestmark = pytest.mark.gpu


def _run(weights, fused: bool, group: int, n=5, frames=7, i16=False):
    os.environ["PTTS_SEANET_TAIL"] = "1" if fused else "0"
    try:
        eng = Engine(weights, max_slots=max(8, n), kv_capacity=64, codec_group=group)
    finally:
        del os.environ["PTTS_SEANET_TAIL"]
    voice = eng.voice_from_prompt(synth.make_voice_prompt(9, seed=3))
    specs = [StreamSpec(synth.make_tokens(5 + i % 7, seed=40 + i), frames, 0, 1e30, noise=synth.make_noise(frames, seed=90 + i)) for i in range(n)]
    slots = eng.open_streams([voice] * n, specs)
    tickets = [eng.step_begin(slots, ahead=f > 0, i16=i16) for f in range(1)]
    out = []
    for f in range(frames):
        if f + 1 < frames:
            nxt = eng.step_begin(slots, ahead=True, i16=i16)
            tickets.append(nxt)
        eng.step_flags(tickets[f])
        if f >= group - 1:
            k = f - (group - 1)
            out.append(eng.step_pcm_i16(tickets[k]) if i16 else eng.step_pcm(tickets[k]))
    for k in range(len(out), frames):
        out.append(eng.step_pcm_i16(tickets[k]) if i16 else eng.step_pcm(tickets[k]))
    eng.close_streams(slots)
    voice.close(); eng.close()
    return np.stack(out)


def test_fused_tail_bit_identical_to_separate_launches():
    w = synth.make_weights(1234)
    ref = _run(w, False, 1)
    assert np.isfinite(ref).all() and np.abs(ref).max() > 0.1
    for group in (1, 2):
        got = _run(w, True, group)
        assert np.array_equal(got, ref), f"group {group}: max diff {np.abs(got - ref).max()}"
    assert np.array_equal(_run(w, True, 1, i16=True), _run(w, False, 1, i16=True))


@pytest.mark.parametrize("n", [24, 64])
def test_fused_tail_many_tiles_per_cta(n):
    """16 tiles per stream and frame: 24 streams = 384 tiles, 64 streams = 1024 tiles on at most 132 persistent CTAs, so every CTA
    walks several tiles (stage reuse, barrier phases, the accumulator hand-back) -- the batch-of-5 case above never does."""
    w = synth.make_weights(1234)
    ref = _run(w, False, 1, n=n, frames=3)
    got = _run(w, True, 1, n=n, frames=3)
    assert np.array_equal(got, ref), f"max diff {np.abs(got - ref).max()}"
    if n == 24:
        assert np.array_equal(_run(w, True, 4, n=n, frames=5), _run(w, False, 1, n=n, frames=5))
