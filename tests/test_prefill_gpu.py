"""Text / voice prefill (reference tts_model.rs:944-964, modules/sdpa.rs:36-171 at Lq > 1): the tensor-core prefill attention
(flowlm_attn_prefill_mma_kernel: K / V of a sequence staged in shared memory once per 64 query rows, mma.sync flash pass)
against the row-per-CTA SIMT kernel (PTTS_PREFILL_TILE=0, f32 q and probabilities) on the same inputs: the KV cache the
prefill leaves is the same, so the first generated frames agree to f16 rounding of q / P.  Parity with the CPU oracle for
the tensor-core path itself is every test of test_parity_gpu.py / test_bench_shape_gpu.py (it is the default)."""
import numpy as np
import pytest

from pocket_tts_b200 import synth

pytestmark = pytest.mark.gpu


def _run(monkeypatch, tiled, voice_rows, token_counts, frames=3):
    from pocket_tts_b200.engine import Engine, StreamSpec
    monkeypatch.setenv("PTTS_PREFILL_TILE", "1" if tiled else "0")
    eng = Engine(synth.make_weights(21), max_slots=8, kv_capacity=256)
    voice = eng.voice_from_prompt(synth.make_voice_prompt(voice_rows, seed=5))
    n = len(token_counts)
    rng = np.random.default_rng(1)
    noise = (rng.standard_normal((n, frames, 32)) * np.sqrt(0.7)).astype(np.float32)
    feed = (rng.standard_normal((n, frames, 32)) * 0.5).astype(np.float32)
    specs = [StreamSpec(synth.make_tokens(t, seed=100 + i), frames, 0, 1e30, noise=noise[i]) for i, t in enumerate(token_counts)]
    slots = eng.open_streams([voice] * n, specs)
    res = []
    for f in range(frames):
        if f:
            for i in range(n):
                eng.set_feedback(int(slots[i]), feed[i, f - 1])
        res.append(eng.step(slots))
    out = (np.stack([r[2] for r in res]), np.stack([r[0] for r in res]), np.stack([r[3] for r in res]))
    eng.close_streams(slots)
    voice.close()
    eng.close()
    return out


@pytest.mark.parametrize("voice_rows,token_counts", [
    (87, [40, 40, 40]),            # the bench shape: one 40-row tile per stream, three warps of it busy
    (12, [1, 17, 5, 33, 16, 2]),   # ragged: single-row tiles, 16 / 17 rows (a warp with one row), 33 rows
    (130, [50]),                   # a voice of more than 8 key blocks, the longest text chunk
    (200, [50, 3]),                # a voice prefilled in four 64-row tiles of its own
])
def test_tensor_core_prefill_attention_matches_simt(monkeypatch, voice_rows, token_counts):
    lat, pcm, logit = _run(monkeypatch, True, voice_rows, token_counts)
    lat0, pcm0, logit0 = _run(monkeypatch, False, voice_rows, token_counts)
    assert np.isfinite(lat).all() and np.isfinite(pcm).all()
    assert np.abs(lat - lat0).max() < 5e-3, np.abs(lat - lat0).max()
    assert np.abs(logit - logit0).max() < 5e-3
    snr = 10 * np.log10((pcm0.astype(np.float64) ** 2).sum() / max(((pcm0.astype(np.float64) - pcm) ** 2).sum(), 1e-30))
    assert snr >= 50.0, snr   # the waveform bar against the oracle is 40 dB
