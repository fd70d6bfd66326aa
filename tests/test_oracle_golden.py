"""Pin the CPU oracle against outputs of the unmodified reference PyTorch package
(tests/golden/*.npz, produced by tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import ptts_oracle as O
from pocket_tts_b200 import synth

CASES = ["cfg1_lsd1", "cfg3_lsd4", "stress_ls05"]
_W = {}


def weights_for(seed, ls):
    key = (int(seed), float(ls))
    if key not in _W:
        _W.clear()
        _W[key] = O.to_torch(synth.make_weights(key[0], layer_scale=key[1]))
    return _W[key]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("kind", ["erf", "tanh"])
def test_oracle_matches_reference_package(golden_dir, case, kind):
    g = np.load(golden_dir / f"{case}.npz")
    W = weights_for(g["weight_seed"], g["layer_scale"])
    voice = O.voice_state_from_prompt(W, synth.make_voice_prompt(int(g["voice_rows"]), seed=7), kind)
    frames = g[f"{kind}_latents"].shape[0]
    r = O.generate_segment(W, voice, g["tokens"], g["noise"], frames, 0, float("inf"),
                           lsd_steps=int(g["lsd_steps"]), gelu_kind=kind)
    assert r["frames"] == frames
    # f32 vs f32, different summation order only
    np.testing.assert_allclose(r["latents"], g[f"{kind}_latents"], atol=2e-4, rtol=0)
    np.testing.assert_allclose(r["eos_logits"], g[f"{kind}_eos_logits"], atol=2e-4, rtol=0)
    np.testing.assert_allclose(r["pcm"], g[f"{kind}_pcm"], atol=5e-4, rtol=0)


@pytest.mark.parametrize("case", CASES)
def test_oracle_decoder_stages_frame0(golden_dir, case):
    """Same stage chain as the reference's assets/ref_decoder_intermediates.safetensors
    (quantized -> after_upsample -> after_decoder_transformer -> audio)."""
    g = np.load(golden_dir / f"{case}.npz")
    W = weights_for(g["weight_seed"], g["layer_scale"])
    trace = {}
    pcm = O.mimi_decode_step(W, torch.from_numpy(g["tanh_latents"][0]), O.MimiState(), "tanh", trace)
    np.testing.assert_allclose(trace["mimi.quantized"].numpy(), g["tanh_stage_quantized"], atol=1e-5)
    np.testing.assert_allclose(trace["mimi.after_upsample"].numpy(), g["tanh_stage_after_upsample"], atol=1e-5)
    np.testing.assert_allclose(trace["mimi.after_decoder_transformer"].numpy(),
                               g["tanh_stage_after_decoder_transformer"], atol=1e-4)
    np.testing.assert_allclose(pcm.numpy(), g["tanh_pcm"][0], atol=5e-4)
