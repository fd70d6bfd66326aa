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


def encoder_weights(g):
    w = dict(synth.make_weights(1234, layer_scale=0.01))
    w.update(synth.make_encoder_weights(int(g["enc_seed"]), layer_scale=float(g["enc_layer_scale"])))
    return O.to_torch(w)


@pytest.mark.parametrize("kind", ["erf", "tanh"])
def test_oracle_voice_cloning_encoder_matches_reference_package(golden_dir, kind):
    """SURVEY 8f N1 (oracle side): PCM -> SEANetEncoder -> encoder transformer -> downsample -> speaker_proj equals the
    unmodified package's `_encode_audio` (22 frames: 352 transformer positions cross the 250-position window; the
    prompt length is ragged, so the zero end-padding of tts_model.rs:514-527 is exercised)."""
    g = np.load(golden_dir / "enc_pcm22.npz")
    W = encoder_weights(g)
    pcm = synth.make_pcm(int(g["n_samples"]), seed=int(g["pcm_seed"]))
    got = O.audio_prompt_from_pcm(W, pcm, kind).numpy()
    want = g[f"{kind}_audio_prompt"]
    assert got.shape == want.shape == (22, 1024)
    np.testing.assert_allclose(got, want, atol=3e-4, rtol=0)


def test_oracle_encoder_chunking_follows_the_reference():
    """RS encodes long prompts in chunks with one carried state (tts_model.rs:528-541) and passes step = 0 for every
    chunk, so only the replicate padding of the downsample restarts at a chunk boundary (conv.rs:114-123): every
    other layer is exactly streaming, and the frames after a boundary differ from the unchunked run only through that
    one convolution's first window."""
    g_w = dict(synth.make_weights(1234, layer_scale=0.01))
    g_w.update(synth.make_encoder_weights(4321, layer_scale=0.5))
    W = O.to_torch(g_w)
    pcm = synth.make_pcm(6 * 1920, seed=5)
    whole = O.audio_prompt_from_pcm(W, pcm, "tanh", chunk_frames=6).numpy()
    halves = O.audio_prompt_from_pcm(W, pcm, "tanh", chunk_frames=3).numpy()
    np.testing.assert_allclose(halves[:3], whole[:3], atol=1e-5)      # before the boundary: identical
    assert np.abs(halves[3] - whole[3]).max() > 1e-3                   # the boundary frame sees replicated padding
    np.testing.assert_allclose(halves[4:], whole[4:], atol=2e-4)      # later frames: streaming again
    assert [O.voice_prompt_chunk_frames(n) for n in (1, 87, 120, 121, 600, 601, 1800, 1801)] == [1, 87, 120, 120, 120, 180, 180, 240]
