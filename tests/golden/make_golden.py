"""Generate tests/golden/*.npz by running the UNMODIFIED reference PyTorch package.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports `pocket_tts` from /root/reference/python-reference (the implementation
the Rust crate is itself parity-tested against, crates/pocket-tts/tests/parity_tests.rs),
builds TTSModel through the package's own constructor path
(`TTSModel._from_pydantic_config_with_weights`, models/tts_model.py:83-170), loads
the seeded random-init weights of pocket_tts_b200/synth.py, and drives the package's
own streaming calls exactly as its extraction scripts do
(scripts/extract_decoder_refs.py:41-99): voice prompt -> text prompt -> AR frames ->
Mimi decode.  Two things are injected from outside, nothing in the package is edited:
  * the per-frame noise x_0 (the package draws it with torch.nn.init.normal_,
    models/flow_lm.py:131-134) is replaced by our seeded tensor so both sides see
    identical noise;
  * for the "tanh" goldens only, F.gelu is switched to approximate="tanh", which is
    deviation D1 of the Rust crate (models/transformer.rs:85).
The "erf" goldens are therefore outputs of the untouched reference.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference/python-reference")

from pocket_tts_b200 import synth  # noqa: E402

torch.set_grad_enabled(False)


def build_reference_model(weights: dict[str, np.ndarray], lsd_steps: int, allow_missing_encoder: bool = True):
    import pocket_tts.conditioners.text as ctext

    # The tokenizer download is the only constructor side effect; token IDs are fed directly.
    ctext.SentencePieceTokenizer.__init__ = lambda self, nbins, tokenizer_path: None
    from pocket_tts.models.tts_model import TTSModel
    from pocket_tts.utils.config import load_config

    cfg = load_config(Path("/root/reference/python-reference/pocket_tts/config/b6369a24.yaml"))
    cfg.weights_path = None
    cfg.weights_path_without_voice_cloning = None
    model = TTSModel._from_pydantic_config_with_weights(cfg, 0.7, lsd_steps, None, -4.0)
    sd = {k: torch.from_numpy(v.copy()) for k, v in weights.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    # only the Mimi *encoder* side (voice cloning from PCM) is not covered by synth weights
    assert all(m.startswith(("mimi.encoder", "mimi.downsample", "flow_lm.flow_net.time_embed")) and
               ("freqs" in m or m.startswith("mimi.")) for m in missing), missing
    if not allow_missing_encoder:
        assert not any(m.startswith("mimi.") for m in missing), missing
    model.eval()
    return model


def run_reference(model, prompt, tokens, noise, frames, gelu_kind, teacher=None):
    from pocket_tts.modules.stateful_module import increment_steps, init_states
    import torch.nn.functional as F

    orig_gelu, orig_normal = F.gelu, torch.nn.init.normal_
    logits = []
    hook = model.flow_lm.out_eos.register_forward_hook(lambda m, i, o: logits.append(float(o.reshape(-1)[-1])))
    step_box = {"i": 0}

    def fake_normal(t, mean=0.0, std=1.0):
        t.copy_(torch.from_numpy(noise[step_box["i"]]).view_as(t))
        return t

    if gelu_kind == "tanh":
        F.gelu = lambda x: orig_gelu(x, approximate="tanh")
    torch.nn.init.normal_ = fake_normal
    try:
        state = init_states(model.flow_lm, batch_size=1, sequence_length=1000)
        model._run_flow_lm_and_increment_step(model_state=state, audio_conditioning=torch.from_numpy(prompt)[None])
        model._run_flow_lm_and_increment_step(model_state=state, text_tokens=torch.from_numpy(tokens.astype(np.int64))[None])
        logits.clear()
        mimi_state = init_states(model.mimi, batch_size=1, sequence_length=1000)
        backbone = torch.full((1, 1, 32), float("nan"))
        lat, pcm, stages = [], [], {}
        for step in range(frames):
            step_box["i"] = step
            nxt, _ = model._run_flow_lm_and_increment_step(model_state=state, backbone_input_latents=backbone)
            lat.append(nxt.reshape(32).clone())
            dec_in = nxt * model.flow_lm.emb_std + model.flow_lm.emb_mean
            quant = model.mimi.quantizer(dec_in.transpose(-1, -2))
            if step == 0:
                # stage-by-stage like scripts/extract_decoder_refs.py:70-99
                up = model.mimi.upsample(quant, mimi_state)
                (tr,) = model.mimi.decoder_transformer(up, mimi_state)
                audio = model.mimi.decoder(tr, mimi_state)
                stages = dict(quantized=quant[0, :, 0].numpy().copy(), after_upsample=up[0].numpy().copy(),
                              after_decoder_transformer=tr[0].numpy().copy())
            else:
                audio = model.mimi.decode_from_latent(quant, mimi_state)
            increment_steps(model.mimi, mimi_state, increment=16)
            pcm.append(audio.reshape(-1).clone())
            backbone = nxt if teacher is None else torch.from_numpy(teacher[step]).view(1, 1, 32)
    finally:
        F.gelu, torch.nn.init.normal_ = orig_gelu, orig_normal
        hook.remove()
    out = dict(latents=torch.stack(lat).numpy(), eos_logits=np.asarray(logits, np.float32), pcm=torch.stack(pcm).numpy())
    out.update({"stage_" + k: v for k, v in stages.items()})
    return out


CASES = [
    # name, weight seed, layer_scale, voice rows, n tokens, frames, lsd steps
    ("cfg1_lsd1", 1234, 0.01, 87, 12, 8, 1),
    ("cfg3_lsd4", 1234, 0.01, 87, 23, 4, 4),
    ("stress_ls05", 99, 0.5, 20, 7, 20, 1),  # large LayerScale: Mimi attention visible; 20 frames cross the 250-window
]


def make_encoder_golden(out_dir: Path):
    """Voice cloning from PCM: the package's own `_encode_audio` (models/tts_model.py:258-262 =
    mimi.encode_to_latent -> transpose -> F.linear(speaker_proj_weight)) on seeded encoder weights and a seeded prompt.
    The prompt length is not a multiple of the frame size, so the package's own end padding is exercised."""
    import torch.nn.functional as F
    weights = dict(synth.make_weights(1234, layer_scale=0.01))
    weights.update(synth.make_encoder_weights(4321, layer_scale=0.5))  # LayerScale large enough for attention to show
    model = build_reference_model(weights, 1, allow_missing_encoder=False)
    n_samples = 22 * 1920 - 700   # 22 frames = 352 encoder-transformer positions: crosses the 250-position window
    pcm = synth.make_pcm(n_samples, seed=3)
    blob = dict(n_samples=n_samples, pcm_seed=3, enc_seed=4321, enc_layer_scale=0.5)
    orig = F.gelu
    try:
        for kind in ("erf", "tanh"):
            if kind == "tanh":
                F.gelu = lambda x: orig(x, approximate="tanh")
            out = model._encode_audio(torch.from_numpy(pcm).view(1, 1, -1))[0].numpy()
            blob[f"{kind}_audio_prompt"] = out.astype(np.float32)
            print("encoder", kind, out.shape, "absmax", float(np.abs(out).max()))
    finally:
        F.gelu = orig
    np.savez_compressed(out_dir / "enc_pcm22.npz", **blob)
    print("wrote", out_dir / "enc_pcm22.npz", os.path.getsize(out_dir / "enc_pcm22.npz"))


def main():
    out_dir = Path(__file__).resolve().parent
    if len(sys.argv) > 1 and sys.argv[1] == "encoder":
        return make_encoder_golden(out_dir)
    cache = {}
    for name, wseed, ls, vrows, ntok, frames, lsd in CASES:
        key = (wseed, ls)
        if key not in cache:
            cache.clear()
            cache[key] = synth.make_weights(wseed, layer_scale=ls)
        weights = cache[key]
        model = build_reference_model(weights, lsd)
        prompt = synth.make_voice_prompt(vrows, seed=7)
        tokens = synth.make_tokens(ntok, seed=11)
        noise = synth.make_noise(frames, seed=5)
        blob = dict(weight_seed=wseed, layer_scale=ls, voice_rows=vrows, tokens=tokens, noise=noise, lsd_steps=lsd)
        for kind in ("erf", "tanh"):
            r = run_reference(model, prompt, tokens, noise, frames, kind)
            blob.update({f"{kind}_{k}": v for k, v in r.items()})
            print(name, kind, "latents absmax", float(np.abs(r["latents"]).max()), "pcm absmax",
                  float(np.abs(r["pcm"]).max()), "eos", r["eos_logits"][:4])
        np.savez_compressed(out_dir / f"{name}.npz", **blob)
        print("wrote", out_dir / f"{name}.npz", os.path.getsize(out_dir / f"{name}.npz"))


if __name__ == "__main__":
    main()
