"""Deterministic random-init weights of the b6369a24 architecture.

The pretrained checkpoint is gated and absent (reference:
crates/pocket-tts/config/b6369a24.yaml:3), so tests and the benchmark use
random-init tensors with the checkpoint's exact safetensors key names and
PyTorch layouts (Linear [out,in], Conv1d [out,in,k], ConvTranspose1d
[in,out,k]; built by the reference at crates/pocket-tts/src/tts_model.rs:296-409).

Every value is rounded to a bf16-representable number. The published
checkpoint was itself saved from a bf16-cast model
(python-reference/pocket_tts/models/tts_model.py:143-145), so bf16 weight
storage in the engine is lossless for these tensors, and the CPU oracle and the
CUDA engine see bit-identical parameters.
"""
from __future__ import annotations

import numpy as np

# Model constants (crates/pocket-tts/config/b6369a24.yaml:6-56)
D_MODEL = 1024
N_HEADS = 16
N_LAYERS = 6
D_FFN = 4096
LDIM = 32
FLOW_DIM = 512
FLOW_DEPTH = 6
MIMI_DIM = 512
MIMI_HEADS = 8
MIMI_LAYERS = 2
MIMI_FFN = 2048
MIMI_CONTEXT = 250
N_BINS = 4000
SAMPLE_RATE = 24000
FRAME_SAMPLES = 1920
UPSAMPLE_STRIDE = 16


def bf16_round(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even f32 -> bf16 -> f32 (numpy only)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(a.shape)


def weight_shapes(with_encoder: bool = False) -> dict[str, tuple[int, ...]]:
    """Key -> shape for every tensor on the generation hot path."""
    s: dict[str, tuple[int, ...]] = {}
    s["flow_lm.input_linear.weight"] = (D_MODEL, LDIM)
    for i in range(N_LAYERS):
        p = f"flow_lm.transformer.layers.{i}."
        s[p + "self_attn.in_proj.weight"] = (3 * D_MODEL, D_MODEL)
        s[p + "self_attn.out_proj.weight"] = (D_MODEL, D_MODEL)
        for n in ("norm1", "norm2"):
            s[p + n + ".weight"] = (D_MODEL,)
            s[p + n + ".bias"] = (D_MODEL,)
        s[p + "linear1.weight"] = (D_FFN, D_MODEL)
        s[p + "linear2.weight"] = (D_MODEL, D_FFN)
    s["flow_lm.out_norm.weight"] = (D_MODEL,)
    s["flow_lm.out_norm.bias"] = (D_MODEL,)
    s["flow_lm.out_eos.weight"] = (1, D_MODEL)
    s["flow_lm.out_eos.bias"] = (1,)
    s["flow_lm.bos_emb"] = (LDIM,)
    s["flow_lm.emb_mean"] = (LDIM,)
    s["flow_lm.emb_std"] = (LDIM,)
    f = "flow_lm.flow_net."
    s[f + "cond_embed.weight"] = (FLOW_DIM, D_MODEL)
    s[f + "cond_embed.bias"] = (FLOW_DIM,)
    s[f + "input_proj.weight"] = (FLOW_DIM, LDIM)
    s[f + "input_proj.bias"] = (FLOW_DIM,)
    for t in range(2):
        q = f + f"time_embed.{t}.mlp."
        s[q + "0.weight"] = (FLOW_DIM, 256)
        s[q + "0.bias"] = (FLOW_DIM,)
        s[q + "2.weight"] = (FLOW_DIM, FLOW_DIM)
        s[q + "2.bias"] = (FLOW_DIM,)
        s[q + "3.alpha"] = (FLOW_DIM,)
    for i in range(FLOW_DEPTH):
        q = f + f"res_blocks.{i}."
        s[q + "in_ln.weight"] = (FLOW_DIM,)
        s[q + "in_ln.bias"] = (FLOW_DIM,)
        s[q + "mlp.0.weight"] = (FLOW_DIM, FLOW_DIM)
        s[q + "mlp.0.bias"] = (FLOW_DIM,)
        s[q + "mlp.2.weight"] = (FLOW_DIM, FLOW_DIM)
        s[q + "mlp.2.bias"] = (FLOW_DIM,)
        s[q + "adaLN_modulation.1.weight"] = (3 * FLOW_DIM, FLOW_DIM)
        s[q + "adaLN_modulation.1.bias"] = (3 * FLOW_DIM,)
    s[f + "final_layer.linear.weight"] = (LDIM, FLOW_DIM)
    s[f + "final_layer.linear.bias"] = (LDIM,)
    s[f + "final_layer.adaLN_modulation.1.weight"] = (2 * FLOW_DIM, FLOW_DIM)
    s[f + "final_layer.adaLN_modulation.1.bias"] = (2 * FLOW_DIM,)
    s["flow_lm.conditioner.embed.weight"] = (N_BINS + 1, D_MODEL)
    s["flow_lm.speaker_proj_weight"] = (D_MODEL, MIMI_DIM)
    s["mimi.quantizer.output_proj.weight"] = (MIMI_DIM, LDIM, 1)
    s["mimi.upsample.convtr.convtr.weight"] = (MIMI_DIM, 1, 2 * UPSAMPLE_STRIDE)
    for i in range(MIMI_LAYERS):
        p = f"mimi.decoder_transformer.transformer.layers.{i}."
        s[p + "self_attn.in_proj.weight"] = (3 * MIMI_DIM, MIMI_DIM)
        s[p + "self_attn.out_proj.weight"] = (MIMI_DIM, MIMI_DIM)
        for n in ("norm1", "norm2"):
            s[p + n + ".weight"] = (MIMI_DIM,)
            s[p + n + ".bias"] = (MIMI_DIM,)
        s[p + "linear1.weight"] = (MIMI_FFN, MIMI_DIM)
        s[p + "linear2.weight"] = (MIMI_DIM, MIMI_FFN)
        s[p + "layer_scale_1.scale"] = (MIMI_DIM,)
        s[p + "layer_scale_2.scale"] = (MIMI_DIM,)
    d = "mimi.decoder.model."
    conv = {
        "0.conv": (512, 512, 7),
        "2.convtr": (512, 256, 12),
        "3.block.1.conv": (128, 256, 3),
        "3.block.3.conv": (256, 128, 1),
        "5.convtr": (256, 128, 10),
        "6.block.1.conv": (64, 128, 3),
        "6.block.3.conv": (128, 64, 1),
        "8.convtr": (128, 64, 8),
        "9.block.1.conv": (32, 64, 3),
        "9.block.3.conv": (64, 32, 1),
        "11.conv": (1, 64, 3),
    }
    for k, shp in conv.items():
        s[d + k + ".weight"] = shp
        cout = shp[1] if k.endswith("convtr") else shp[0]
        s[d + k + ".bias"] = (cout,)
    return s


def _fan_in(name: str, shape: tuple[int, ...]) -> int:
    if len(shape) == 2:
        return shape[1]
    if len(shape) == 3:
        if "convtr" in name:
            # ConvTranspose1d [in, out/g, k]: each output sample sums in * (k/stride) taps;
            # every decoder convtr has k = 2*stride (seanet.rs:346-356, conv.rs:326-336).
            return shape[0] * 2 if shape[1] != 1 else 2
        return shape[1] * shape[2]
    return 1


def make_weights(seed: int = 1234, layer_scale: float = 0.01, gain: float = 1.0) -> dict[str, np.ndarray]:
    """Random-init, bf16-representable f32 tensors keyed like the checkpoint.

    Linear/Conv weights ~ N(0, gain/fan_in) keep activations O(1) so the bf16
    tolerances are meaningful; norm scales/biases, LayerScale, emb_std/emb_mean
    are perturbed away from their identity defaults so a kernel that drops one
    of them fails parity.
    """
    rng = np.random.default_rng(seed)
    out: dict[str, np.ndarray] = {}
    for name, shape in weight_shapes().items():
        n = int(np.prod(shape))
        g = rng.standard_normal(n, dtype=np.float32).reshape(shape)
        if name.endswith("layer_scale_1.scale") or name.endswith("layer_scale_2.scale"):
            w = layer_scale * (1.0 + 0.25 * g)
        elif name.endswith("emb_std"):
            w = 1.0 + 0.25 * np.abs(g)
        elif name.endswith("emb_mean"):
            w = 0.1 * g
        elif name.endswith("bos_emb"):
            w = g
        elif name.endswith(".alpha"):
            w = 1.0 + 0.1 * g
        elif len(shape) == 1 and name.endswith(".weight"):
            w = 1.0 + 0.1 * g  # LayerNorm scale
        elif len(shape) == 1:
            w = 0.05 * g  # biases (LayerNorm, Linear, Conv)
        elif name.endswith("conditioner.embed.weight"):
            w = g  # LUT rows ~ N(0,1) like nn.Embedding
        else:
            w = g * np.float32(np.sqrt(gain / _fan_in(name, shape)))
        out[name] = bf16_round(w.astype(np.float32))
    return out


def encoder_weight_shapes() -> dict[str, tuple[int, ...]]:
    """Mimi encoder side (voice cloning from PCM): SEANetEncoder, encoder transformer, downsample.
    Key names and layouts of the checkpoint (reference build sites: models/seanet.rs:148-247, models/mimi.rs:60-98)."""
    s: dict[str, tuple[int, ...]] = {}
    e = "mimi.encoder.model."
    conv = {
        "0.conv": (64, 1, 7),
        "1.block.1.conv": (32, 64, 3), "1.block.3.conv": (64, 32, 1),
        "3.conv": (128, 64, 8),
        "4.block.1.conv": (64, 128, 3), "4.block.3.conv": (128, 64, 1),
        "6.conv": (256, 128, 10),
        "7.block.1.conv": (128, 256, 3), "7.block.3.conv": (256, 128, 1),
        "9.conv": (512, 256, 12),
        "11.conv": (512, 512, 3),
    }
    for k, shp in conv.items():
        s[e + k + ".weight"] = shp
        s[e + k + ".bias"] = (shp[0],)
    for i in range(MIMI_LAYERS):
        p = f"mimi.encoder_transformer.transformer.layers.{i}."
        s[p + "self_attn.in_proj.weight"] = (3 * MIMI_DIM, MIMI_DIM)
        s[p + "self_attn.out_proj.weight"] = (MIMI_DIM, MIMI_DIM)
        for n in ("norm1", "norm2"):
            s[p + n + ".weight"] = (MIMI_DIM,)
            s[p + n + ".bias"] = (MIMI_DIM,)
        s[p + "linear1.weight"] = (MIMI_FFN, MIMI_DIM)
        s[p + "linear2.weight"] = (MIMI_DIM, MIMI_FFN)
        s[p + "layer_scale_1.scale"] = (MIMI_DIM,)
        s[p + "layer_scale_2.scale"] = (MIMI_DIM,)
    s["mimi.downsample.conv.conv.weight"] = (MIMI_DIM, MIMI_DIM, 2 * UPSAMPLE_STRIDE)
    return s


def make_encoder_weights(seed: int = 4321, layer_scale: float = 0.01) -> dict[str, np.ndarray]:
    """Seeded encoder-side tensors from their OWN generator, so adding them never changes make_weights()' tensors
    (the committed goldens depend on those)."""
    rng = np.random.default_rng(seed)
    out: dict[str, np.ndarray] = {}
    for name, shape in encoder_weight_shapes().items():
        g = rng.standard_normal(int(np.prod(shape)), dtype=np.float32).reshape(shape)
        if name.endswith("layer_scale_1.scale") or name.endswith("layer_scale_2.scale"):
            w = layer_scale * (1.0 + 0.25 * g)
        elif len(shape) == 1 and name.endswith(".weight"):
            w = 1.0 + 0.1 * g
        elif len(shape) == 1:
            w = 0.05 * g
        else:
            w = g * np.float32(np.sqrt(1.0 / _fan_in(name, shape)))
        out[name] = bf16_round(w.astype(np.float32))
    return out


def make_pcm(n_samples: int, seed: int = 3) -> np.ndarray:
    """Synthetic 24 kHz mono prompt: a few decaying partials plus noise, peak ~0.5."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / 24000.0
    x = sum(a * np.sin(2 * np.pi * f * t + ph) for a, f, ph in zip((0.3, 0.15, 0.1), (180.0, 410.0, 1230.0), rng.uniform(0, 6.28, 3)))
    x = x * (0.6 + 0.4 * np.sin(2 * np.pi * 3.0 * t)) + 0.02 * rng.standard_normal(n_samples)
    return x.astype(np.float32)


def make_voice_prompt(n_rows: int = 87, seed: int = 7) -> np.ndarray:
    """Synthetic `audio_prompt` [T, 1024] with the magnitude of the reference's
    assets/ref_voice_conditioning.safetensors (max-abs ~0.9, 87 rows)."""
    rng = np.random.default_rng(seed)
    return (0.2 * rng.standard_normal((n_rows, D_MODEL), dtype=np.float32)).astype(np.float32)


def make_tokens(n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(0, N_BINS, size=n, dtype=np.int64).astype(np.int32)


def make_noise(frames: int, seed: int, temp: float = 0.7) -> np.ndarray:
    """Injected x_0 per frame: N(0, temp) like flow_lm.rs:39-48 (std = sqrt(temp))."""
    rng = np.random.default_rng(seed)
    return (np.sqrt(np.float32(temp)) * rng.standard_normal((frames, LDIM), dtype=np.float32)).astype(np.float32)


def make_config_yaml(**override) -> str:
    """A model YAML in the schema of the reference's config.rs:6-108 (the layout of config/b6369a24.yaml) filled with the
    b6369a24 dimensions; `override` replaces leaf values by dotted path (tests of ptts_config_check)."""
    cfg = {
        "flow_lm": {"dtype": "float32", "flow": {"depth": 6, "dim": 512},
                    "transformer": {"d_model": 1024, "hidden_scale": 4, "max_period": 10000, "num_heads": 16, "num_layers": 6},
                    "lookup_table": {"dim": 1024, "n_bins": 4000, "tokenizer": "sentencepiece"}},
        "mimi": {"dtype": "float32", "sample_rate": 24000, "channels": 1, "frame_rate": 12.5,
                 "seanet": {"dimension": 512, "channels": 1, "n_filters": 64, "n_residual_layers": 1, "ratios": [6, 5, 4],
                            "kernel_size": 7, "residual_kernel_size": 3, "last_kernel_size": 3, "dilation_base": 2,
                            "pad_mode": "constant", "compress": 2},
                 "transformer": {"d_model": 512, "num_heads": 8, "num_layers": 2, "layer_scale": 0.01, "context": 250,
                                 "dim_feedforward": 2048, "input_dimension": 512, "output_dimensions": [512]},
                 "quantizer": {"dimension": 32, "output_dimension": 512}},
    }
    for path, val in override.items():
        node = cfg
        keys = path.split(".")
        for k in keys[:-1]:
            node = node[k]
        node[keys[-1]] = val

    def emit(node, indent):
        out = []
        for k, v in node.items():
            if isinstance(v, dict):
                out.append(" " * indent + f"{k}:")
                out += emit(v, indent + 2)
            elif isinstance(v, list):
                out.append(" " * indent + f"{k}:")
                out += [" " * indent + f"- {x}" for x in v]
            else:
                out.append(" " * indent + f"{k}: {v}")
        return out
    return "# generated by pocket_tts_b200.synth.make_config_yaml\n" + "\n".join(emit(cfg, 0)) + "\n"
