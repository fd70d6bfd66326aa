"""Bring-up probe: phase-by-phase check of the persistent FlowLM step kernel against numpy on its own buffers
(every phase of layer 0 except attention, whose inputs live in the KV cache)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    from pocket_tts_b200 import synth
    from pocket_tts_b200.engine import Engine, StreamSpec
    n = int(sys.argv[2])
    w = synth.make_weights(1234)
    eng = Engine(w, max_slots=max(n, 2), kv_capacity=256, lm_step_kernel=True)
    voice = eng.voice_from_prompt(synth.make_voice_prompt(23, seed=7))
    specs = [StreamSpec(synth.make_tokens(5 + (i % 7), seed=100 + i), 4, 0, 1e30, noise=synth.make_noise(4, seed=200 + i)) for i in range(n)]
    slots = eng.open_streams([voice] * n, specs)
    eng.step(slots)
    out = {"x": np.stack([eng.debug_read("flowlm.x", r) for r in range(n)])}
    for name in ("lm.hA", "lm.attnA", "lm.ffnA"):
        out[name] = eng.debug_read(name, 0, cap=1 << 20)
    out["ws"] = np.concatenate([eng.debug_read("lm.ws", i, cap=1 << 20) for i in range(2)])
    np.savez(sys.argv[3], **out)
    sys.exit(0)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5


def run(stop):
    f = f"/tmp/lmchk_{stop}.npz"
    env = dict(os.environ, PTTS_LM_STOP=str(stop))
    subprocess.run([sys.executable, __file__, "child", str(n), f], env=env, check=True, capture_output=True)
    return np.load(f)


def image(raw, kblocks):
    """f16 image [kb][64 rows][64 k], SWIZZLE_128B -> [64 rows, kb*64] float32"""
    b = raw.view(np.uint8)[: kblocks * 8192].reshape(kblocks, 64, 8, 16)
    out = np.empty((kblocks, 64, 8, 16), np.uint8)
    for r in range(64):
        for c in range(8):
            out[:, r, c] = b[:, r, c ^ (r & 7)]
    return out.reshape(kblocks, 64, 64 * 2).view(np.float16).astype(np.float32).transpose(1, 0, 2).reshape(64, kblocks * 64)


from pocket_tts_b200 import synth  # noqa: E402
W = synth.make_weights(1234)
P = "flow_lm.transformer.layers.0."
f16 = lambda a: a.astype(np.float16).astype(np.float32)


def ln(x, w, b, eps=1e-5):
    m = x.mean(-1, keepdims=True)
    v = ((x - m) ** 2).mean(-1, keepdims=True)
    return (x - m) / np.sqrt(v + eps) * w + b


def rep(name, got, want):
    d = np.abs(got - want)
    print(f"{name:34s} max err {d.max():.3e}  (ref max {np.abs(want).max():.3f})", flush=True)


d1 = run(1)
x0 = d1["x"][:n]
h0 = image(d1["lm.hA"], 16)[:n]
rep("ph0 LN1(x) image", h0, f16(ln(x0, W[P + "norm1.weight"], W[P + "norm1.bias"])))
d2 = run(2)
S = 6
ws = d2["ws"][: S * 64 * 3072].reshape(S, 64, 3072).sum(0)[:n]
rep("ph1 in_proj partial sum", ws, h0 @ f16(W[P + "self_attn.in_proj.weight"]).T)
d3 = run(3)
a = image(d3["lm.attnA"], 16)[:n]
print("ph2 attention image: max", np.abs(a).max(), "nan", np.isnan(a).any(), flush=True)
d4 = run(4)
ws = d4["ws"][: 16 * 64 * 1024].reshape(16, 64, 1024).sum(0)[:n]
rep("ph3 out_proj partial sum", ws, a @ f16(W[P + "self_attn.out_proj.weight"]).T)
d5 = run(5)
x1 = d5["x"][:n]
rep("ph4 x + out_proj", x1, x0 + ws)
h1 = image(d5["lm.hA"], 16)[:n]
rep("ph4 LN2 image", h1, f16(ln(x1, W[P + "norm2.weight"], W[P + "norm2.bias"])))
d6 = run(6)
ws1 = d6["ws"][: 4 * 64 * 4096].reshape(4, 64, 4096).sum(0)[:n]
rep("ph5 linear1 partial sum", ws1, h1 @ f16(W[P + "linear1.weight"]).T)
d7 = run(7)
g = image(d7["lm.ffnA"], 64)[:n]
gelu = lambda v: 0.5 * v * (1 + np.tanh(0.7978845608028654 * v * (1 + 0.044715 * v * v)))
rep("ph6 gelu image", g, f16(gelu(ws1)))
d8 = run(8)
ws2 = d8["ws"][: 16 * 64 * 1024].reshape(16, 64, 1024).sum(0)[:n]
rep("ph7 linear2 partial sum", ws2, g @ f16(W[P + "linear2.weight"]).T)
d9 = run(9)
rep("ph8 x + linear2", d9["x"][:n], x1 + ws2)

