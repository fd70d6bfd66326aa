"""Bring-up probe (not a test): the persistent FlowLM step kernel against the per-layer launch path on the same
streams, then a timing of both.  python tests/lm_step_probe.py [streams] [frames]"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pocket_tts_b200 import synth  # noqa: E402
from pocket_tts_b200.engine import Engine, StreamSpec  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w = synth.make_weights(1234)
prompt = synth.make_voice_prompt(23, seed=7)
out = {}
for mega in (False, True):
    eng = Engine(w, max_slots=max(n, 2), kv_capacity=256, lm_step_kernel=mega)
    voice = eng.voice_from_prompt(prompt)
    specs = [StreamSpec(synth.make_tokens(5 + (i % 7), seed=100 + i), frames, 0, 1e30, noise=synth.make_noise(frames, seed=200 + i)) for i in range(n)]
    slots = eng.open_streams([voice] * n, specs)
    lats, logits, pcms, xs, hs = [], [], [], [], []
    for f in range(frames):
        pcm, fin, lat, lg = eng.step(slots)
        lats.append(lat); logits.append(lg); pcms.append(pcm)
        xs.append(np.stack([eng.debug_read("flowlm.x", r) for r in range(min(n, 3))]))
        hs.append(np.stack([eng.debug_read("flowlm.h", r) for r in range(min(n, 3))]))
    out[mega] = (np.stack(lats), np.stack(logits), np.stack(pcms), np.stack(xs), np.stack(hs))
    # timing: device-resident steps
    for s in slots:
        eng.close_stream(int(s))
    specs = [StreamSpec(synth.make_tokens(40, seed=100 + i), 100, 0, 1e30, temp=0.7, seed=i) for i in range(n)]
    slots = eng.open_streams([voice] * n, specs)
    for _ in range(10):
        eng.step_device(slots)
    eng.sync()
    t0 = time.perf_counter()
    for _ in range(60):
        eng.step_device(slots)
    eng.sync()
    dt = (time.perf_counter() - t0) / 60
    ms = eng.step_timed(slots)
    print(f"mega={mega}: {dt*1e6:.1f} us/step pipelined; sequential stages (lm, flow, front+mimi, seanet) ms = {ms[:4].round(4).tolist()} total {ms[5]:.4f}", flush=True)
    eng.sync()
    voice.close(); eng.close()
a, b = out[False], out[True]
for name, i in (("latents", 0), ("eos_logit", 1), ("pcm", 2), ("flowlm.x", 3), ("flowlm.h", 4)):
    d = np.abs(a[i] - b[i])
    print(f"{name}: max|legacy - mega| per frame {d.reshape(d.shape[0], -1).max(1).round(6).tolist()}  (ref max {np.abs(a[i]).max():.3f})")
