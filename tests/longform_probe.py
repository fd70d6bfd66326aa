"""Probe (not a test): BASELINE configs[4] per-GPU slice -- `requests` concurrent long-form requests of 6 chunks x 125 frames
(60 s of speech) with a 300 ms `[pause:..]` between chunks, continuous batching on one engine, PCM of every frame copied
back to the host.  python tests/longform_probe.py [requests] [python|ahead|native|native16]"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pocket_tts_b200 import synth
from pocket_tts_b200.engine import Engine, StreamSpec
from pocket_tts_b200.tts_model import BatchScheduler

requests_n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mode = sys.argv[2] if len(sys.argv) > 2 else "native16"
chunks, frames, tokens, pause_ms = 6, 125, 40, 300
eng = Engine(synth.make_weights(1234), max_slots=requests_n, kv_capacity=tokens + frames + 3)
voice = eng.voice_from_prompt(synth.make_voice_prompt(87, seed=7))
reqs = []
for r in range(requests_n):
    segs = []
    for c in range(chunks):
        if c:
            segs.append(("pause", pause_ms))
        segs.append(("text", StreamSpec(synth.make_tokens(tokens, seed=r * 16 + c), frames, 3, 1e30, temp=0.7, seed=r * 16 + c)))
    reqs.append(segs)
sched = BatchScheduler(eng, voice)
t0 = time.perf_counter()
out = sched.run(reqs, ahead=mode == "ahead", native=mode.startswith("native"), i16=mode == "native16")
dt = time.perf_counter() - t0
want = chunks * frames * 1920 + (chunks - 1) * pause_ms * 24
assert all(o.shape == (want,) for o in out) and all(np.isfinite(o.astype(np.float32)).all() for o in out[:4])
audio_s = requests_n * want / 24000.0
print(f"[{mode}] {requests_n} requests x {want / 24000:.1f} s = {audio_s:.0f} audio-s in {dt:.2f} s wall -> {audio_s / dt:.0f} audio-s per wall-s "
      f"(speech only: {requests_n * chunks * frames * 0.08 / dt:.0f}), launches {eng.launch_count()}")
