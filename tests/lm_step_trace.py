"""Bring-up probe: per-phase %globaltimer stamps of the persistent FlowLM step kernel (PTTS_LM_TRACE=1).
python tests/lm_step_trace.py [streams] [warm frames]"""
import os
import sys
from pathlib import Path

os.environ["PTTS_LM_TRACE"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pocket_tts_b200 import synth  # noqa: E402
from pocket_tts_b200.engine import Engine, StreamSpec  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 60
eng = Engine(synth.make_weights(1234), max_slots=n, kv_capacity=256, lm_step_kernel=True)
voice = eng.voice_from_prompt(synth.make_voice_prompt(87, seed=7))
specs = [StreamSpec(synth.make_tokens(40, seed=100 + i), 200, 0, 1e30, temp=0.7, seed=i) for i in range(n)]
slots = eng.open_streams([voice] * n, specs)
for _ in range(warm):
    eng.step_device(slots)
eng.sync()
ms = eng.step_timed(slots)
print("sequential stages ms (lm, flow, front+mimi, seanet):", ms[:4].round(4).tolist(), "total", round(float(ms[5]), 4), flush=True)
eng.sync()
