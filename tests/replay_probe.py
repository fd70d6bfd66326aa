"""Bring-up probe (not a test): graph-replayed time of the four FlowLM decode Linear shapes (ptts_profile_gemm_replay) and
the step time, under whatever PTTS_* switches the environment carries.  python tests/replay_probe.py [streams] [int8: 0|1]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pocket_tts_b200 import synth
from pocket_tts_b200.engine import Engine, StreamSpec
streams = int(sys.argv[1]) if len(sys.argv) > 1 else 64
int8 = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
eng = Engine(synth.make_weights(1234), max_slots=streams, kv_capacity=40 + 125 + 3, int8_weights=int8)
voice = eng.voice_from_prompt(synth.make_voice_prompt(87, seed=7))
specs = [StreamSpec(synth.make_tokens(40, seed=1000 + i), 125, 3, 1e30, temp=0.7, seed=i) for i in range(streams)]
slots = eng.open_streams([voice] * streams, specs)
for rep in range(2):
    eng.sync()
    t = time.perf_counter()
    for _ in range(40):
        eng.step_device(slots)
    eng.sync()
    us = (time.perf_counter() - t) / 40 * 1e6
rp = eng.gemm_replay(streams, 20)
tot_b = sum(v["bytes"] for v in rp.values()); tot_us = sum(v["us"] for v in rp.values())
print(f"step {us:.1f} us | " + " ".join(f"{k} {v['us']:.2f}" for k, v in rp.items()) + f" | pooled {tot_b / tot_us / 1e3:.0f} GB/s = {tot_b / tot_us / 1e3 / 6533.5:.3f}", flush=True)
