"""Bring-up probe (not a test): per-kernel-class time of one batched open (text prefill) at the bench shape.
python tests/prefill_probe.py [streams] [tokens]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pocket_tts_b200 import synth
from pocket_tts_b200.engine import Engine, StreamSpec
streams = int(sys.argv[1]) if len(sys.argv) > 1 else 64
tokens = int(sys.argv[2]) if len(sys.argv) > 2 else 40
eng = Engine(synth.make_weights(1234), max_slots=streams, kv_capacity=tokens + 125 + 3)
voice = eng.voice_from_prompt(synth.make_voice_prompt(87, seed=7))
specs = [StreamSpec(synth.make_tokens(tokens, seed=1000 + i), 125, 3, 1e30, temp=0.7, seed=i) for i in range(streams)]
for rep in range(3):
    eng.sync(); t = time.perf_counter()
    slots = eng.open_streams([voice] * streams, specs)
    eng.sync(); dt = time.perf_counter() - t
    eng.close_streams(slots)
print(f"open + prefill of {streams} x {tokens} tokens: {dt * 1e3:.3f} ms wall")
eng.profile(True)
slots = eng.open_streams([voice] * streams, specs)
eng.sync()
eng.profile(False)
rep = eng.profile_report()
tot = sum(v["ms"] for v in rep.values())
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:28s} launches {v['launches']:3d}  {1000 * v['ms'] / v['launches']:8.1f} us/launch  {1000 * v['ms']:8.1f} us total")
print(f"  sum {tot * 1e3:.1f} us (event-bracketed, PDL off)")
