"""Bring-up probe (not a test): durations of path A (FlowLM + flow head) and path B (Mimi + SEANet) inside the pipelined step
(python tests/overlap_probe.py [streams]; run with PTTS_DIAG_TIMES=1, PTTS_DIAG_SKIP=1|2 for one path alone; the library prints the event stamps of the last step at every sync)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pocket_tts_b200 import synth
from pocket_tts_b200.engine import Engine, StreamSpec
streams = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eng = Engine(synth.make_weights(1234), max_slots=streams, kv_capacity=40 + 125 + 3)
voice = eng.voice_from_prompt(synth.make_voice_prompt(87, seed=7))
specs = [StreamSpec(synth.make_tokens(40, seed=1000 + i), 125, 3, 1e30, temp=0.7, seed=i) for i in range(streams)]
slots = eng.open_streams([voice] * streams, specs)
for rep in range(3):
    eng.sync()
    t = time.perf_counter()
    for _ in range(30):
        eng.step_device(slots)
    eng.sync()
    print(f"30 steps: {(time.perf_counter() - t) / 30 * 1e6:.1f} us/step", flush=True)
