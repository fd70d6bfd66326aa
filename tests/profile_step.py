"""One mid-utterance decode step of the bench workload inside a cudaProfilerStart/Stop range, for
`ncu --profile-from-start off` (launch list and full-set captures; see profiles/).
python tests/profile_step.py [streams] [warm steps] [profiled steps] [step_kernel: 0|1]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pocket_tts_b200 import synth
from pocket_tts_b200.engine import Engine, StreamSpec

streams = int(sys.argv[1]) if len(sys.argv) > 1 else 64
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 60
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
step_kernel = bool(int(sys.argv[4])) if len(sys.argv) > 4 else None
eng = Engine(synth.make_weights(1234), max_slots=streams, kv_capacity=40 + 125 + 3, lm_step_kernel=step_kernel)
voice = eng.voice_from_prompt(synth.make_voice_prompt(87, seed=7))
specs = [StreamSpec(synth.make_tokens(40, seed=1000 + i), 125, 3, 1e30, temp=0.7, seed=i) for i in range(streams)]
slots = eng.open_streams([voice] * streams, specs)
for _ in range(warm):
    eng.step_device(slots)
eng.sync()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(steps):
    eng.step_device(slots)
eng.sync()
torch.cuda.cudart().cudaProfilerStop()
ms = eng.step_timed(slots)
print("stage ms [flowlm, flow head, mimi transformer, seanet, control, total]:", ms[:6].round(4).tolist())
print("launches per step:", eng.launch_count() // (warm + steps + 1))
