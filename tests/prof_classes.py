import sys
sys.path.insert(0, '.')
from pocket_tts_b200 import synth
from pocket_tts_b200.engine import Engine, StreamSpec
eng = Engine(synth.make_weights(1234), max_slots=64, kv_capacity=170)
voice = eng.voice_from_prompt(synth.make_voice_prompt(87, seed=7))
slots = eng.open_streams([voice] * 64, [StreamSpec(synth.make_tokens(40, seed=i), 125, 3, 1e30, temp=0.7, seed=i) for i in range(64)])
for _ in range(60): eng.step_device(slots)
eng.sync()
for _ in range(3):
    eng.profile(True); eng.step_device(slots); eng.profile(False)
rep = eng.profile_report(); ovh = eng.profile_overhead_us()
tot = 0
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]['ms']):
    us = 1000*v['ms']/v['launches']-ovh; tot += us*v['launches']/3
    print(f"{k:24s} {v['launches']/3:4.0f} x {us:7.1f} us = {us*v['launches']/3:7.1f}")
print('sum', tot, 'overhead per pair', ovh)
