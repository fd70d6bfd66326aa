"""Bring-up probe: first frame / stage at which a codec group of 2 frames differs from single-frame passes."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pocket_tts_b200 import synth
from pocket_tts_b200.engine import Engine, StreamSpec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 24
w = synth.make_weights(1234)
taps = ["mimi.after_decoder_transformer", "seanet.convtr2", "seanet.convtr5", "seanet.convtr8", "pcm"]
out = {}
for run, cg in enumerate((1, 1, 2, 4, 2)):
    eng = Engine(w, max_slots=4 * n + 2, kv_capacity=64, codec_group=cg)
    voice = eng.voice_from_prompt(synth.make_voice_prompt(11, seed=7))
    specs = [StreamSpec(synth.make_tokens(5, seed=300 + i), frames, 0, 1e30, noise=synth.make_noise(frames, seed=i)) for i in range(n)]
    slots = eng.open_streams([voice] * n, specs)
    out[run] = {}
    for k in range(frames):
        eng.step_device(slots)
        if k % cg == cg - 1:
            eng.sync()
            out[run][k - (cg - 1)] = {t: eng.debug_read(t, (n - 1) * cg) for t in taps}   # [row][frames of the group][...]
    voice.close(); eng.close()
for run in (1, 2, 3, 4):
    print("run", run)
    for k in sorted(out[run]):
        line = f"frame {k:2d}:"
        for t in taps:
            d = np.abs(out[0][k][t] - out[run][k][t])
            bad = np.flatnonzero(d > 0)
            line += f"  {t.split('.')[-1][:10]}: {bad.size}/{d.size} first {bad[0] if bad.size else -1} max {d.max():.1e}"
        if any(np.abs(out[0][k][t] - out[run][k][t]).max() > 0 for t in taps):
            print(line)
