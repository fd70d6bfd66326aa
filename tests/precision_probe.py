"""Design probe (not a test): predict the error of a bf16-operand engine by rounding GEMM
activation operands / KV rows to bf16 inside the CPU oracle."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import ptts_oracle as O
from pocket_tts_b200 import synth

def bf(x): return x.bfloat16().float()
def split2(x):  # hi+lo bf16 pair ~ 16 mantissa bits
    hi = x.bfloat16().float(); return hi + (x - hi).bfloat16().float()
def f16(x): return x.half().float()

def snr(ref, x):
    return 10*np.log10((ref**2).sum()/max(((ref-x)**2).sum(),1e-30))

def run(frames, a, kv, W, voice, tokens, noise, teacher=None):
    O._a, O._kv = a, kv
    try:
        return O.generate_segment(W, voice, tokens, noise, frames, 0, float("inf"), teacher_latents=teacher)
    finally:
        O._a = O._kv = (lambda x: x)

if __name__ == "__main__":
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    W = O.to_torch(synth.make_weights(1234))
    prompt = synth.make_voice_prompt(87); tokens = synth.make_tokens(40, 3); noise = synth.make_noise(frames, 5)
    ident = lambda x: x
    O._a = O._kv = ident
    voice = O.voice_state_from_prompt(W, prompt)
    ref = run(frames, ident, ident, W, voice, tokens, noise)
    for name, a, kv in [("bf16 act + bf16 kv", bf, bf), ("bf16 act + f32 kv", bf, ident), ("f32 act + bf16 kv", ident, bf),
                        ("split-bf16 act + bf16 kv", split2, bf), ("fp16 act + fp16 kv", f16, f16), ("split2 act+kv", split2, split2)]:
        O._a, O._kv = a, kv
        v = O.voice_state_from_prompt(W, prompt)
        free = run(frames, a, kv, W, v, tokens, noise)
        tf = run(frames, a, kv, W, v, tokens, noise, teacher=ref["latents"])
        e_free = np.abs(free["latents"]-ref["latents"]).max(axis=1)
        e_tf = np.abs(tf["latents"]-ref["latents"]).max(axis=1)
        print(f"{name:28s} teacher-forced lat maxabs {e_tf.max():.2e} (mean {e_tf.mean():.2e}) pcm SNR {snr(ref['pcm'], tf['pcm']):.1f} dB | "
              f"free lat maxabs {e_free.max():.2e} first8 {e_free[:8].max():.2e} pcm SNR {snr(ref['pcm'], free['pcm']):.1f} dB "
              f"eos maxerr {np.abs(free['eos_logits']-ref['eos_logits']).max():.2e}")
