"""The persistent FlowLM step kernel (csrc/lm_step.cuh: one launch for the language-model half of a decode step) against
the golden vectors of the unmodified reference package, the CPU oracle and the per-layer launch path.

Same tolerances as tests/test_parity_gpu.py (BASELINE.json north_star): teacher-forced latents max-abs <= 1e-2, PCM SNR
>= 40 dB, EOS logit / frame bookkeeping as the reference's."""
import numpy as np
import pytest

from pocket_tts_b200 import synth

pytestmark = pytest.mark.gpu

LAT_TOL = 1e-2
SNR_MIN = 40.0


def snr(ref, x):
    return 10 * np.log10((ref ** 2).sum() / max(((ref - x) ** 2).sum(), 1e-30))


def make_engine(seed=1234, ls=0.01, step_kernel=True, **kw):
    from pocket_tts_b200.engine import Engine
    w = synth.make_weights(int(seed), layer_scale=float(ls))
    return Engine(w, max_slots=kw.pop("max_slots", 8), kv_capacity=kw.pop("kv_capacity", 512), lm_step_kernel=step_kernel, **kw), w


@pytest.mark.parametrize("case", ["cfg1_lsd1", "cfg3_lsd4", "stress_ls05"])
def test_step_kernel_golden_teacher_forced(golden_dir, case):
    from pocket_tts_b200.engine import StreamSpec
    g = np.load(golden_dir / f"{case}.npz")
    eng, _ = make_engine(g["weight_seed"], g["layer_scale"])
    eng.set_lsd_steps(int(g["lsd_steps"]))
    voice = eng.voice_from_prompt(synth.make_voice_prompt(int(g["voice_rows"]), seed=7))
    frames = g["tanh_latents"].shape[0]
    slots = eng.open_streams([voice], [StreamSpec(g["tokens"], frames, 0, 1e30, noise=g["noise"][:frames])])
    lat, pcm, logit = [], [], []
    eng.launch_count(reset=True)  # the prefill above is the per-layer path
    for f in range(frames):
        if f:
            eng.set_feedback(int(slots[0]), g["tanh_latents"][f - 1])
        p, fin, l, lg = eng.step(slots)
        lat.append(l[0]); pcm.append(p[0]); logit.append(lg[0])
        assert bool(fin[0]) == (f == frames - 1)
    launches = eng.launch_count()
    eng.close_stream(int(slots[0])); voice.close(); eng.close()
    assert np.abs(np.stack(lat) - g["tanh_latents"]).max() <= LAT_TOL
    assert snr(g["tanh_pcm"], np.stack(pcm)) >= SNR_MIN
    assert np.abs(np.array(logit) - g["tanh_eos_logits"]).max() < 2e-2
    # the language-model half really is one launch: step kernel + one flow head per LSD step + EOS bookkeeping; the codec
    # half is 28 launches (the per-layer path needs 52 + 3 per LSD step for the language-model half alone)
    assert launches / frames <= 32 + int(g["lsd_steps"]), launches / frames


def test_step_kernel_ragged_batch_mixed_voices_matches_oracle():
    """Streams of different lengths, two different voices in one batch (so the shared-prefix staging only covers part of
    the rows of a CTA and the others stream their prefix through the per-warp rings) and a voice longer than the staged
    window (128 rows): each stream against its own oracle run."""
    from oracle import ptts_oracle as O
    from pocket_tts_b200.engine import StreamSpec
    eng, wnp = make_engine(max_slots=16, kv_capacity=256)
    W = O.to_torch(wnp)
    prompts = [synth.make_voice_prompt(33, seed=21), synth.make_voice_prompt(12, seed=22), synth.make_voice_prompt(140, seed=23)]
    voices = [eng.voice_from_prompt(p) for p in prompts]
    ovs = [O.voice_state_from_prompt(W, p) for p in prompts]
    plan = [(5, 3, 0), (17, 4, 1), (1, 2, 0), (9, 4, 2), (30, 3, 1), (3, 3, 2), (40, 2, 0)]
    specs, refs, vs = [], [], []
    for i, (ntok, frames, vi) in enumerate(plan):
        tok, noise = synth.make_tokens(ntok, seed=100 + i), synth.make_noise(frames, seed=200 + i)
        specs.append(StreamSpec(tok, frames, 0, 1e30, noise=noise))
        refs.append(O.generate_segment(W, ovs[vi], tok, noise, frames, 0, float("inf")))
        vs.append(voices[vi])
    slots = eng.open_streams(vs, specs)
    active = list(range(len(specs)))
    got_lat = [[] for _ in specs]
    got_pcm = [[] for _ in specs]
    step = 0
    while active:
        for i in active:
            if step > 0:
                eng.set_feedback(int(slots[i]), refs[i]["latents"][step - 1])
        pcm, fin, lat, _ = eng.step(slots[active])
        nxt = []
        for j, i in enumerate(active):
            got_lat[i].append(lat[j]); got_pcm[i].append(pcm[j])
            if fin[j]:
                eng.close_stream(int(slots[i]))
            else:
                nxt.append(i)
        active = nxt
        step += 1
    for v in voices:
        v.close()
    eng.close()
    for i, r in enumerate(refs):
        assert len(got_lat[i]) == r["frames"]
        assert np.abs(np.stack(got_lat[i]) - r["latents"]).max() <= LAT_TOL, i
        assert snr(r["pcm"], np.stack(got_pcm[i])) >= SNR_MIN, i


def test_step_kernel_matches_per_layer_path_at_64_streams_and_is_reproducible():
    """64 streams (the benchmarked batch), free running for a few frames: the step kernel and the per-layer launches see
    the same inputs, so they may differ only by the f32 summation order of their split-K partials; two runs of the step
    kernel are bit-identical (fixed reduction orders, no atomics)."""
    from pocket_tts_b200.engine import StreamSpec
    n, frames = 64, 4
    prompt = synth.make_voice_prompt(87, seed=7)
    outs = []
    for step_kernel in (False, True, True):
        eng, _ = make_engine(step_kernel=step_kernel, max_slots=n, kv_capacity=128)
        voice = eng.voice_from_prompt(prompt)
        specs = [StreamSpec(synth.make_tokens(20 + (i % 21), seed=100 + i), frames, 0, 1e30, noise=synth.make_noise(frames, seed=300 + i)) for i in range(n)]
        slots = eng.open_streams([voice] * n, specs)
        lat, lg, pcm = [], [], []
        for f in range(frames):
            p, fin, l, g = eng.step(slots)
            lat.append(l); lg.append(g); pcm.append(p)
        outs.append((np.stack(lat), np.stack(lg), np.stack(pcm)))
        voice.close(); eng.close()
    a, b, c = outs
    assert np.array_equal(b[0], c[0]) and np.array_equal(b[1], c[1]) and np.array_equal(b[2], c[2])
    # frame 0 has no feedback: tight; later frames drift apart with the AR feedback of rounding differences
    assert np.abs(a[0][0] - b[0][0]).max() <= 1e-2 and np.abs(a[1][0] - b[1][0]).max() <= 1e-2
    assert np.abs(a[0] - b[0]).max() <= 5e-2
    assert snr(a[2][0], b[2][0]) >= SNR_MIN
