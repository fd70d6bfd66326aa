"""Small-batch Linear (csrc/gemv.cuh): 1-4 activation rows, the shapes of a single-utterance decode step (BASELINE
configs[0]; reference Linear sites modules/attention.rs:129,280, models/transformer.rs:85, modules/mlp.rs:322-368), through
the C ABI against a float64 product of the same f16-rounded operands, against the tensor-core GEMM on the same inputs, and
with one-byte weight codes (reference quantize.rs:65-94).  Every test asserts that the GEMV family was really selected."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def f16r(a):
    return np.asarray(a, np.float32).astype(np.float16).astype(np.float32)


def ref(a, w, bias=None, act=0):
    d = torch.from_numpy(f16r(a)).double() @ torch.from_numpy(f16r(w)).double().T
    if bias is not None:
        d = d + torch.from_numpy(np.asarray(bias, np.float32)).double()
    if act == 1:
        d = torch.nn.functional.gelu(d, approximate="tanh")
    elif act == 2:
        d = torch.nn.functional.silu(d)
    return d.float().numpy()


# rows, feats, k, act: every (row bucket, chunks-per-lane) instantiation; feature counts that leave warps without work, need
# a second batch per warp (10240 = flow adaLN), or are not a multiple of anything (1000, 72)
CASES = [
    (1, 3072, 1024, 0), (1, 1024, 1024, 0), (1, 4096, 1024, 1), (1, 1024, 4096, 0), (1, 10240, 512, 0), (1, 512, 1024, 2),
    (2, 3072, 1024, 0), (2, 1024, 4096, 0), (2, 10240, 512, 0), (2, 1000, 1024, 1),
    (3, 4096, 1024, 1), (3, 1024, 4096, 0), (3, 512, 512, 2),
    (4, 3072, 1024, 0), (4, 1024, 4096, 0), (4, 10240, 512, 0), (4, 72, 512, 0), (4, 32, 512, 0),
]


@pytest.mark.parametrize("rows,feats,k,act", CASES)
def test_gemv_matches_reference_and_tensor_core_path(rows, feats, k, act):
    from pocket_tts_b200.engine import test_gemm as run, gemv_launches
    rng = np.random.default_rng(rows * 131 + feats + k)
    a = rng.standard_normal((rows, k), dtype=np.float32)
    w = rng.standard_normal((feats, k), dtype=np.float32) / np.sqrt(k)
    bias = rng.standard_normal(feats, dtype=np.float32)
    before = gemv_launches()
    got = run(a, w, bias, mode=0, act=act)
    assert gemv_launches() == before + 1, "the small-batch GEMV was not selected"
    tc = run(a, w, bias, mode=2, act=act)          # the same product on tcgen05 (weights on MMA-M)
    assert gemv_launches() == before + 1
    want = ref(a, w, bias, act)
    assert np.abs(got - want).max() < 2e-3
    assert np.abs(got - tc).max() < 1e-4           # f32 accumulation on both sides: only the order differs


@pytest.mark.parametrize("rows,feats,k", [(1, 3072, 1024), (1, 1024, 4096), (1, 10240, 512), (2, 4096, 1024), (3, 1024, 4096),
                                          (4, 3072, 1024), (4, 1024, 4096), (4, 512, 512), (2, 1000, 1024)])
def test_gemv_int8_codes(rows, feats, k):
    """Byte codes expanded in registers == the tensor-core path streaming the same codes (both exact products of f16 x
    integer, f32 accumulate) up to summation order, and == A . (codes * scale)^T of the reference's per-tensor scheme."""
    from pocket_tts_b200.engine import test_gemm_int8 as run, gemv_launches
    rng = np.random.default_rng(rows + feats + k)
    a = rng.standard_normal((rows, k), dtype=np.float32)
    w = (rng.standard_normal((feats, k), dtype=np.float32) / np.sqrt(k)).astype(np.float32)
    w[0, :8] = [w.max(), -np.abs(w).max(), 0.0, 1e-9, -1e-9, w.min(), 0.5 * w.max(), -0.5 * w.max()]
    before = gemv_launches()
    got, scale = run(a, w, storage=2)
    assert gemv_launches() == before + 1, "the small-batch GEMV was not selected"
    tc, scale2 = run(a, w, storage=1)
    assert scale == scale2
    codes = np.clip(np.rint(w / np.float32(scale)), -127, 127)
    want = (torch.from_numpy(f16r(a)).double() @ torch.from_numpy(codes).double().T * scale).float().numpy()
    assert np.abs(got - want).max() < 2e-3
    assert np.abs(got - tc).max() < 1e-4


def test_gemv_not_selected_beyond_four_rows():
    from pocket_tts_b200.engine import test_gemm as run, gemv_launches
    rng = np.random.default_rng(5)
    a = rng.standard_normal((5, 1024), dtype=np.float32)
    w = rng.standard_normal((256, 1024), dtype=np.float32) / 32
    before = gemv_launches()
    got = run(a, w, None)
    assert gemv_launches() == before
    assert np.abs(got - ref(a, w)).max() < 2e-3


@pytest.mark.parametrize("n,int8", [(1, False), (3, False), (2, True)])
def test_small_batch_step_runs_on_gemv_and_matches_the_tensor_core_step(n, int8):
    """1-3 utterances through the engine with the GEMV family on (default) and off (ptts_engine_cfg.reserved[1] = 1):
    teacher-forced latents, EOS logits and PCM of both agree to accumulation-order noise."""
    from pocket_tts_b200 import synth
    from pocket_tts_b200.engine import Engine, StreamSpec, gemv_launches
    weights = synth.make_weights(3)
    voice_rows = synth.make_voice_prompt(12, seed=4)
    frames = 6
    rng = np.random.default_rng(9)
    noise = (rng.standard_normal((n, frames, 32)) * np.sqrt(0.7)).astype(np.float32)
    feed = (rng.standard_normal((n, frames, 32)) * 0.5).astype(np.float32)
    tokens = [np.arange(5 + i, 17 + 2 * i, dtype=np.int32) for i in range(n)]
    outs = []
    for off in (False, True):
        eng = Engine(weights, max_slots=4, kv_capacity=256, gemv_off=off, int8_weights=int8)
        voice = eng.voice_from_prompt(voice_rows)
        slots = eng.open_streams([voice] * n, [StreamSpec(tokens[i], frames, 0, 1e30, noise=noise[i]) for i in range(n)])
        before = gemv_launches()
        lat, pcm, logit = [], [], []
        for f in range(frames):
            if f > 0:
                for i in range(n):
                    eng.set_feedback(int(slots[i]), feed[i, f - 1])
            p, fin, l, lg = eng.step(slots)
            lat.append(l.copy()); pcm.append(p.copy()); logit.append(lg.copy())
        used = gemv_launches() - before
        assert (used > 0) == (not off), f"GEMV launches {used} with the family {'off' if off else 'on'}"
        outs.append((np.stack(lat), np.stack(pcm), np.stack(logit)))
        eng.close_streams(slots)
        voice.close()
        eng.close()
    (l0, p0, g0), (l1, p1, g1) = outs
    assert np.abs(l0 - l1).max() < 1e-2   # f16 roundings of the activations flip on accumulation-order noise; the parity bar itself
    assert np.abs(g0 - g1).max() < 1e-2
    assert np.abs(p0 - p1).max() < 1e-2


@pytest.mark.parametrize("n,int8", [(1, False), (4, False), (2, True)])
def test_fused_layernorm_prologue_is_bit_identical_to_the_layernorm_launches(n, int8, monkeypatch):
    """The LayerNorm in front of in_proj / linear1 runs inside the GEMV's prologue (same arithmetic per row as
    ln_rows_kernel, rounded to f16 once): latents, logits and PCM must not change by a bit against PTTS_GEMV_LN=0."""
    from pocket_tts_b200 import synth
    from pocket_tts_b200.engine import Engine, StreamSpec
    weights = synth.make_weights(11)
    voice_rows = synth.make_voice_prompt(9, seed=2)
    frames = 5
    rng = np.random.default_rng(3)
    noise = (rng.standard_normal((n, frames, 32)) * np.sqrt(0.7)).astype(np.float32)
    tokens = [np.arange(3 + i, 10 + 3 * i, dtype=np.int32) for i in range(n)]
    outs = []
    for ln_fused in ("1", "0"):
        monkeypatch.setenv("PTTS_GEMV_LN", ln_fused)
        eng = Engine(weights, max_slots=4, kv_capacity=256, int8_weights=int8)
        voice = eng.voice_from_prompt(voice_rows)
        slots = eng.open_streams([voice] * n, [StreamSpec(tokens[i], frames, 0, 1e30, noise=noise[i]) for i in range(n)])
        eng.launch_count(reset=True)
        res = [eng.step(slots) for _ in range(frames)]      # free-running: any difference would compound
        launches = eng.launch_count()
        outs.append((np.stack([r[2] for r in res]), np.stack([r[0] for r in res]), np.stack([r[3] for r in res]), launches))
        eng.close_streams(slots)
        voice.close()
        eng.close()
    (l0, p0, g0, n0), (l1, p1, g1, n1) = outs
    np.testing.assert_array_equal(l0, l1)
    np.testing.assert_array_equal(g0, g1)
    np.testing.assert_array_equal(p0, p1)
    assert n1 - n0 == 11 * frames, (n0, n1)     # the 11 FlowLM LayerNorm launches of a step are gone


@pytest.mark.skipif(not __import__("os").environ.get("PTTS_TEST_FLOW_SMALL"),
                    reason="flow_small.cuh is opt-in (PTTS_FLOW_SMALL=1) and so is its test (PTTS_TEST_FLOW_SMALL=1)")
@pytest.mark.parametrize("n,lsd,int8", [(1, 1, False), (2, 4, False), (3, 1, False), (4, 2, False), (1, 4, True)])
def test_small_batch_flow_head_matches_the_fused_cluster_kernel(n, lsd, int8, monkeypatch):
    """flow_head_small_kernel (1-4 rows: one cluster of 8 CTAs, GEMV warps, operands replicated through distributed shared
    memory; csrc/flow_small.cuh) against flow_head_kernel (tcgen05, PTTS_FLOW_SMALL=0) on the same conditioning: the same
    arithmetic up to summation order, for one and several Euler steps (reference flow_lm.rs:7-22, mlp.rs:135-171,275-383)."""
    from pocket_tts_b200 import synth
    from pocket_tts_b200.engine import Engine, StreamSpec
    weights = synth.make_weights(17)
    voice_rows = synth.make_voice_prompt(10, seed=8)
    frames = 4
    rng = np.random.default_rng(12)
    noise = (rng.standard_normal((n, frames, 32)) * np.sqrt(0.7)).astype(np.float32)
    feed = (rng.standard_normal((n, frames, 32)) * 0.5).astype(np.float32)
    tokens = [np.arange(4 + i, 12 + 2 * i, dtype=np.int32) for i in range(n)]
    outs = []
    for small in ("1", "0"):
        monkeypatch.setenv("PTTS_FLOW_SMALL", small)
        eng = Engine(weights, max_slots=4, kv_capacity=256, int8_weights=int8)
        eng.set_lsd_steps(lsd)
        voice = eng.voice_from_prompt(voice_rows)
        slots = eng.open_streams([voice] * n, [StreamSpec(tokens[i], frames, 0, 1e30, noise=noise[i]) for i in range(n)])
        lat, pcm = [], []
        for f in range(frames):
            if f > 0:
                for i in range(n):
                    eng.set_feedback(int(slots[i]), feed[i, f - 1])
            p, fin, l, lg = eng.step(slots)
            lat.append(l.copy()); pcm.append(p.copy())
        outs.append((np.stack(lat), np.stack(pcm)))
        eng.close_streams(slots)
        voice.close()
        eng.close()
    (l0, p0), (l1, p1) = outs
    assert np.isfinite(l0).all()
    assert np.abs(l0 - l1).max() < 2e-3, np.abs(l0 - l1).max()
    snr = 10 * np.log10((p1.astype(np.float64) ** 2).sum() / max(((p1.astype(np.float64) - p0) ** 2).sum(), 1e-30))
    assert snr >= 50.0, snr
