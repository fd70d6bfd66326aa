"""Isolated kernels through the C ABI vs plain torch f32 references (operands rounded to f16 on both
sides, so the only difference is accumulation order)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def f16r(a):
    return np.asarray(a, np.float32).astype(np.float16).astype(np.float32)


def ref_gemm(a, w, bias=None, act=0):
    d = torch.from_numpy(f16r(a)).double() @ torch.from_numpy(f16r(w)).double().T
    if bias is not None:
        d = d + torch.from_numpy(np.asarray(bias, np.float32)).double()
    if act == 1:
        d = F.gelu(d, approximate="tanh")
    elif act == 2:
        d = F.silu(d)
    return d.float().numpy()


GEMM_CASES = [
    # rows, feats, k, mode (0 auto, 1 act-as-M, 2 weight-as-M), split_k, act
    (64, 3072, 1024, 0, 1, 0),      # FlowLM in_proj at B=64 (swap-AB)
    (1, 1024, 1024, 0, 1, 0),       # B=1
    (64, 1024, 4096, 0, 8, 0),      # linear2 with cluster split-K
    (64, 4096, 1024, 0, 1, 1),      # linear1 + tanh-GELU
    (200, 512, 512, 0, 1, 2),       # flow head width, SiLU, ragged rows
    (1024, 1536, 512, 0, 1, 0),     # Mimi in_proj at B=64 (activation-as-M)
    (300, 640, 512, 1, 1, 0),       # ragged rows and features, activation-as-M
    (1000, 72, 64, 1, 1, 0),        # tiny K, odd feature count
    (64, 32, 512, 0, 1, 0),         # flow final linear (32 features)
    (30720, 256, 256, 1, 1, 0),     # SEANet convtr8 shape: persistent kernel, two feature tiles
    (20000, 64, 64, 1, 1, 2),       # SEANet res9b shape: persistent kernel, ragged last tile, single K block
]


@pytest.mark.parametrize("rows,feats,k,mode,split,act", GEMM_CASES)
@pytest.mark.parametrize("simt", [0, 1])
def test_gemm(rows, feats, k, mode, split, act, simt):
    from pocket_tts_b200.engine import test_gemm as run
    rng = np.random.default_rng(rows * 7 + feats)
    a = rng.standard_normal((rows, k), dtype=np.float32)
    w = rng.standard_normal((feats, k), dtype=np.float32) / np.sqrt(k)
    bias = rng.standard_normal(feats, dtype=np.float32)
    got = run(a, w, bias, mode=mode, split_k=split, act=act, use_simt=simt)
    want = ref_gemm(a, w, bias, act)
    err = np.abs(got - want).max()
    assert err < 2e-3, f"max abs err {err}"


INT8_CASES = [
    # rows, feats, k, split_k
    (64, 3072, 1024, 2),    # FlowLM in_proj at B=64: resident decode mode, 2-CTA cluster
    (64, 4096, 1024, 1),    # linear1: 16 k-blocks through the stage ring
    (64, 1024, 4096, 4),    # linear2: ring + cluster split-K
    (1, 1024, 1024, 8),     # B=1, 8-CTA cluster, two k-blocks per CTA
    (200, 640, 512, 1),     # ragged rows / features (prefill shape)
    (256, 128, 64, 1),      # one k-block
]


@pytest.mark.parametrize("rows,feats,k,split", INT8_CASES)
def test_gemm_int8_storage(rows, feats, k, split):
    """BASELINE configs[3]: one-byte weight codes streamed from HBM and expanded to f16 in shared memory must be
    bit-identical to streaming an f16 copy of the same codes, and both equal A . (codes * scale)^T of the reference's
    per-tensor scheme (crates/pocket-tts/src/quantize.rs:65-94)."""
    from pocket_tts_b200.engine import test_gemm_int8 as run
    rng = np.random.default_rng(rows + feats + k)
    a = rng.standard_normal((rows, k), dtype=np.float32)
    w = (rng.standard_normal((feats, k), dtype=np.float32) / np.sqrt(k)).astype(np.float32)  # f32: the codes are defined on f32 division
    w[0, :8] = [w.max(), -np.abs(w).max(), 0.0, 1e-9, -1e-9, w.min(), 0.5 * w.max(), -0.5 * w.max()]  # extremes of the code range
    got, scale = run(a, w, split_k=split, storage=1)
    f16_codes, scale2 = run(a, w, split_k=split, storage=0)
    assert scale == scale2 == np.float32(np.abs(w).max() / np.float32(127.0))
    np.testing.assert_array_equal(got, f16_codes)
    codes = np.clip(np.rint(w / np.float32(scale)), -127, 127)
    want = (torch.from_numpy(f16r(a)).double() @ torch.from_numpy(codes).double().T * scale).float().numpy()
    err = np.abs(got - want).max()
    assert err < 2e-3, f"max abs err {err}"


CONV_CASES = [(3, 16, 512, 512, 7), (2, 96, 256, 128, 3), (2, 480, 128, 64, 3), (1, 1920, 64, 64, 3), (9, 16, 512, 64, 7),
              (12, 1920, 64, 64, 3), (40, 480, 128, 64, 3)]  # the last two run the persistent kernel


@pytest.mark.parametrize("n,t,cin,cout,k", CONV_CASES)
def test_streaming_conv1d(n, t, cin, cout, k):
    from pocket_tts_b200.engine import test_conv1d as run
    rng = np.random.default_rng(t + k)
    x = rng.standard_normal((n, t, cin), dtype=np.float32)
    prev = rng.standard_normal((n, k - 1, cin), dtype=np.float32)
    w = rng.standard_normal((cout, cin, k), dtype=np.float32) / np.sqrt(cin * k)
    b = rng.standard_normal(cout, dtype=np.float32)
    got = run(x, prev, w, b)
    xp = torch.from_numpy(f16r(np.concatenate([prev, x], axis=1))).permute(0, 2, 1)
    want = F.conv1d(xp.double(), torch.from_numpy(f16r(w)).double(), torch.from_numpy(b).double()).permute(0, 2, 1).float().numpy()
    err = np.abs(got - want).max()
    assert err < 2e-3, f"max abs err {err}"


CONVTR_CASES = [(3, 16, 512, 256, 6), (2, 96, 256, 128, 5), (2, 480, 128, 64, 4), (9, 16, 512, 256, 6), (40, 480, 128, 64, 4)]


@pytest.mark.parametrize("n,t,cin,cout,s", CONVTR_CASES)
def test_streaming_convtr1d(n, t, cin, cout, s):
    """y over [prev_row | x] must equal the reference's overlap-add with carried `partial`
    (modules/conv.rs:219-267): partial = tail of convtr(prev_row) minus bias."""
    from pocket_tts_b200.engine import test_convtr1d as run
    rng = np.random.default_rng(t + s)
    x = rng.standard_normal((n, t, cin), dtype=np.float32)
    prev = rng.standard_normal((n, cin), dtype=np.float32)
    w = rng.standard_normal((cin, cout, 2 * s), dtype=np.float32) / np.sqrt(2 * cin)
    b = rng.standard_normal(cout, dtype=np.float32)
    got = run(x, prev, w, b, s)
    xin = torch.from_numpy(f16r(np.concatenate([prev[:, None], x], axis=1))).permute(0, 2, 1).double()
    full = F.conv_transpose1d(xin, torch.from_numpy(f16r(w)).double(), torch.from_numpy(b).double(), stride=s)
    want = full[:, :, s:s + t * s].permute(0, 2, 1).float().numpy()  # drop prev_row's own first s samples and the new tail
    err = np.abs(got - want).max()
    assert err < 2e-3, f"max abs err {err}"
