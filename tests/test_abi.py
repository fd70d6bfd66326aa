"""The C-ABI library loads, exports every symbol include/ptts.h declares, and refuses to compute
without a CUDA device (no CPU fallback)."""
import re
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def built_lib():
    from pocket_tts_b200 import build
    build.build()
    from pocket_tts_b200 import _lib
    return _lib


def header_symbols(name):
    text = (ROOT / "include" / name).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ptts_[a-z0-9_]+)\s*\(", text)))


def test_exports_match_header(built_lib):
    """Product ABI (ptts.h) and test hooks (ptts_internal.h) are separate headers; the library exports every symbol of
    both and the ctypes binding lists exactly those."""
    L = built_lib.lib()
    product, internal = header_symbols("ptts.h"), header_symbols("ptts_internal.h")
    assert len(product) >= 30 and not set(product) & set(internal)
    assert not [s for s in product if "test" in s or "debug" in s or "profile" in s]
    for s in product + internal:
        assert hasattr(L, s), f"{s} declared in a header but not exported"
    assert sorted(built_lib.PRODUCT_SYMBOLS) == product
    assert sorted(built_lib.INTERNAL_SYMBOLS) == internal
    assert L.ptts_abi_version() == 2


def test_config_check_reads_the_reference_yaml_schema(built_lib, tmp_path):
    """ptts_config_check (config.rs:111): a YAML in the reference's schema with the b6369a24 dimensions passes (block lists,
    nested maps, comments), a changed dimension is named.  Needs no device."""
    from pocket_tts_b200 import synth
    from pocket_tts_b200.engine import config_check
    p = tmp_path / "m.yaml"
    p.write_text(synth.make_config_yaml())
    config_check(p)
    p.write_text(synth.make_config_yaml(**{"flow_lm.transformer.num_heads": 12}))
    with pytest.raises(built_lib.PttsError) as ei:
        config_check(p)
    assert "flow_lm.transformer.num_heads" in str(ei.value)
    p.write_text(synth.make_config_yaml(**{"mimi.seanet.ratios": [6, 5, 8]}))
    with pytest.raises(built_lib.PttsError) as ei:
        config_check(p)
    assert "ratios" in str(ei.value)
    with pytest.raises(built_lib.PttsError):
        config_check(tmp_path / "missing.yaml")


def test_sass_is_blackwell_native(built_lib):
    """tcgen05.mma / TMA / TMEM loads must be in the shipped SASS (B200_PROFILING.md evidence table)."""
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", str(built_lib.LIB_PATH)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    # mma.sync is allowed only in the two attention kernels whose per-(sequence, head) products are far below a tcgen05
    # M=128 issue -- the 16-query Mimi window attention and the prefill attention (<= 64 query rows x ~130 keys x 64 dims);
    # every GEMM-shaped op (Linear / Conv / ConvTranspose / flow head) must stay on tcgen05
    fn, legacy = "", set()
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
        elif "HMMA.16816" in line:
            legacy.add(fn)
    assert all("mimi_attn" in f or "attn_prefill_mma" in f for f in legacy), legacy


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(built_lib):
    from pocket_tts_b200.engine import test_gemm
    with pytest.raises(built_lib.PttsError) as ei:
        test_gemm(np.zeros((4, 64), np.float32), np.zeros((8, 64), np.float32))
    assert ei.value.code == -2  # PTTS_ERR_CUDA


def test_product_never_imports_oracle():
    for p in (ROOT / "pocket_tts_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p


def test_plain_c_caller(built_lib, tmp_path):
    """include/ptts.h is valid C11 and a plain C host can link the library (the boundary the Rust crate binds,
    INTEGRATION.md): struct sizes match the ctypes binding, null arguments are refused, and without a device the engine
    refuses to exist (PTTS_ERR_CUDA) instead of computing on the CPU."""
    import ctypes as C
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    exe = tmp_path / "abi_check"
    lib_dir = built_lib.LIB_PATH.parent
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "c_abi" / "abi_check.c"),
                        "-o", str(exe), f"-L{lib_dir}", "-lptts_cuda", f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    args = [str(exe)] + ([] if torch.cuda.is_available() else ["--no-gpu"])
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok header" in r.stdout and "FAIL" not in r.stdout
    # the same sizes the ctypes structures have
    assert C.sizeof(built_lib.EngineCfg) == 64 and C.sizeof(built_lib.TensorDesc) == 56 and C.sizeof(built_lib.StreamParams) == 32
    assert C.sizeof(built_lib.Segment) == 56


def test_library_is_built_from_the_sources_in_the_tree():
    """The shipped binary carries the content hash of what it was compiled from (pocket_tts_b200/build.py); loading it with
    other sources in the tree rebuilds or refuses (never a stale kernel)."""
    from pocket_tts_b200 import build
    build.build()
    assert build.STAMP.exists() and build.STAMP.read_text().strip() == build.source_hash()
    assert not build.needs_build()
