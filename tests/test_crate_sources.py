"""crates/pocket-tts-cuda cannot be compiled in this image (no cargo / rustc): at least keep its FFI honest -- src/ffi.rs must
declare exactly the product ABI of include/ptts.h, with the argument counts of the C prototypes, and its constants must match."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def c_protos():
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "ptts.h").read_text(), flags=re.S)
    out = {}
    for m in re.finditer(r"\b(ptts_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def rust_protos():
    text = (ROOT / "crates" / "pocket-tts-cuda" / "src" / "ffi.rs").read_text()
    text = re.sub(r"//.*", "", text)
    block = text[text.index('extern "C" {'):]
    out = {}
    for m in re.finditer(r"pub fn (ptts_[a-z0-9_]+)\s*\(([^;]*?)\)\s*(?:->\s*[^;]+)?;", block, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    return out


def test_ffi_declares_exactly_the_product_abi():
    c, r = c_protos(), rust_protos()
    assert len(c) >= 30
    assert sorted(c) == sorted(r), (sorted(set(c) - set(r)), sorted(set(r) - set(c)))
    for name, n in c.items():
        assert r[name] == n, (name, n, r[name])


def test_ffi_constants_match_the_header():
    h = (ROOT / "include" / "ptts.h").read_text()
    rs = (ROOT / "crates" / "pocket-tts-cuda" / "src" / "ffi.rs").read_text()
    for name in ("PTTS_ABI_VERSION", "PTTS_STEP_PCM", "PTTS_STEP_AHEAD", "PTTS_STEP_PCM_I16", "PTTS_FRAME_OVERRUN", "PTTS_SEG_TEXT", "PTTS_SEG_PAUSE"):
        cv = int(re.search(rf"#define {name} (-?\d+)", h).group(1))
        rv = int(re.search(rf"pub const {name}: \w+ = (-?\d+);", rs).group(1))
        assert cv == rv, name
    for name, val in (("PTTS_ERR_INVALID", -1), ("PTTS_ERR_CUDA", -2), ("PTTS_ERR_CAPACITY", -3), ("PTTS_ERR_STATE", -4)):
        assert re.search(rf"{name} = {val}", h) and re.search(rf"pub const {name}: i32 = {val};", rs)


def test_facade_keeps_the_reference_surface():
    src = (ROOT / "crates" / "pocket-tts-cuda" / "src" / "lib.rs").read_text()
    for method in ("pub fn load(", "pub fn load_with_params(", "pub fn get_voice_state<", "pub fn get_voice_state_from_prompt_file<",
                   "pub fn get_voice_state_from_prompt_tensor(", "pub fn get_voice_state_from_tensor(", "pub fn generate(",
                   "pub fn generate_stream<", "pub fn generate_stream_long<", "pub struct SegmentIter", "impl Iterator for SegmentIter"):
        assert method in src, method
    for field in ("pub temp: f32", "pub lsd_decode_steps: usize", "pub eos_threshold: f32", "pub noise_clamp: Option<f32>",
                  "pub sample_rate: usize", "pub dim: usize", "pub ldim: usize"):
        assert field in src, field
