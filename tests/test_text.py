"""Host text front end and output formats (N4): the reference's in-tree known-answer tests replayed,
and the Unigram tokenizer restatement checked against the `tokenizers` library the reference links."""
import io
import struct
import wave
from pathlib import Path

import numpy as np
import pytest

from pocket_tts_b200 import audio as A
from pocket_tts_b200 import text as T

REF_TOKENIZER = Path("/root/reference/crates/pocket-tts/assets/tokenizer.json")

SENTENCES = [
    "        Hello, world!", "Hello world.", "The quick brown fox jumps over the lazy dog.",
    "It costs 1,000 dollars... maybe more; who knows?", "naïve café — déjà vu", "日本語のテキスト", "emoji 🙂 inside",
    "  leading and   multiple   spaces ", "x", "", "UPPER lower MiXeD 12345 !@#$%", "supercalifragilisticexpialidocious",
]


# ------------------------------------------------------------------------------------------------ pause.rs KATs
def test_pause_kats():
    # crates/pocket-tts/src/pause.rs:192-262
    p = T.parse_explicit_pauses("Hello [pause:500ms] world")
    assert len(p) == 1 and p[0].duration_ms == 500 and p[0].original == "[pause:500ms]"
    p = T.parse_explicit_pauses("Test [pause:1s] and [pause:1.5s]")
    assert [x.duration_ms for x in p] == [1000, 1500]
    p = T.parse_natural_pauses("Hello... world")
    assert len(p) == 1 and p[0].duration_ms == T.ELLIPSIS_MS
    p = T.parse_natural_pauses("Hello, world")
    assert len(p) == 1 and p[0].duration_ms == T.COMMA_MS
    assert T.parse_natural_pauses("That costs 1,000 dollars") == []
    assert T.strip_pause_markers("Hello [pause:500ms] world [pause:1s] done") == "Hello   world   done"
    clean, pauses = T.parse_text_with_pauses("Hello... [pause:500ms] world, done")
    assert clean == "Hello...   world, done" and len(pauses) == 3
    assert T.silence_samples(500, 24000) == 12000 and T.silence_samples(1000, 24000) == 24000


def test_long_form_segments():
    # tts_model.rs:1079-1107: text and pauses interleaved in order; the explicit marker's position is where its
    # replacement space sits in the clean text
    segs = T.long_form_segments("Hello [pause:500ms] world")
    assert segs == [("text", "Hello "), ("pause", 500), ("text", " world")]
    segs = T.long_form_segments("Wait... then go, now [pause:1s]")
    assert [s for s in segs if s[0] == "pause"] == [("pause", 500), ("pause", 200), ("pause", 1000)]
    assert [s[1].strip() for s in segs if s[0] == "text"] == ["Wait", "then go", "now"]
    assert T.long_form_segments("[pause:0ms] only text") == [("text", "  only text")]  # 0 ms markers are dropped (pause.rs:163)
    # byte offsets, not char offsets: multi-byte text in front of a marker
    segs = T.long_form_segments("café [pause:250ms] déjà")
    assert segs == [("text", "café "), ("pause", 250), ("text", " déjà")]
    # integration_tests.rs:265-325: a 500 ms pause adds exactly 12000 samples at 24 kHz
    assert sum(T.silence_samples(v) for k, v in T.long_form_segments("Hello [pause:500ms] world") if k == "pause") == 12000


# ------------------------------------------------------------------------------------------------ tokenizer
def _synthetic_vocab():
    rng = np.random.default_rng(0)
    pieces = ["<unk>", "<s>", "</s>", "<pad>"] + [f"<0x{b:02X}>" for b in range(256)]
    scores = [0.0] * len(pieces)
    base = ["▁", "▁the", "▁a", "the", "he", "ll", "o", "▁hello", "▁world", "wor", "ld", "▁", ",", ".", "!", "▁H", "ello",
            "▁qu", "ick", "▁brown", "▁fox", "s", "▁over", "▁l", "azy", "▁dog", "é", "caf", "▁▁", "▁▁▁▁", "1", "0", "00"]
    base += list("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ")
    seen = set(pieces)
    for p in base:
        if p not in seen:
            seen.add(p)
            pieces.append(p)
            scores.append(float(-rng.uniform(2.0, 12.0)))
    return list(zip(pieces, scores))


def _hf(vocab, prepend, bos):
    tk = pytest.importorskip("tokenizers")
    from tokenizers import Tokenizer, models, pre_tokenizers, processors
    t = Tokenizer(models.Unigram(vocab, unk_id=0, byte_fallback=True))
    t.pre_tokenizer = pre_tokenizers.Metaspace(replacement="▁", prepend_scheme=prepend, split=False)
    if bos:
        t.post_processor = processors.TemplateProcessing(single="<s> $A", special_tokens=[("<s>", 1)])
    return t


@pytest.mark.parametrize("prepend,bos", [("always", False), ("never", True)])
def test_unigram_matches_tokenizers_library_synthetic(prepend, bos):
    """native build: Metaspace(Always), no BOS (conditioners/text.rs:57-80); WASM json: never + `<s>` (assets/tokenizer.json)."""
    vocab = _synthetic_vocab()
    mine = T.UnigramTokenizer(vocab, 0, True, prepend, (1,) if bos else ())
    ref = _hf(vocab, prepend, bos)
    for s in SENTENCES:
        if s == "":
            continue  # the library returns only specials for empty input; covered below
        assert mine.encode(s) == ref.encode(s).ids, s
    assert mine.encode("") == ([1] if bos else [])


@pytest.mark.skipif(not REF_TOKENIZER.exists(), reason="reference assets not present on this box")
def test_unigram_matches_reference_tokenizer_json():
    tk = pytest.importorskip("tokenizers")
    ref = tk.Tokenizer.from_file(str(REF_TOKENIZER))
    mine = T.UnigramTokenizer.from_tokenizer_json(REF_TOKENIZER)
    assert mine.vocab_size == 4000  # n_bins (config/b6369a24.yaml), checked by LUTConditioner::new (text.rs:38-45)
    rng = np.random.default_rng(1)
    words = "the of and to in is that it was for on are as with his they at be this from have or by one had not but what all".split()
    extra = [" ".join(rng.choice(words, size=int(rng.integers(1, 40)))) + rng.choice(list(".!?")) for _ in range(200)]
    for s in SENTENCES + extra:
        if s:
            assert mine.encode(s) == ref.encode(s).ids, s
    ids = mine.encode(T.prepare_text_prompt("Hello, world!"))
    assert ids[0] == 1 and len(ids) == 12  # cfg1 = 12 ids incl. <s>
    # native-style construction over the same vocabulary: '▁' always prepended, no BOS
    native = T.UnigramTokenizer(mine.vocab, 0, True, "always", ())
    assert len(native.encode("Hello world.")) == len(mine.encode(" Hello world.")) - 1


def _sp_piece(piece, score, ty=None):
    def varint(v):
        out = bytearray()
        while v >= 0x80:
            out.append((v & 0x7F) | 0x80)
            v >>= 7
        out.append(v)
        return bytes(out)
    msg = b"\x0a" + varint(len(piece.encode())) + piece.encode() + b"\x15" + struct.pack("<f", score)
    if ty is not None:
        msg += b"\x18" + varint(ty)
    return b"\x0a" + varint(len(msg)) + msg


def test_sentencepiece_protobuf_kats():
    # conditioners/text.rs:388-421
    assert T.read_varint(bytes([0xAC, 0x02, 0x01]), 0) == (300, 2)
    assert T.read_varint(bytes([0xAC, 0x02, 0x01]), 2) == (1, 3)
    vocab, unk = T.parse_sentencepiece_vocab(_sp_piece("<unk>", -1.0, 2) + _sp_piece("hello", -2.5, 1))
    assert unk == 0 and [p for p, _ in vocab] == ["<unk>", "hello"]
    assert abs(vocab[0][1] + 1.0) < 1e-6 and abs(vocab[1][1] + 2.5) < 1e-6
    with pytest.raises(ValueError, match="No vocabulary found"):
        T.parse_sentencepiece_vocab(b"")
    # unknown trailing fields of ModelProto (trainer spec etc.) are skipped
    blob = _sp_piece("<unk>", 0.0, 2) + _sp_piece("▁a", -1.0, 1) + b"\x12\x03abc" + b"\x18\x05"
    tok = T.UnigramTokenizer.from_sentencepiece_model(blob)
    assert tok.vocab_size == 2 and tok.encode("a") == [1]


# ------------------------------------------------------------------------------------------------ sentence packing
def test_split_into_best_sentences():
    # tts_model.rs:603-684 with a one-token-per-word counter
    count = lambda s: len(s.split())
    # sentences are trimmed (the 8 leading spaces of a short prompt come back when the segment re-prepares its chunk)
    assert T.split_into_best_sentences("hello world", count) == ["Hello world."]
    assert T.prepare_text_prompt("Hello world.") == "        Hello world."
    text = "One two three. Four five six! Seven eight? Nine; ten: eleven"
    assert T.split_into_best_sentences(text, count) == ["One two three. Four five six! Seven eight? Nine; ten: eleven."]
    long = " ".join(f"w{i}" for i in range(30)) + ". " + " ".join(f"v{i}" for i in range(30)) + "."
    chunks = T.split_into_best_sentences(long, count)
    assert len(chunks) == 2 and chunks[0].startswith("W0") and chunks[1].startswith("v0")  # 30 + 30 > 50
    huge = " ".join(f"w{i}" for i in range(100))
    chunks = T.split_into_best_sentences("Short one. " + huge, count)
    assert chunks[0] == "Short one." and [len(c.split()) for c in chunks[1:]] == [35, 35, 30]  # word batches of 35
    # a batch that is still over the limit is halved once (two tokens per word)
    chunks = T.split_into_best_sentences(huge, lambda s: 2 * len(s.split()))
    assert [len(c.split()) for c in chunks] == [17, 18, 17, 18, 15, 15]
    # a failing counter counts as 50 tokens (`unwrap_or`)
    def boom(s):
        raise RuntimeError
    assert T.split_into_best_sentences("A b. C d.", boom) == ["A b.", "C d."]
    assert T.split_into_best_sentences("", count) == ["."]


def test_model_facade_text_path_without_gpu():
    """generate_stream's host half: chunks -> (tokens, max_gen_len, frames_after_eos) without touching the engine."""
    from pocket_tts_b200.tts_model import TTSModel
    m = TTSModel.__new__(TTSModel)  # no engine: only the host text path is exercised
    m.tokenizer = T.UnigramTokenizer(_synthetic_vocab(), 0, True, "always", ())
    m.temp, m.eos_threshold, m.sample_rate = 0.7, -4.0, 24000
    req = m.long_form_request("Hello world, the quick brown fox [pause:300ms] jumps over the lazy dog.")
    kinds = [k for k, _ in req]
    assert kinds == ["text", "pause", "text", "pause", "text"]
    assert [v for k, v in req if k == "pause"] == [200, 300]
    spec = req[0][1]
    assert spec.max_gen_len == (2 + 2) * 13 and spec.frames_after_eos == 5 and spec.tokens.dtype == np.int32


# ------------------------------------------------------------------------------------------------ output formats
def test_pcm_i16_and_wav():
    # audio.rs:124-144: clamp, * 32767, truncate toward zero
    x = np.array([[0.0, 1.0, -1.0, 2.0, -2.0, 0.5, -0.5, 1e-5, np.nan]], np.float32)
    assert A.pcm_i16(x)[:, 0].tolist() == [0, 32767, -32767, 32767, -32767, 16383, -16383, 0, 0]
    b = A.pcm_i16_le_bytes(x)
    assert len(b) == 18 and b[2:4] == struct.pack("<h", 32767)
    stereo = np.stack([np.linspace(-1, 1, 7), np.linspace(1, -1, 7)]).astype(np.float32)
    inter = np.frombuffer(A.pcm_i16_le_bytes(stereo), "<i2").reshape(7, 2)  # interleaved by sample
    assert inter[0].tolist() == [-32767, 32767]
    with pytest.raises(ValueError, match="shape"):
        A.pcm_i16_le_bytes(np.zeros(5, np.float32))
    rng = np.random.default_rng(0)
    a = (rng.standard_normal((1, 4800)) * 0.3).astype(np.float32)
    buf = io.BytesIO(A.wav_bytes(a, 24000))
    with wave.open(buf) as w:  # stdlib parser as the independent check of the container
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 24000, 4800)
        np.testing.assert_array_equal(np.frombuffer(w.readframes(4800), "<i2"), A.pcm_i16(a)[:, 0])
    assert A.stream_chunk_bytes(a[None, :, :1920]) == A.pcm_i16_le_bytes(a[:, :1920])
    np.testing.assert_allclose(np.abs(A.normalize_peak(a)).max(), 1.0, rtol=1e-6)
    assert not A.normalize_peak(np.zeros((1, 4), np.float32)).any()


def test_noise_clamp_rejection_sampling():
    """models/flow_lm.rs:39-65: truncated normal by rejection; std = sqrt(temp); temp = 0 gives exact zeros."""
    from pocket_tts_b200.tts_model import sample_clamped_noise
    z = sample_clamped_noise(200, 0.7, 1.0, seed=3)
    assert z.shape == (200, 32) and z.dtype == np.float32
    assert np.abs(z).max() <= 1.0
    # N(0, 0.7) truncated at |v| <= 1 has std 0.5239 (scipy.stats.truncnorm)
    assert abs(z.std() - 0.5239) < 0.01 and abs(z.mean()) < 0.02
    np.testing.assert_array_equal(z, sample_clamped_noise(200, 0.7, 1.0, seed=3))
    assert not np.array_equal(z, sample_clamped_noise(200, 0.7, 1.0, seed=4))
    wide = sample_clamped_noise(200, 0.7, 10.0, seed=3)
    assert abs(wide.std() - np.sqrt(0.7)) < 0.01          # an inactive clamp leaves N(0, temp)
    assert not sample_clamped_noise(5, 0.0, 1.0).any()
    with pytest.raises(ValueError):
        sample_clamped_noise(5, 0.7, 0.0)
