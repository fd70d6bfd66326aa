"""Host text front end of the generation path: everything the reference does to a string before
token ids reach the FlowLM prefill.  Pure host code, no device work.

  pauses      crates/pocket-tts/src/pause.rs:52-185      explicit `[pause:Xms|Xs]` + natural (`...`, `,`) pauses
  segments    tts_model.rs:1074-1127                      text / silence interleaving of generate_stream_long
  prepare     tts_model.rs:1194-1237                      prompt normalisation, frames_after_eos
  sentences   tts_model.rs:603-684                        split_into_best_sentences (<= 50 tokens per chunk)
  tokenizer   conditioners/text.rs:57-80,329-348          Unigram + Metaspace, from the sentencepiece `.model`
                                                          protobuf (native) or a `tokenizer.json` (WASM build)

The reference tokenises with the third-party `tokenizers` crate (Cargo.lock: tokenizers 0.22); its Unigram Viterbi
search (strict `>` relaxation, unknown pieces fused, byte fallback) is restated here from its published algorithm and
checked against the `tokenizers` Python package in tests/test_text.py.
"""
from __future__ import annotations

import json
import re
import struct
from dataclasses import dataclass
from pathlib import Path

# ------------------------------------------------------------------------------------------------ pauses (pause.rs)
ELLIPSIS_MS, COMMA_MS, PERIOD_MS, SEMICOLON_MS = 500, 200, 400, 300  # pause.rs:22-31 (only the first two are used)
_EXPLICIT = re.compile(r"\[pause:(\d+(?:\.\d+)?)(ms|s)\]")  # pause.rs:34-37
_ELLIPSIS = re.compile(r"\.{3,}")                            # pause.rs:39


@dataclass(frozen=True)
class PauseMarker:
    original: str
    duration_ms: int
    position: int  # BYTE offset (the reference slices UTF-8 `str`s)


def _duration_ms(value: str, unit: str) -> int:
    v = float(value)
    return int(v) if unit == "ms" else int(v * 1000.0)  # `as u32` truncates (pause.rs:60-64)


def _byte_pos(text: str, char_pos: int) -> int:
    return len(text[:char_pos].encode("utf-8"))


def parse_explicit_pauses(text: str) -> list[PauseMarker]:
    """pause.rs:52-74"""
    return [PauseMarker(m.group(0), _duration_ms(m.group(1), m.group(2)), _byte_pos(text, m.start()))
            for m in _EXPLICIT.finditer(text)]


def parse_natural_pauses(text: str) -> list[PauseMarker]:
    """pause.rs:77-113: ellipses, and commas that are not between two digits."""
    out = [PauseMarker(m.group(0), ELLIPSIS_MS, _byte_pos(text, m.start())) for m in _ELLIPSIS.finditer(text)]
    for i, c in enumerate(text):
        if c != ",":
            continue
        prev_digit = i > 0 and text[i - 1].isascii() and text[i - 1].isdigit()
        next_digit = i + 1 < len(text) and text[i + 1].isascii() and text[i + 1].isdigit()
        if not prev_digit or not next_digit:
            out.append(PauseMarker(",", COMMA_MS, _byte_pos(text, i)))
    out.sort(key=lambda p: p.position)  # stable, like sort_by_key
    return out


def strip_pause_markers(text: str) -> str:
    """pause.rs:116-118"""
    return _EXPLICIT.sub(" ", text)


def parse_text_with_pauses(text: str) -> tuple[str, list[PauseMarker]]:
    """pause.rs:130-180 -> (clean_text, pauses sorted by byte position in clean_text).  Explicit markers of 0 ms are
    dropped; each marker shrinks to the one space that replaces it."""
    clean = strip_pause_markers(text)
    pauses = parse_natural_pauses(clean)
    offset = 0
    for m in _EXPLICIT.finditer(text):
        pos = max(_byte_pos(text, m.start()) - offset, 0)
        ms = _duration_ms(m.group(1), m.group(2))
        if ms > 0:
            pauses.append(PauseMarker(m.group(0), ms, pos))
        offset += len(m.group(0).encode("utf-8")) - 1
    pauses.sort(key=lambda p: p.position)
    return clean, pauses


def silence_samples(duration_ms: int, sample_rate: int = 24000) -> int:
    """pause.rs:183-185"""
    return (int(duration_ms) * int(sample_rate)) // 1000


def long_form_segments(text: str) -> list[tuple[str, object]]:
    """The segment list of generate_stream_long (tts_model.rs:1079-1107): [("text", str) | ("pause", ms)].
    Natural-pause punctuation stays attached to nothing: the comma / ellipsis itself is skipped, like the reference."""
    clean, pauses = parse_text_with_pauses(text)
    raw = clean.encode("utf-8")
    segs: list[tuple[str, object]] = []
    last = 0
    for p in pauses:
        if p.position > last:
            seg = raw[last:p.position].decode("utf-8", errors="strict")
            if seg.strip():
                segs.append(("text", seg))
        segs.append(("pause", p.duration_ms))
        last = p.position + (1 if p.original.startswith("[pause:") else len(p.original.encode("utf-8")))
    if last < len(raw):
        seg = raw[last:].decode("utf-8")
        if seg.strip():
            segs.append(("text", seg))
    return segs


# ------------------------------------------------------------------------------------------------ prompt preparation
def prepare_text_prompt(text: str) -> str:
    """tts_model.rs:1194-1227"""
    text = strip_pause_markers(text).strip()
    if not text:
        return "."
    text = text.replace("\n", " ").replace("\r", " ").replace("  ", " ")
    word_count = len(text.split())
    if not text[0].isupper():
        text = text[0].upper() + text[1:]
    if text[-1].isalnum():
        text += "."
    if word_count < 5:
        text = " " * 8 + text
    return text


def estimate_frames_after_eos(text: str) -> int:
    """tts_model.rs:1230-1237"""
    return 5 if len(text.split()) <= 4 else 3


def max_gen_len(prepared: str) -> int:
    """tts_model.rs:968: (words(prepared) + 2) * 13"""
    return (len(prepared.split()) + 2) * 13


def estimate_generation_steps(text: str) -> int:
    """tts_model.rs:1127-1130"""
    return max_gen_len(prepare_text_prompt(text))


# ------------------------------------------------------------------------------------------------ tokenizer
class _Trie:
    __slots__ = ("root",)

    def __init__(self):
        self.root: dict = {}

    def add(self, key: bytes, value: int):
        node = self.root
        for b in key:
            node = node.setdefault(b, {})
        node.setdefault(-1, value)  # first id wins for duplicate pieces, like the crate's token_to_ids map

    def prefixes(self, data: bytes, start: int):
        """(end, id) for every vocabulary piece that is a prefix of data[start:], shortest first."""
        node = self.root
        for i in range(start, len(data)):
            node = node.get(data[i])
            if node is None:
                return
            v = node.get(-1)
            if v is not None:
                yield i + 1, v


class UnigramTokenizer:
    """Unigram language-model tokenizer with a Metaspace pre-tokenizer (`split = false`), as the reference builds it.

    native (`.model`, text.rs:57-80): prepend '▁' always, no BOS.  `tokenizer.json` (assets/, WASM): whatever the file
    says (the shipped one: prepend never, `<s>` = 1 in front).  Not modelled: the json file's `added_tokens` pre-pass
    (a literal "<s>" / "</s>" / "<unk>" / "<pad>" typed into the text is split off before the Unigram model there;
    here, as in the native `.model` path, it goes through the Unigram search like any other characters)."""
    UNK_PENALTY = 10.0  # tokenizers `K_UNK_PENALTY`

    def __init__(self, vocab: list[tuple[str, float]], unk_id: int | None = 0, byte_fallback: bool = True,
                 prepend_scheme: str = "always", bos_ids: tuple[int, ...] = (), replacement: str = "▁"):
        if not vocab:
            raise ValueError("No vocabulary found")  # text.rs:231-233
        self.vocab = [(p, float(s)) for p, s in vocab]
        self.unk_id, self.byte_fallback = unk_id, byte_fallback
        self.prepend_scheme, self.bos_ids, self.replacement = prepend_scheme, tuple(bos_ids), replacement
        self.piece_to_id: dict[str, int] = {}
        self.trie = _Trie()
        for i, (p, _) in enumerate(self.vocab):
            self.piece_to_id.setdefault(p, i)
            self.trie.add(p.encode("utf-8"), i)
        self.unk_score = min(s for _, s in self.vocab) - self.UNK_PENALTY

    # ---- construction
    @classmethod
    def from_tokenizer_json(cls, path: str | Path) -> "UnigramTokenizer":
        d = json.loads(Path(path).read_text(encoding="utf-8"))
        m = d["model"]
        if m.get("type") != "Unigram":
            raise ValueError(f"unsupported tokenizer model {m.get('type')}")
        pre = d.get("pre_tokenizer") or {}
        bos: list[int] = []
        post = d.get("post_processor") or {}
        if post.get("type") == "TemplateProcessing":
            for item in post.get("single", []):
                if "SpecialToken" in item:
                    bos += post["special_tokens"][item["SpecialToken"]["id"]]["ids"]
                else:
                    break  # only leading specials are modelled (the shipped file has exactly `<s> $A`)
        return cls([(p, s) for p, s in m["vocab"]], m.get("unk_id"), bool(m.get("byte_fallback", False)),
                   pre.get("prepend_scheme", "always"), tuple(bos), pre.get("replacement", "▁"))

    @classmethod
    def from_sentencepiece_model(cls, path_or_bytes) -> "UnigramTokenizer":
        data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else Path(path_or_bytes).read_bytes()
        vocab, unk = parse_sentencepiece_vocab(bytes(data))
        return cls(vocab, unk, True, "always", ())

    @property
    def vocab_size(self) -> int:
        return len(self.vocab)

    # ---- encoding
    def _metaspace(self, text: str) -> str:
        s = text.replace(" ", self.replacement)
        if self.prepend_scheme in ("always", "first") and not s.startswith(self.replacement):
            s = self.replacement + s   # one un-split piece, so `first` == `always`
        return s

    def _viterbi(self, sentence: str) -> list[str]:
        data = sentence.encode("utf-8")
        n = len(data)
        if n == 0:
            return []
        NEG = float("-inf")
        best_score = [NEG] * (n + 1)
        best_start = [-1] * (n + 1)
        best_id = [-1] * (n + 1)
        best_score[0] = 0.0
        best_start[0] = 0
        pos = 0
        while pos < n:
            here = best_score[pos]
            b0 = data[pos]
            mblen = 1 if b0 < 0x80 else 2 if b0 < 0xE0 else 3 if b0 < 0xF0 else 4
            mblen = min(mblen, n - pos)
            single = False
            for end, tid in self.trie.prefixes(data, pos):
                cand = self.vocab[tid][1] + here
                if best_start[end] < 0 or cand > best_score[end]:
                    best_score[end], best_start[end], best_id[end] = cand, pos, tid
                if end - pos == mblen:
                    single = True
            if not single:
                end = pos + mblen
                cand = self.unk_score + here
                if best_start[end] < 0 or cand > best_score[end]:
                    best_score[end], best_start[end], best_id[end] = cand, pos, self.unk_id if self.unk_id is not None else -2
            pos += mblen
        out: list[str] = []
        fused: list[str] = []
        end = n
        while end > 0:
            start = best_start[end]
            piece = data[start:end].decode("utf-8", errors="replace")
            if self.unk_id is not None and best_id[end] == self.unk_id:
                fused.append(piece)           # consecutive unknown characters become one piece (fuse_unk)
            else:
                if fused:
                    out.append("".join(reversed(fused)))
                    fused = []
                out.append(piece)
            end = start
        if fused:
            out.append("".join(reversed(fused)))
        out.reverse()
        return out

    def encode(self, text: str, add_special_tokens: bool = True) -> list[int]:
        ids: list[int] = list(self.bos_ids) if add_special_tokens else []
        if text == "":
            return ids
        for piece in self._viterbi(self._metaspace(text)):
            tid = self.piece_to_id.get(piece)
            if tid is not None:
                ids.append(tid)
                continue
            if self.byte_fallback:
                bs = [self.piece_to_id.get(f"<0x{b:02X}>") for b in piece.encode("utf-8")]
                if all(b is not None for b in bs):
                    ids.extend(bs)
                    continue
            if self.unk_id is None:
                raise ValueError("piece outside the vocabulary and no unk id")
            ids.append(self.unk_id)
        return ids

    __call__ = encode

    def count_tokens(self, text: str) -> int:
        """conditioners/text.rs:341-348"""
        return len(self.encode(text, True))


def read_varint(data: bytes, pos: int) -> tuple[int, int]:
    """conditioners/text.rs:239-260"""
    result = shift = 0
    while True:
        if pos >= len(data):
            raise ValueError("Unexpected end of data while reading varint")
        b = data[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift >= 64:
            raise ValueError("Varint too large")


def _skip(data: bytes, pos: int, wire: int, end: int) -> int:
    if wire == 0:
        return read_varint(data, pos)[1]
    if wire == 2:
        ln, pos = read_varint(data, pos)
        return pos + ln
    if wire == 5:
        return pos + 4
    if wire == 1:
        return pos + 8
    return end  # unknown wire type: give up on this message (text.rs:205-208,226-228)


def parse_sentencepiece_vocab(data: bytes) -> tuple[list[tuple[str, float]], int]:
    """(pieces with scores, unk id) from a sentencepiece ModelProto: field 1 = repeated SentencePiece{1: piece,
    2: score f32, 3: type (2 = UNKNOWN)} (conditioners/text.rs:84-236)."""
    vocab: list[tuple[str, float]] = []
    unk = 0
    pos = 0
    while pos < len(data):
        tag, pos = read_varint(data, pos)
        field, wire = tag >> 3, tag & 7
        if field == 1 and wire == 2:
            ln, pos = read_varint(data, pos)
            end = pos + ln
            piece, score, p = "", 0.0, pos
            while p < end:
                t, p = read_varint(data, p)
                f, w = t >> 3, t & 7
                if f == 1 and w == 2:
                    l2, p = read_varint(data, p)
                    piece = data[p:p + l2].decode("utf-8", errors="replace")
                    p += l2
                elif f == 2 and w == 5:
                    if p + 4 <= len(data):
                        (score,) = struct.unpack_from("<f", data, p)
                        p += 4
                elif f == 3 and w == 0:
                    ty, p = read_varint(data, p)
                    if ty == 2:
                        unk = len(vocab)
                else:
                    p = _skip(data, p, w, end)
            if piece:
                vocab.append((piece, float(score)))
            pos = end
        elif wire in (0, 1, 2, 5):
            pos = _skip(data, pos, wire, len(data))
        else:
            break
    if not vocab:
        raise ValueError("No vocabulary found in SentencePiece model")
    return vocab, unk


# ------------------------------------------------------------------------------------------------ sentence packing
MAX_TOKENS_PER_CHUNK = 50   # tts_model.rs:604
WORDS_PER_BATCH = 35        # tts_model.rs:641
_SENTENCE_END = ".!?;:"     # tts_model.rs:610


def _split_inclusive(text: str, seps: str) -> list[str]:
    out, cur = [], []
    for ch in text:
        cur.append(ch)
        if ch in seps:
            out.append("".join(cur))
            cur = []
    if cur:
        out.append("".join(cur))
    return out


def split_into_best_sentences(text: str, count_tokens) -> list[str]:
    """tts_model.rs:603-684.  `count_tokens(str) -> int` (a failure counts as MAX_TOKENS_PER_CHUNK, like `unwrap_or`)."""
    def count(s: str) -> int:
        try:
            return int(count_tokens(s))
        except Exception:
            return MAX_TOKENS_PER_CHUNK

    prepared = prepare_text_prompt(text)
    raw = [s.strip() for s in _split_inclusive(prepared, _SENTENCE_END)]
    raw = [s for s in raw if s]
    if not raw:
        return [prepared]
    chunks: list[str] = []
    cur, cur_n = "", 0
    for sent in raw:
        n = count(sent)
        if n > MAX_TOKENS_PER_CHUNK:
            if cur:
                chunks.append(cur)
                cur, cur_n = "", 0
            words = sent.split()
            for i in range(0, len(words), WORDS_PER_BATCH):
                batch = words[i:i + WORDS_PER_BATCH]
                s = " ".join(batch)
                if count(s) <= MAX_TOKENS_PER_CHUNK:
                    chunks.append(s)
                else:
                    mid = len(batch) // 2
                    chunks.append(" ".join(batch[:mid]))
                    chunks.append(" ".join(batch[mid:]))
            continue
        if not cur:
            cur, cur_n = sent, n
        elif cur_n + n > MAX_TOKENS_PER_CHUNK:
            chunks.append(cur)
            cur, cur_n = sent, n
        else:
            cur, cur_n = cur + " " + sent, cur_n + n
    if cur:
        chunks.append(cur)
    return chunks
