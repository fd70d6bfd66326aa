"""Output formats of the generation path: the reference's 16-bit PCM conversion and WAV container
(crates/pocket-tts/src/audio.rs:110-184) and the `/stream` chunk framing (pocket-tts-cli/src/server/handlers.rs:259-276:
raw little-endian i16 bytes of each frame, in order).  Host code over numpy; the audio itself comes from the engine.
"""
from __future__ import annotations

import io
import struct
from pathlib import Path

import numpy as np


def _channels_first(audio: np.ndarray) -> np.ndarray:
    a = np.asarray(audio, np.float32)
    if a.ndim != 2:
        raise ValueError(f"Expected audio tensor with shape [channels, samples], got {list(a.shape)}")  # audio.rs:112-117
    return a


def pcm_i16(audio: np.ndarray) -> np.ndarray:
    """f32 [channels, samples] -> i16 [samples, channels] (interleaved): clamp to [-1, 1], scale by 32767, truncate
    toward zero; NaN -> 0 (Rust `as i16` semantics, audio.rs:139-141)."""
    a = _channels_first(audio)
    v = np.clip(a, -1.0, 1.0) * np.float32(32767.0)
    v = np.where(np.isnan(v), np.float32(0.0), v)
    return np.trunc(v).astype(np.int16).T.copy()


def pcm_i16_le_bytes(audio: np.ndarray) -> bytes:
    """audio.rs:110-122"""
    return pcm_i16(audio).astype("<i2").tobytes()


def stream_chunk_bytes(frame: np.ndarray) -> bytes:
    """One `/stream` body chunk from a generate_stream item f32 [1, 1, 1920] (handlers.rs:266-272: squeeze(0) then
    pcm_i16_le_bytes)."""
    f = np.asarray(frame, np.float32)
    return pcm_i16_le_bytes(f.reshape(f.shape[-2], f.shape[-1]) if f.ndim == 3 else f)


def wav_bytes(audio: np.ndarray, sample_rate: int) -> bytes:
    """16-bit integer PCM WAV (`hound` WavSpec{bits_per_sample: 16, Int}, audio.rs:155-184): canonical 44-byte header."""
    a = _channels_first(audio)
    data = pcm_i16_le_bytes(a)
    ch = a.shape[0]
    block = ch * 2
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, 1, ch, int(sample_rate), int(sample_rate) * block, block, 16) + b"data" + struct.pack("<I", len(data))
    return hdr + data


def write_wav(path: str | Path | io.IOBase, audio: np.ndarray, sample_rate: int) -> None:
    """audio.rs:146-150"""
    b = wav_bytes(audio, sample_rate)
    if hasattr(path, "write"):
        path.write(b)
    else:
        Path(path).write_bytes(b)


def normalize_peak(audio: np.ndarray) -> np.ndarray:
    """audio.rs:186-193"""
    a = np.asarray(audio, np.float32)
    m = float(np.abs(a).max()) if a.size else 0.0
    return a * np.float32(1.0 / m) if m > 0.0 else a.copy()
