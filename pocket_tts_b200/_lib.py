"""ctypes binding of include/ptts.h.  There is no fallback: if libptts_cuda.so is missing the
import fails, and every compute entry point fails without a CUDA device."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libptts_cuda.so"

# every symbol include/ptts.h (product ABI) and include/ptts_internal.h (test hooks, probes) declare; tests/test_abi.py
# checks both headers against these lists
PRODUCT_SYMBOLS = [
    "ptts_last_error", "ptts_abi_version", "ptts_engine_create", "ptts_engine_destroy", "ptts_engine_set_lsd_steps",
    "ptts_engine_set_codec_group",
    "ptts_voice_from_prompt", "ptts_voice_from_pcm", "ptts_audio_prompt_from_pcm", "ptts_voice_destroy", "ptts_voice_len",
    "ptts_voice_save", "ptts_voice_load", "ptts_config_check",
    "ptts_streams_open", "ptts_step", "ptts_step_begin", "ptts_step_flags", "ptts_step_pcm", "ptts_step_pcm_i16", "ptts_step_device",
    "ptts_sync", "ptts_stream_set_feedback", "ptts_stream_close", "ptts_streams_close", "ptts_stream_frames",
    "ptts_sched_create", "ptts_sched_destroy", "ptts_sched_submit", "ptts_sched_run", "ptts_sched_result_samples", "ptts_sched_result", "ptts_sched_result_view",
    "ptts_sched_steps",
]
INTERNAL_SYMBOLS = [
    "ptts_debug_read", "ptts_launch_count", "ptts_step_timed", "ptts_cuda_stream", "ptts_profile_enable", "ptts_profile_report",
    "ptts_profile_overhead", "ptts_test_gemm", "ptts_test_gemm_int8", "ptts_test_gemv_launches", "ptts_test_gemm_trace", "ptts_test_conv1d", "ptts_test_convtr1d",
    "ptts_test_noise", "ptts_debug_f16_overflow", "ptts_profile_gemm_replay",
]
SYMBOLS = PRODUCT_SYMBOLS + INTERNAL_SYMBOLS


class TensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 4), ("data", C.c_void_p)]


class EngineCfg(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_slots", C.c_int32), ("max_batch", C.c_int32),
                ("kv_capacity", C.c_int32), ("weight_mode", C.c_int32), ("use_cuda_graph", C.c_int32),
                ("debug_gemm", C.c_int32), ("reserved", C.c_int32 * 9)]


class StreamParams(C.Structure):
    _fields_ = [("max_gen_len", C.c_int32), ("frames_after_eos", C.c_int32), ("eos_threshold", C.c_float),
                ("temp", C.c_float), ("seed", C.c_uint64), ("noise", C.c_void_p)]


class Segment(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_tokens", C.c_int32), ("tokens", C.c_void_p), ("params", StreamParams),
                ("pause_ms", C.c_int32), ("reserved", C.c_int32)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: run `python -m pocket_tts_b200.build` (nvcc, sm_100a). "
                          "There is no CPU or PyTorch fallback for this path.")
    from . import build as _build
    if _build.needs_build():
        # the binary in the tree was not built from the sources in the tree (content hash, not file times): rebuild, or
        # refuse -- never run a stale kernel silently
        try:
            _build.build()
        except Exception as ex:
            raise ImportError(f"{LIB_PATH} is stale (source hash differs from {_build.STAMP.name}) and cannot be rebuilt here: {ex}")
    L = C.CDLL(str(LIB_PATH))
    vp, i32, i64, fp = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_float)
    L.ptts_last_error.restype = C.c_char_p
    L.ptts_abi_version.restype = i32
    L.ptts_engine_create.argtypes = [C.POINTER(EngineCfg), C.POINTER(TensorDesc), i32, C.POINTER(vp)]
    L.ptts_engine_destroy.argtypes = [vp]
    L.ptts_engine_destroy.restype = None
    L.ptts_engine_set_lsd_steps.argtypes = [vp, i32]
    L.ptts_engine_set_codec_group.argtypes = [vp, i32]
    L.ptts_voice_from_prompt.argtypes = [vp, vp, i32, C.POINTER(vp)]
    L.ptts_voice_from_pcm.argtypes = [vp, vp, i32, C.POINTER(vp)]
    L.ptts_audio_prompt_from_pcm.argtypes = [vp, vp, i32, vp, i32, C.POINTER(i32)]
    L.ptts_voice_destroy.argtypes = [vp, vp]
    L.ptts_voice_destroy.restype = None
    L.ptts_voice_len.argtypes = [vp]
    L.ptts_streams_open.argtypes = [vp, i32, C.POINTER(vp), vp, vp, C.POINTER(StreamParams), vp]
    L.ptts_step.argtypes = [vp, vp, i32, vp, vp, vp, vp]
    L.ptts_step_begin.argtypes = [vp, vp, i32, i32]
    L.ptts_step_begin.restype = i64
    L.ptts_step_flags.argtypes = [vp, i64, vp, vp, vp]
    L.ptts_step_pcm.argtypes = [vp, i64, vp]
    L.ptts_step_device.argtypes = [vp, vp, i32]
    L.ptts_sync.argtypes = [vp]
    L.ptts_stream_set_feedback.argtypes = [vp, i32, vp]
    L.ptts_stream_close.argtypes = [vp, i32]
    L.ptts_stream_frames.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
    L.ptts_debug_read.argtypes = [vp, C.c_char_p, i32, vp, i64]
    L.ptts_debug_read.restype = i64
    L.ptts_launch_count.argtypes = [vp, i32]
    L.ptts_launch_count.restype = i64
    L.ptts_step_timed.argtypes = [vp, vp, i32, vp]
    L.ptts_cuda_stream.argtypes = [vp]
    L.ptts_cuda_stream.restype = vp
    L.ptts_profile_enable.argtypes = [vp, i32]
    L.ptts_profile_report.argtypes = [vp, C.c_char_p, i64]
    L.ptts_profile_report.restype = i64
    L.ptts_profile_overhead.argtypes = [vp, vp]
    L.ptts_test_gemm.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32]
    L.ptts_test_gemm_int8.argtypes = [i32, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    L.ptts_test_gemv_launches.argtypes = []
    L.ptts_test_gemv_launches.restype = i64
    L.ptts_test_gemm_trace.argtypes = [i32, i32, i32, i32, i32, i32, i32, vp, vp, i32, vp]
    L.ptts_test_conv1d.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32]
    L.ptts_test_convtr1d.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32]
    L.ptts_voice_save.argtypes = [vp, vp, C.c_char_p, i32]
    L.ptts_voice_load.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.ptts_config_check.argtypes = [C.c_char_p]
    L.ptts_step_pcm_i16.argtypes = [vp, i64, vp]
    L.ptts_streams_close.argtypes = [vp, vp, i32]
    L.ptts_sched_create.argtypes = [vp, vp, i32, C.POINTER(vp)]
    L.ptts_sched_destroy.argtypes = [vp]
    L.ptts_sched_destroy.restype = None
    L.ptts_sched_submit.argtypes = [vp, C.POINTER(Segment), i32]
    L.ptts_sched_submit.restype = i64
    L.ptts_sched_run.argtypes = [vp, i32]
    L.ptts_sched_result_samples.argtypes = [vp, i64]
    L.ptts_sched_result_samples.restype = i64
    L.ptts_sched_result.argtypes = [vp, i64, vp, i64]
    L.ptts_sched_result_view.argtypes = [vp, i64, C.POINTER(vp), C.POINTER(i64)]
    L.ptts_sched_steps.argtypes = [vp]
    L.ptts_sched_steps.restype = i64
    L.ptts_test_noise.argtypes = [i32, C.c_uint64, i32, vp]
    L.ptts_profile_gemm_replay.argtypes = [vp, i32, i32, vp, vp]
    L.ptts_debug_f16_overflow.argtypes = [vp, C.POINTER(i64)]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is C.c_int:  # default: status code
            fn.restype = i32
    _lib = L
    return L


class PttsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ptts error {code}: {msg}")
        self.code = code


def check(status: int) -> int:
    if status < 0:
        raise PttsError(status, lib().ptts_last_error().decode("utf-8", "replace"))
    return status
