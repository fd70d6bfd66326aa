"""Bring-up probe: a handful of launches of one GEMM shape (for ncu source-level sampling)."""
import ctypes as C, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pocket_tts_b200 import _lib
L = _lib.lib()
rows, feats, k, mode, split = [int(x) for x in sys.argv[1:6]]
us = C.c_float(); n = C.c_int32(); st = np.zeros(16 * 8, np.int64)
_lib.check(L.ptts_test_gemm_trace(0, rows, feats, k, mode, split, 5, C.byref(us), st.ctypes.data_as(C.c_void_p), 8, C.byref(n)))
print(us.value)
