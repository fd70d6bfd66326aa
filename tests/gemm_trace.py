"""Bring-up probe (not a test): where does one GEMM launch spend its time?"""
import ctypes as C, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pocket_tts_b200 import _lib
L = _lib.lib()
cases = [(122880, 64, 64, 11, 0, "res9b"), (122880, 64, 192, 10, 0, "res9a-like"), (30720, 256, 256, 10, 0, "convtr8"), (30720, 128, 64, 11, 0, "res6b")]
for rows, feats, k, mode, split, nm in cases:
    us = C.c_float(); n = C.c_int32(); st = np.zeros(16 * 8, np.int64)
    _lib.check(L.ptts_test_gemm_trace(0, rows, feats, k, mode, split, 20, C.byref(us), st.ctypes.data_as(C.c_void_p), 8, C.byref(n)))
    print(f"{nm}: rows={rows} F={feats} K={k}: {us.value:.2f} us/launch back-to-back")
    s = st[:16]
    for j in range(3):
        print(f"   tile {j}: acc_ready {s[j*4]}  staged {s[j*4+1]}  barrier {s[j*4+2]}  stored {s[j*4+3]}")
    print("   epilogue entry", s[12], "exit", s[13])
