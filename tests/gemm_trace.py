"""Bring-up probe (not a test): where does one GEMM launch spend its time?"""
import ctypes as C, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pocket_tts_b200 import _lib
L = _lib.lib()
names = ["entry", "setup", "tma0", "prod_done", "tile0", "mma_done", "acc_ready", "staged", "epi_done", "freed"]
cases = [(64, 3072, 1024, 0, 0), (64, 3072, 1024, 0, 4), (64, 3072, 1024, 0, 8), (64, 1024, 1024, 0, 0), (64, 1024, 4096, 0, 0),
         (64, 4096, 1024, 0, 0), (64, 512, 512, 0, 0), (64, 512, 512, 0, 8), (1024, 1536, 512, 0, 0), (1024, 2048, 512, 0, 0), (1024, 512, 2048, 0, 0)]
for rows, feats, k, mode, split in cases:
    us = C.c_float(); n = C.c_int32(); st = np.zeros(10 * 8, np.int64)
    _lib.check(L.ptts_test_gemm_trace(0, rows, feats, k, mode, split, 50, C.byref(us), st.ctypes.data_as(C.c_void_p), 8, C.byref(n)))
    print(f"rows={rows} F={feats} K={k} split={split or 'auto'}: {us.value:.2f} us/launch back-to-back")
    print("   cta 0", {nm: int(v) for nm, v in zip(names, st[0:10])})
