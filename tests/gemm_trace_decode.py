"""Bring-up probe (not a test): stage stamps (%globaltimer, ns) of decode-shaped GEMM launches.
slots: 0 entry, 1 prologue done, 2 first activation TMA issued, 3 producer done, 4 first stage landed, 5 MMAs issued,
6 accumulator ready, 7 staged in smem, 8 stored, 9 exit."""
import ctypes as C, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pocket_tts_b200 import _lib
L = _lib.lib()
cases = [(64, 512, 512, "flow.mlp"), (64, 1024, 1024, "out_proj"), (64, 3072, 1024, "in_proj"), (64, 4096, 1024, "linear1"),
         (64, 1024, 4096, "linear2"), (1024, 1536, 512, "mimi.in_proj"), (1024, 2048, 512, "mimi.linear1")]
MAXC = 160
for rows, feats, k, nm in cases:
    us = C.c_float(); n = C.c_int32(); st = np.zeros(16 * MAXC, np.int64)
    _lib.check(L.ptts_test_gemm_trace(0, rows, feats, k, 0, 0, 20, C.byref(us), st.ctypes.data_as(C.c_void_p), MAXC, C.byref(n)))
    s = st.reshape(MAXC, 16)[: n.value]
    print(f"{nm}: rows={rows} F={feats} K={k}: {us.value:.2f} us/launch back-to-back, {n.value} CTAs")
    ok = s[:, 9] >= 0
    rel = s[ok][:, :10] - s[ok][:, :1]
    print("   median per-CTA stamps rel. entry:", np.median(rel, axis=0).astype(int).tolist())
    print("   entry spread", int(s[ok][:, 0].max() - s[ok][:, 0].min()), " kernel span", int(s[ok][:, 9].max() - s[ok][:, 0].min()))
