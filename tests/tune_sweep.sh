#!/bin/bash
# Bring-up probe (not a test): per-call-site (feature tile, split cap) of the codec GEMMs, step time and codec-alone time.
run() { echo "== $*"; env "$@" python tests/overlap_probe.py 64 2>&1 | grep "us/step" | tail -1; env "$@" PTTS_DIAG_SKIP=2 python tests/overlap_probe.py 64 2>&1 | grep "us/step" | tail -1; }
run A=1
run PTTS_TUNE_CONV0=128,128
run PTTS_TUNE_CONV0=256,128
run PTTS_TUNE_CONV0=64,128
run PTTS_TUNE_MLIN2=128,128
run PTTS_TUNE_MLIN2=256,128
run PTTS_TUNE_CT2=256,96
run PTTS_TUNE_MLIN1=256,128
run PTTS_TUNE_MINPROJ=256,96
run PTTS_TUNE_MOUTPROJ=128,128
run PTTS_TUNE_CT5=128,0
