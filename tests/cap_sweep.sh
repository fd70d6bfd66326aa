#!/bin/bash
run() { echo -n "$* : "; env "$@" python tests/replay_probe.py 64 2>&1 | tail -1; }
run A=1
run PTTS_LIN2_CTAS=64
run PTTS_LIN2_CTAS=64 PTTS_OUTPROJ_CTAS=64
run PTTS_OUTPROJ_CTAS=64
run PTTS_INPROJ_CTAS=96
run PTTS_LIN2_CTAS=64 PTTS_LIN1_CTAS=128
run PTTS_LIN1_CTAS=32
