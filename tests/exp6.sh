run() { echo "== $*"; env "$@" PTTS_DIAG_TIMES=1 python tests/overlap_probe.py 64 2>&1 | tail -2; }
run PTTS_DIAG_SKIP=1
run PTTS_DIAG_SKIP=3
run PTTS_DIAG_B_STOP=1
run PTTS_DIAG_B_STOP=2
run PTTS_DIAG_B_STOP=0
