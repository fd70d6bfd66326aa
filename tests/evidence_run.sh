#!/bin/bash
# Round evidence (not a test): bench lines and the full-set ncu capture of the small-batch GEMV.  Whole-step captures:
# see profiles/README.md (the .ncu-rep files are deleted after their CSV export: gpurun_out/ carries at most 64 MiB back).
# bash tests/evidence_run.sh <tag>   -> gpurun_out/<tag>_*
T=${1:-r02}
O=gpurun_out
python bench.py > $O/${T}_bench_b200x1.json 2> $O/${T}_bench_b200x1.err; tail -c 300 $O/${T}_bench_b200x1.err
for s in 1 4; do python bench.py --streams $s --steps 3 --warmup 3 --no-cpu-baseline --longform 0 > $O/${T}_bench_f16_b$s.json 2> $O/${T}_bench_f16_b$s.err; done
python bench.py --streams 1 --int8 --steps 3 --warmup 3 --no-cpu-baseline --longform 0 > $O/${T}_bench_int8_b1.json 2> $O/${T}_bench_int8_b1.err
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:gemv -c 12 -o /tmp/${T}_gemv_full -f python tests/profile_step.py 1 8 1 > $O/${T}_ncu_gemv.log 2>&1
ncu -i /tmp/${T}_gemv_full.ncu-rep --page raw --csv > $O/${T}_gemv_raw.csv 2>/dev/null
ls -la $O | tail -12
