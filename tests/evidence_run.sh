#!/bin/bash
# Round evidence (not a test): GPU test log, smoke, bench lines, full-set ncu rows of the new kernels.  Whole-step captures:
# see profiles/README.md (.ncu-rep files stay in /tmp: gpurun_out/ carries at most 64 MiB back).
# bash tests/evidence_run.sh <tag>   -> gpurun_out/<tag>_*
T=${1:-r02}
O=gpurun_out
python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1; tail -2 $O/${T}_smoke.log
python -m pytest tests -q -m gpu > $O/${T}_pytest_gpu.log 2>&1; tail -2 $O/${T}_pytest_gpu.log
python bench.py > $O/${T}_bench_b200x1.json 2> $O/${T}_bench_b200x1.err; tail -c 300 $O/${T}_bench_b200x1.err
for s in 1 4 16 256 512; do python bench.py --streams $s --steps 3 --warmup 3 --no-cpu-baseline --longform 0 > $O/${T}_bench_f16_b$s.json 2> $O/${T}_bench_f16_b$s.err; done
python bench.py --streams 256 --int8 --steps 3 --warmup 3 --no-cpu-baseline --longform 0 > $O/${T}_bench_int8_b256.json 2> $O/${T}_bench_int8_b256.err
python bench.py --streams 1 --int8 --steps 3 --warmup 3 --no-cpu-baseline --longform 0 > $O/${T}_bench_int8_b1.json 2> $O/${T}_bench_int8_b1.err
ncu --set full --import-source on --clock-control none -k regex:"prefill_mma|rope_append" -c 4 -o /tmp/${T}_prefill_full -f python tests/prefill_probe.py 64 40 > $O/${T}_ncu_prefill.log 2>&1
ncu -i /tmp/${T}_prefill_full.ncu-rep --page raw --csv > $O/${T}_prefill_raw.csv 2>/dev/null
ls -la $O | tail -5
