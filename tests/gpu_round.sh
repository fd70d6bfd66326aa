# One GPU session of the round: bench line, ncu launch list and full-set capture of one mid-utterance decode step.
set -x
python bench.py > gpurun_out/r02_bench_b200x1.json 2> gpurun_out/r02_bench_b200x1.err
python bench.py --streams 1 --no-cpu-baseline --longform 0 --steps 3 > gpurun_out/r02_bench_f16_b1.json 2>&1
python bench.py --streams 1 --lm-step-kernel --no-cpu-baseline --longform 0 --steps 3 > gpurun_out/r02_bench_f16_b1_stepkernel.json 2>&1
python bench.py --streams 16 --no-cpu-baseline --longform 0 --steps 3 > gpurun_out/r02_bench_f16_b16.json 2>&1
python bench.py --streams 16 --lm-step-kernel --no-cpu-baseline --longform 0 --steps 3 > gpurun_out/r02_bench_f16_b16_stepkernel.json 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_ncu.csv python tests/profile_step.py 64 8 1 > gpurun_out/ncu1.log 2>&1
ncu --profile-from-start off --set full --import-source on --clock-control none -f -o /tmp/r02_full python tests/profile_step.py 64 8 1 > gpurun_out/ncu2.log 2>&1
ncu -i /tmp/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
