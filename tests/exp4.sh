run() { echo "== $*"; env "$@" PTTS_DIAG_TIMES=1 python tests/overlap_probe.py 64 2>&1 | tail -2; }
for cg in 2 4; do for g in 96 112 128; do for kb in 200 100 64; do
  b=$((148-g))
  run PTTS_CODEC_GROUP=$cg PTTS_LM_STEP_KERNEL=1 PTTS_LM_CTAS=$g PTTS_B_SMS=$b PTTS_GEMM_SMEM_KB=$kb
done; done; done
run PTTS_CODEC_GROUP=1 PTTS_GEMM_SMEM_KB=100
run PTTS_CODEC_GROUP=1 PTTS_GEMM_SMEM_KB=64
run PTTS_CODEC_GROUP=2 PTTS_GEMM_SMEM_KB=64
