run() { echo "== $*"; env "$@" PTTS_DIAG_TIMES=1 python tests/overlap_probe.py 64 2>&1 | tail -2; }
for cg in 1 2; do
run PTTS_CODEC_GROUP=$cg PTTS_MAX_CTAS_B=24 PTTS_B_SMS=64
run PTTS_CODEC_GROUP=$cg PTTS_MAX_CTAS_B=16 PTTS_B_SMS=32
run PTTS_CODEC_GROUP=$cg PTTS_MAX_CTAS_B=8 PTTS_B_SMS=16
done
run PTTS_CODEC_GROUP=1 PTTS_PRIO=0
