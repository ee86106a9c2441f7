set -x
OUT=gpurun_out; TAG=r02c
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-aux --no-graph > $OUT/ncu_bench_plain_$TAG.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-aux --no-graph > $OUT/ncu_launches_$TAG.log 2>&1
NCU_STEPS=1 ncu --set full --clock-control none --import-source on -k regex:'attn_|act_bwd|conv_tc2_kernel<128' -c 200 -f -o /tmp/attn_$TAG python scripts/ncu_target.py > $OUT/ncu_attn_$TAG.log 2>&1
ncu -i /tmp/attn_$TAG.ncu-rep --page raw --csv > $OUT/attn_${TAG}_raw.csv
ls -l $OUT/*$TAG*
