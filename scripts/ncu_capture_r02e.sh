#!/bin/bash
# Round-2 closing evidence (run under gpurun; one GPU).  Every program exits 0 WITHOUT ncu before it runs under ncu.
#   1. the full -m gpu test suite
#   2. the default bench line (python bench.py)
#   3. ncu launch list of one eager projection step with gpu__time_duration.sum + DRAM bytes read / written per launch
#      (summarised by scripts/summarize_traffic.py; also the source of bench.py's roofline.traffic)
#   4. scripts/bench_membound.py / bench_attn.py: CUDA-event GB/s of the HBM-bound kernels at the bench shapes
set -x
TAG=${1:-r02e}
OUT=gpurun_out
python -m pytest tests -m gpu -x -q > $OUT/t_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/t_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err || exit 1
python bench.py --steps 100 --warmup 5 --no-aux --no-cpu-baseline > $OUT/bench_steps100_$TAG.json 2>/dev/null
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-aux --no-graph > $OUT/ncu_bench_plain_$TAG.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --print-units base --clock-control none --csv \
    --log-file $OUT/traffic_$TAG.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-aux --no-graph > $OUT/ncu_traffic_$TAG.log 2>&1
python bench.py --workload pairs > $OUT/pairs_n1_$TAG.json 2>/dev/null
python scripts/bench_membound.py > $OUT/membound_$TAG.txt 2>&1
timeout 120 python scripts/bench_attn.py > $OUT/attn_$TAG.txt 2>&1
ls -l $OUT/*$TAG*
