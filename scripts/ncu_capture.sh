#!/bin/bash
# Round-end ncu evidence (run under gpurun; one GPU).  Each program first exits 0 WITHOUT ncu, then runs under ncu.
#   1. launch list of one eager projection step through bench.py (gpu__time_duration.sum, --clock-control none)
#   2. ncu --set full of the HBM-bound / attention kernels and of the convolution kernels of one step (scripts/ncu_target.py, one eager step;
#      the first 12 conv launches belong to set_targets' VGG pass and are skipped)
# Raw pages are exported to CSV on the box (the .ncu-rep files exceed what gpurun brings back) and summarised by scripts/ncu_summary.py.
set -x
TAG=${1:-r02b}
OUT=gpurun_out
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-aux --no-graph > $OUT/ncu_bench_plain_$TAG.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-aux --no-graph > $OUT/ncu_launches_$TAG.log 2>&1
NCU_STEPS=1 python scripts/ncu_target.py || exit 1
NCU_STEPS=1 ncu --set full --clock-control none --import-source on \
    -k regex:'upfir2|fir4|attn_|torgb|vgg_conv1|lpips_|act_bwd|pointwise|maxpool' -c 200 -f -o /tmp/mem_$TAG python scripts/ncu_target.py > $OUT/ncu_mem_$TAG.log 2>&1
ncu -i /tmp/mem_$TAG.ncu-rep --page raw --csv > $OUT/mem_${TAG}_raw.csv
NCU_STEPS=1 ncu --set full --clock-control none --import-source on \
    -k regex:'conv_tc|conv_halo' --launch-skip 12 -c 200 -f -o /tmp/conv_$TAG python scripts/ncu_target.py > $OUT/ncu_conv_$TAG.log 2>&1
ncu -i /tmp/conv_$TAG.ncu-rep --page raw --csv > $OUT/conv_${TAG}_raw.csv
ls -l $OUT/*$TAG*
