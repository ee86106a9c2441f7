#!/bin/bash
# ncu --set full of the dominant kernel of the final build (conv_tc2_kernel: CTA-pair tcgen05 implicit-GEMM conv), 12 launches of one eager step
# (the first 12 conv launches belong to set_targets' VGG pass and are skipped); one GPU.
set -x
TAG=r02f; OUT=gpurun_out
NCU_STEPS=1 python scripts/ncu_target.py || exit 1
NCU_STEPS=1 ncu --set full --clock-control none --import-source on -k regex:'conv_tc2_kernel' --launch-skip 12 -c 14 -f -o /tmp/top_$TAG python scripts/ncu_target.py > $OUT/ncu_top_$TAG.log 2>&1
ncu -i /tmp/top_$TAG.ncu-rep --page raw --csv > $OUT/top_${TAG}_raw.csv
ls -l $OUT/*top_$TAG*
