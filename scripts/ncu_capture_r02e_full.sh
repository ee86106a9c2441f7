#!/bin/bash
# ncu --set full of the kernels changed after the r02c capture (LPIPS tap+pool forward / backward, streaming 1x1 convolution); one GPU.
set -x
TAG=r02e; OUT=gpurun_out
NCU_STEPS=1 python scripts/ncu_target.py || exit 1
NCU_STEPS=1 ncu --set full --clock-control none --import-source on -k regex:'lpips_tap_pool|pointwise' -c 60 -f -o /tmp/tap_$TAG python scripts/ncu_target.py > $OUT/ncu_tap_$TAG.log 2>&1
ncu -i /tmp/tap_$TAG.ncu-rep --page raw --csv > $OUT/tap_${TAG}_raw.csv
ls -l $OUT/*tap_$TAG*
