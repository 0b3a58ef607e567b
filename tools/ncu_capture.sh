#!/bin/bash
# Usage (on the GPU box, under gpurun): tools/ncu_capture.sh <tag> <kernel-regex> <skip> <count> <command...>
# Captures `ncu --set full` for <count> launches of kernels matching the regex after skipping <skip> of
# them, keeps the (large) report in /tmp and writes only text extracts to gpurun_out/: the raw metric
# page and the source page (SASS + stall samples; needs -lineinfo).
set -u
tag=$1; regex=$2; skip=$3; count=$4; shift 4
mkdir -p gpurun_out
rep=/tmp/${tag}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$regex" --launch-skip "$skip" \
    --launch-count "$count" -f -o "$rep" "$@" > gpurun_out/${tag}.log 2>&1
ncu -i ${rep}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i ${rep}.ncu-rep --page source --csv > gpurun_out/${tag}_src.csv 2>/dev/null
gzip -f gpurun_out/${tag}_src.csv
ls -la gpurun_out/${tag}*
