#!/bin/bash
# Usage (on the GPU box, under gpurun): tools/ncu_step.sh <tag> <one_step.py args...>
# `ncu --set full` over every kernel of ONE pass of the hot path (tools/one_step.py brackets it with
# cudaProfilerStart/Stop).  The report stays in /tmp; gpurun_out/ gets the raw metric page and the
# source page (SASS + stall samples) of the first launch of every distinct kernel.
set -u
tag=$1; shift
mkdir -p gpurun_out
rep=/tmp/${tag}
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o "$rep" \
    python tools/one_step.py "$@" > gpurun_out/${tag}.log 2>&1
ncu -i ${rep}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i ${rep}.ncu-rep --page source --csv 2>/dev/null | python tools/first_blocks.py | gzip > gpurun_out/${tag}_src.csv.gz
ls -la gpurun_out/${tag}*
