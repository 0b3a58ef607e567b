#!/bin/bash
# Usage (on the GPU box, under gpurun): tools/sanitize.sh <tag>
# compute-sanitizer memcheck, racecheck, initcheck and synccheck over tools/sanitize.py; the logs go to
# gpurun_out/<tag>_{memcheck,racecheck,initcheck,synccheck}.log (copy the summaries to profiles/).
set -u
tag=${1:-sanitize}
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py \
        > gpurun_out/${tag}_${tool}.log 2>&1
    echo "$tool rc=$?" >> gpurun_out/${tag}_${tool}.log
    tail -n 4 gpurun_out/${tag}_${tool}.log
done
