#!/bin/bash
# usage: tools/ncu_kernel.sh <kernel-regex> <tag> [skip] [batch] [count]   (runs under gpurun)
set -u
K=$1; TAG=$2; SKIP=${3:-3}; BATCH=${4:-128}; COUNT=${5:-1}
mkdir -p gpurun_out
CMD="python bench.py --config ${CONFIG:-2} --steps 1 --warmup 3 --latency-iters 5 --batch $BATCH --no-cpu --no-extra"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $COUNT -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$TAG.log | cut -c1-300
