#!/bin/bash
# Runs on the GPU box (under gpurun): plain bench run, then the ncu launch list of the same command,
# then one full ncu capture of the kernel named in $1 (regex, default k_fit_blobs).
# Outputs land in gpurun_out/ (copied back); summaries are committed under profiles/.
set -u
KERNEL=${1:-k_fit_small}
TAG=${2:-r01}
mkdir -p gpurun_out
CMD="python bench.py --config ${CONFIG:-2} --steps 2 --warmup 3 --latency-iters 20 --no-cpu --no-extra"
$CMD > gpurun_out/bench_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/bench_plain_$TAG.log; exit 1; }
tail -1 gpurun_out/bench_plain_$TAG.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s 3 -c 2 -f -o gpurun_out/prof_${KERNEL}_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
