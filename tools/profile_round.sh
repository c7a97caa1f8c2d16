#!/bin/bash
# usage (under gpurun): tools/profile_round.sh <tag> [configs, default "2 5"]   (one config per call keeps gpurun_out under 64 MiB)
#   config 2 (128 frames): plain bench line, ncu launch list, ncu --set full of every pipeline kernel of one batch
#   config 5 (4K, batch 16): plain bench line, ncu --set full of every pipeline kernel of one batch
TAG=${1:-r02}
CFGS=${2:-2 5}
mkdir -p gpurun_out
K='^k_(pre|blur|tile|ccl|bou|sel|sca|fit|qua|dec)'
for CFG in $CFGS; do
  B=128; [ $CFG = 5 ] && B=16
  CMD="python bench.py --config $CFG --steps 2 --warmup 3 --latency-iters 20 --batch $B --no-cpu --no-extra"
  $CMD > gpurun_out/bench_plain_${TAG}_c$CFG.json 2> gpurun_out/bench_plain_${TAG}_c$CFG.err || { echo "plain run failed (config $CFG)"; tail -5 gpurun_out/bench_plain_${TAG}_c$CFG.err; continue; }
  python tools/show_bench.py gpurun_out/bench_plain_${TAG}_c$CFG.json 2>/dev/null | sed -n 1,3p
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_c$CFG.csv $CMD > /dev/null 2>&1
  echo "launch list (config $CFG) rc=$?"
  # one batch = 14 kernels (15 with the huge tier / blur ...): skip the first warm-up batches, take 16 launches
  ncu --set full --clock-control none --import-source on -k regex:$K -s 45 -c 16 -f -o gpurun_out/prof_${TAG}_c$CFG $CMD > gpurun_out/ncu_${TAG}_c$CFG.log 2>&1
  echo "full capture (config $CFG) rc=$?"; tail -1 gpurun_out/ncu_${TAG}_c$CFG.log | cut -c1-200
done
ls -la gpurun_out/prof_${TAG}_c*.ncu-rep
