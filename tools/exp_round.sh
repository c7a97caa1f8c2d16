#!/bin/bash
# usage (under gpurun): tools/exp_round.sh <tag> "<exp masks>" [pytest -k expression] [lanes for mask 0]
# parity tests of the default build, then per-kernel times / throughput of every B200TAG_EXP variant (tools/exp_kernels.py)
TAG=${1:-x}; MASKS=${2:-0}; KEXPR=${3:-}; L0=${4:-2}
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$KEXPR" > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
fi
for M in $MASKS; do
  L=2; [ "$M" = 0 ] && L=$L0
  B200TAG_EXP=$M timeout 200 python tools/exp_kernels.py --lanes $L 2> gpurun_out/${TAG}_exp$M.err | tee -a gpurun_out/${TAG}_exp.jsonl
done
B200TAG_EXP=0 timeout 200 python tools/exp_kernels.py --config 5 --batch 16 2>> gpurun_out/${TAG}_exp0.err | tee -a gpurun_out/${TAG}_exp.jsonl
