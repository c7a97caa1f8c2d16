#!/bin/bash
# usage (under gpurun --gpus 8): tools/scale_only.sh <tag> N [N ...]   -- the bench line at each N (no H2D probe)
TAG=$1; shift
mkdir -p gpurun_out
for N in "$@"; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_scale_n$N.json 2> gpurun_out/${TAG}_scale_n$N.err
  echo "N=$N rc=$?"
  python tools/show_bench.py gpurun_out/${TAG}_scale_n$N.json | grep -v "^  [a-z_0-9]* *[0-9.]* ms"
done
