#!/bin/bash
# usage (under gpurun --gpus 8): tools/scale_check.sh <tag> [N ...]   -- bare H2D probe, then the bench at each N
TAG=${1:-r02}; shift
NS=${@:-8 4}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
lscpu | head -25 > gpurun_out/${TAG}_lscpu.txt 2>&1
numactl -H >> gpurun_out/${TAG}_lscpu.txt 2>&1
timeout 400 python tools/h2d_probe.py --gpus 1,2,4,8 --iters 10 --out gpurun_out/h2d_probe_${TAG}.json > gpurun_out/${TAG}_h2d.log 2>&1
echo "probe rc=$?"; tail -1 gpurun_out/${TAG}_h2d.log | cut -c1-1500
for N in $NS; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_scale_n$N.json 2> gpurun_out/${TAG}_scale_n$N.err
  echo "N=$N rc=$?"; tail -c 600 gpurun_out/${TAG}_scale_n$N.err
  python tools/show_bench.py gpurun_out/${TAG}_scale_n$N.json | grep -v "^  [a-z_0-9]* *[0-9.]* ms"
done
