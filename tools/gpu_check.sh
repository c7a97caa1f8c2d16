#!/bin/bash
# usage (under gpurun): tools/gpu_check.sh <tag> [pytest-args]   -- gpu tests, then one short bench line, into gpurun_out/
TAG=${1:-x}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_tests.log
tail -12 gpurun_out/${TAG}_tests.log
timeout 400 python bench.py --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
tail -c 1200 gpurun_out/${TAG}_bench.err
python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -60
