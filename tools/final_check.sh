#!/bin/bash
# usage (under gpurun): tools/final_check.sh <tag> [noref]  -- what the driver runs at round end: gpu tests, smoke, both bench arms
# ("noref" skips the CPU reference arm: 90 s of host time that does not depend on the kernels)
TAG=${1:-final}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
if [ "$2" != noref ]; then
  T0=$(date +%s)
  timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "reference arm rc=$? ($(( $(date +%s) - T0 )) s)"
fi
T0=$(date +%s)
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "our arm rc=$? ($(( $(date +%s) - T0 )) s)"
tail -c 600 gpurun_out/${TAG}_bench.err
python tools/show_bench.py gpurun_out/${TAG}_bench.json
if [ "$2" != noref ]; then
python - $TAG <<'P'
import json,sys
d=json.loads(open("gpurun_out/%s_bench_ref.json" % sys.argv[1] if len(sys.argv)>1 else "gpurun_out/final_bench_ref.json").read().strip().splitlines()[-1])
print("reference arm:", round(d["value"],1), d["unit"], d["cpu_baseline"]["cores"], "threads", d["cpu_baseline"].get("cpu_model"), "ms/step", round(d["ms_per_step"]))
P
fi
