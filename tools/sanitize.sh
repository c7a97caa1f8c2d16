#!/bin/bash
# usage (under gpurun): tools/sanitize.sh <tag>   -- compute-sanitizer memcheck + racecheck + initcheck on small frames
TAG=${1:-san}
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'P'
import numpy as np, sys
sys.path.insert(0, ".")
from ros_vision_b200 import detector as D, synth
D.load_library()
for (w, h, fmt, dec, sigma, fam) in ((328, 248, "yuyv", 2, 0.0, "tag36h11"), (320, 240, "bgr", 1, 0.8, "tag36h11"), (320, 240, "gray", 2, 0.0, "tag16h5")):
    sc = synth.make_scene(w, h, 7, 2, side_range=(40, 80), noise_sigma=4.0, family=fam)
    fr = {"gray": lambda g: g, "yuyv": synth.gray_to_yuyv, "bgr": synth.gray_to_bgr}[fmt](sc.gray)
    det = D.GpuDetector(w, h, fmt, quad_decimate=dec, quad_sigma=sigma, max_batch=2, families=[fam])
    det.DetectBatch([fr, fr])
    print(fmt, dec, fam, [int(i) for i in det.Detections(0)["id"]], det.FrameInfo(0).num_points)
    det.close()
P
for TOOL in memcheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $TOOL --error-exitcode 9 python /tmp/san_case.py > gpurun_out/${TAG}_$TOOL.log 2>&1
  echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/${TAG}_$TOOL.log | tail -3
done
