#!/usr/bin/env python3
"""Records the classic CPU detector's (oracle/classic_detector.c) answers on the reference's own test images, the way
src/apriltags_cuda/test/gpu_detector_test.cu:104-157 runs them: the CPU detector on cv::COLOR_BGR2GRAY of the frame, the
GPU detector (here: the GPU-semantics oracle) on the YUYV conversion of the same frame with the test's intrinsics, then
CpuAndGpuEqual -- same id, centre and the four corners within 0.5 px.

Run in the build container (needs /root/reference and cv2); the output tests/golden/classic_full_frames.json travels
with the repo, /root/reference does not.
"""
import json
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import match_corner_sets  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

REF = "/root/reference/src/apriltags_cuda/test/data/"
CAM = (905.495617, 609.916016, 907.909470, 352.682645)  # gpu_detector_test.cu:63-66
DIST = (0.059238, -0.075154, -0.003801, 0.001113, 0.0)  # gpu_detector_test.cu:69-73


def dets_json(d):
    return [{"id": int(x["id"]), "hamming": int(x["hamming"]), "decision_margin": float(x["decision_margin"]),
             "c": x["c"].tolist(), "p": x["p"].tolist()} for x in d]


def main():
    out = {}
    for name, expect in (("colorimage.jpg", 1), ("colorimage_notags.jpg", 0)):
        bgr = cv2.imread(REF + name)
        H, W = bgr.shape[:2]
        gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)            # gpu_detector_test.cu:106,115,125
        yuyv = cv2.cvtColor(bgr, cv2.COLOR_BGR2YUV_YUYV)         # :43-47
        cpu, nq = po.classic_detect(po.make_config(W, H, "gray", 2, 0.0), gray)
        gpu = po.detect(po.make_config(W, H, "yuyv", 2, 0.0, camera=CAM, dist=DIST), yuyv).detections
        assert len(cpu) == expect, (name, len(cpu))               # CpuDetectsAprilTag / CpuNoAprilTagDetections
        assert len(gpu) == expect, (name, len(gpu))
        rec = {"width": W, "height": H, "classic_on_bgr2gray": dets_json(cpu), "classic_quads": nq,
               "gpu_semantics_on_yuyv": dets_json(gpu)}
        if expect:                                                # CpuAndGpuEqual, :122-157
            assert int(cpu[0]["id"]) == int(gpu[0]["id"])
            dc = float(np.abs(cpu[0]["c"] - gpu[0]["c"]).max())
            dp = float(np.abs(cpu[0]["p"] - gpu[0]["p"]).max())
            assert dc < 0.5 and dp < 0.5, (dc, dp)
            rec["cpu_vs_gpu_max_abs_diff_px"] = {"centre": dc, "corners": dp}
        out[name] = rec
    g = cv2.imread(REF + "grayimage.jpg", cv2.IMREAD_UNCHANGED)
    cpu, nq = po.classic_detect(po.make_config(1280, 800, "gray", 2, 0.0), g)
    gpu = po.detect(po.make_config(1280, 800, "gray", 2, 0.0), g).detections
    assert [int(x["id"]) for x in cpu] == [int(x["id"]) for x in gpu] == [585]
    out["grayimage.jpg"] = {"width": 1280, "height": 800, "classic": dets_json(cpu), "classic_quads": nq,
                            "gpu_semantics": dets_json(gpu),
                            "cpu_vs_gpu_max_abs_diff_px": {"centre": float(np.abs(cpu[0]["c"] - gpu[0]["c"]).max()),
                                                           "corners": float(np.abs(cpu[0]["p"] - gpu[0]["p"]).max())}}
    with open(os.path.join(HERE, "classic_full_frames.json"), "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out.items():
        print(k, v.get("cpu_vs_gpu_max_abs_diff_px"), [d["id"] for d in v.get("classic_on_bgr2gray", v.get("classic", []))])


if __name__ == "__main__":
    main()
