#!/usr/bin/env python
"""End-to-end throughput of the camera wire format: config-2 scenes as 4:2:2 JPEG bitstreams
in host memory -> b200tag_enqueue_mjpg (luminance decode + detection on one stream) -> detections on the host.
Informational: the headline bench (bench.py) stays on raw YUYV frames.  Usage: python tools/bench_mjpg.py [--batch 128]
[--steps 20] [--lanes 3] [--quality 75] [--decoder native|sequential|nvjpeg]; with --decoder nvjpeg,
B200TAG_NVJPEG_BACKEND=hardware|gpu|hybrid|default selects the nvJPEG backend."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import cv2
    from ros_vision_b200 import detector as D, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--lanes", type=int, default=3)
    ap.add_argument("--quality", type=int, default=75)
    ap.add_argument("--unique", type=int, default=16)
    ap.add_argument("--decoder", default="native", choices=["native", "sequential", "nvjpeg"],
                    help="native: the engine's parallel decode kernels; sequential: its warp-per-frame kernel; nvjpeg: the library")
    a = ap.parse_args()
    if a.decoder != "native":
        os.environ["B200TAG_MJPG_DECODER"] = a.decoder
    D.load_library()
    w, h = 1280, 800
    jpgs, ntags = [], 0
    for i in range(a.unique):
        frame, fmt, _, _, dec, sigma, sc = synth.config_frame(2, i)
        bgr = synth.gray_to_bgr(sc.gray, np.random.default_rng(i))
        ok, buf = cv2.imencode(".jpg", bgr, [cv2.IMWRITE_JPEG_QUALITY, a.quality, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                              cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422])
        assert ok
        jpgs.append(buf.tobytes())
        ntags += len(sc.tags)
    batch = [jpgs[i % len(jpgs)] for i in range(a.batch)]
    dets = [D.GpuDetector(w, h, "gray", quad_decimate=2, max_batch=a.batch) for _ in range(a.lanes)]
    for it in range(a.warmup):
        for d in dets:
            d.EnqueueMjpg(batch)
        for d in dets:
            d.Finish()
    found = sum(len(dets[0].Detections(f)) for f in range(a.unique))
    t0 = time.perf_counter()
    pending = []
    for it in range(a.steps):
        d = dets[it % a.lanes]
        if len(pending) == a.lanes:
            pending.pop(0).Finish()
        d.EnqueueMjpg(batch)
        pending.append(d)
    for d in pending:
        d.Finish()
    dt = time.perf_counter() - t0
    print(json.dumps({"metric": "mjpg_frames_per_second_end_to_end", "value": a.batch * a.steps / dt, "unit": "frames/s",
                      "ms_per_step": 1e3 * dt / a.steps, "batch": a.batch, "lanes": a.lanes, "steps": a.steps,
                      "jpeg_bytes_per_frame": int(np.mean([len(j) for j in jpgs])), "quality": a.quality,
                      "decoder": a.decoder if a.decoder != "nvjpeg" else "nvjpeg/" + dets[0].mjpg_backend,
                      "parallel_frames_last_batch": dets[0].MjpgParallelFrames(), "tags_found": found, "tags_present": ntags}))
    for d in dets:
        d.close()


if __name__ == "__main__":
    main()
