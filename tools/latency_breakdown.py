#!/usr/bin/env python3
"""Per-kernel device time of ONE frame (config 2 by default): where the single-frame latency goes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ros_vision_b200 import detector as D, synth
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
frame, fmt, w, h, dec, sigma, sc = synth.config_frame(cfg, 0)
det = D.GpuDetector(w, h, fmt, quad_decimate=dec, quad_sigma=sigma, max_batch=1)
dev = torch.from_numpy(np.ascontiguousarray(frame).reshape(-1)).cuda()
for _ in range(5):
    det.DetectDevice(dev.data_ptr(), 1)
prof = det.ProfileDevice(dev.data_ptr(), 1, iters=50)
tot = sum(ms for _, ms in prof)
for name, ms in prof:
    print(f"{name:16s} {ms * 1e3:8.1f} us")
print(f"{'sum':16s} {tot * 1e3:8.1f} us")
pb = D.PinnedBuffer(frame.size)
pb.array[:] = np.ascontiguousarray(frame).reshape(-1)
lat = []
for i in range(300):
    t0 = time.perf_counter(); det.DetectPointers([pb.ptr]); lat.append((time.perf_counter() - t0) * 1e6)
lat = np.sort(lat[20:])
print(f"host frame in -> detections out: p50 {lat[len(lat)//2]:.1f} us, p99 {lat[int(len(lat)*0.99)]:.1f} us")
lat = []
for i in range(300):
    t0 = time.perf_counter(); det.DetectDevice(dev.data_ptr(), 1); lat.append((time.perf_counter() - t0) * 1e6)
lat = np.sort(lat[20:])
print(f"device frame -> detections out: p50 {lat[len(lat)//2]:.1f} us")
