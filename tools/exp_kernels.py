#!/usr/bin/env python3
"""Per-kernel CUDA-event times and device-resident throughput of config 2 for the kernel variants selected by
B200TAG_EXP (csrc/kernels.h: exp_flags) -- one process per variant, since the switch is read once.
usage (on the GPU box): B200TAG_EXP=<mask> python tools/exp_kernels.py [--lanes 2,3,4] [--batch 128] [--steps 10]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lanes", default="2")
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--config", type=int, default=2)
    a = ap.parse_args()
    import torch
    from ros_vision_b200 import detector as D
    D.load_library()
    torch.cuda.set_device(0)
    bench.select_config(a.config)
    T = bench.Timing(torch, None, 1, 0)
    frames = bench.make_frames()
    out = {"exp": os.environ.get("B200TAG_EXP", "0"), "config": a.config, "batch": a.batch, "value": {}}
    for dl in [int(x) for x in a.lanes.split(",")]:
        det, dev_batch, _, ms, _, _ = bench.device_resident_leg(T, D, 0, frames, a.batch, dl, a.steps, 3)
        out["value"][str(dl)] = round(a.batch * dl * a.steps / (ms * 1e-3))
        prof = det.ProfileDevice(dev_batch.data_ptr(), a.batch, iters=5)
        det.close()
        del dev_batch
    out["kernels_ms"] = {n: round(m, 4) for n, m in prof}
    out["sum_ms"] = round(sum(m for _, m in prof), 4)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
