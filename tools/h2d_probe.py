#!/usr/bin/env python3
"""Bare host->device bandwidth of the box, N GPUs at once, no kernels.

    python tools/h2d_probe.py [--gpus 1,2,4,8] [--mb 256] [--iters 20] [--out profiles/h2d_probe_rNN.json]

Every worker owns one GPU: a pinned host block of --mb MB and a device buffer, and times `iters` back-to-back
cudaMemcpyAsync(host -> device) of the whole block with CUDA events on its own stream (one cudaMemcpyAsync per block,
exactly what b200tag_enqueue_host_block issues).  All workers start together (barrier); the table gives the per-GPU and
the aggregate GB/s for
  procs    one process per GPU (how bench.py / torchrun runs),
  threads  one process, one thread per GPU,
  procs+2M one process per GPU, the host block 2 MB aligned and registered with cudaHostRegister.
This is the ceiling of bench.py's raw-frame `e2e` number at N GPUs: frames/s <= aggregate GB/s / bytes per frame.
"""
from __future__ import annotations

import argparse
import json
import mmap
import os
import sys
import threading
import time

import numpy as np


def _copy_loop(torch, dev, mb, iters, aligned, start, results, idx):
    torch.cuda.set_device(dev)
    nbytes = mb << 20
    keep = None
    if aligned:
        # 2 MB aligned anonymous mapping, registered as pinned memory
        raw = mmap.mmap(-1, nbytes + (2 << 20))
        arr = np.frombuffer(raw, dtype=np.uint8)
        off = (-arr.ctypes.data) % (2 << 20)
        arr = arr[off:off + nbytes]
        arr[:] = 1
        rc = torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, nbytes, 0)
        if int(rc) != 0:
            results[idx] = {"error": f"cudaHostRegister rc={int(rc)}"}
            start.wait()
            return
        host = torch.from_numpy(arr)
        keep = raw
    else:
        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        host.fill_(1)
    dst = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{dev}")
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        for _ in range(3):
            dst.copy_(host, non_blocking=True)
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.wait()
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(iters):
            dst.copy_(host, non_blocking=True)
        e1.record(st)
    st.synchronize()
    ms = e0.elapsed_time(e1)
    results[idx] = {"gpu": dev, "gbs": nbytes * iters / (ms * 1e-3) / 1e9, "ms": ms}
    if aligned:
        torch.cuda.cudart().cudaHostUnregister(arr.ctypes.data)
    del keep


def _proc_main(dev, n, mb, iters, aligned, barrier, q):
    import torch

    class B:
        def wait(self):
            barrier.wait()
    res = [None]
    _copy_loop(torch, dev, mb, iters, aligned, B(), res, 0)
    q.put(res[0])


def run_procs(n, mb, iters, aligned):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(n)
    q = ctx.Queue()
    ps = [ctx.Process(target=_proc_main, args=(d, n, mb, iters, aligned, barrier, q)) for d in range(n)]
    for p in ps:
        p.start()
    out = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    return out


def run_threads(n, mb, iters):
    import torch
    start = threading.Barrier(n)
    res = [None] * n
    ts = [threading.Thread(target=_copy_loop, args=(torch, d, mb, iters, False, start, res, d)) for d in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    have = torch.cuda.device_count()
    rows = []
    for n in [int(x) for x in args.gpus.split(",")]:
        if n > have:
            continue
        for mode in ("procs", "threads", "procs+2M"):
            if mode == "procs+2M" and n not in (1, max(int(x) for x in args.gpus.split(",") if int(x) <= have)):
                continue
            t0 = time.time()
            try:
                if mode == "threads":
                    r = run_threads(n, args.mb, args.iters)
                else:
                    r = run_procs(n, args.mb, args.iters, mode.endswith("2M"))
                per = [x["gbs"] for x in r if x and "gbs" in x]
                rows.append({"n_gpus": n, "mode": mode, "per_gpu_gbs_min": min(per), "per_gpu_gbs_max": max(per),
                             "aggregate_gbs": sum(per), "errors": [x for x in r if x and "error" in x]})
            except Exception as e:  # noqa: BLE001
                rows.append({"n_gpus": n, "mode": mode, "error": repr(e)[:200]})
            print(json.dumps(rows[-1]), f"({time.time() - t0:.1f} s)", flush=True)
    cpu = ""
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                cpu = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    doc = {"what": "cudaMemcpyAsync pinned host -> device, all GPUs at once, no kernels", "block_mb": args.mb, "iters": args.iters,
           "host_cpu": cpu, "host_threads": os.cpu_count(), "gpus_visible": have, "rows": rows}
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(doc, f, indent=1)
    print(json.dumps(doc))


if __name__ == "__main__":
    sys.exit(main())
