#!/usr/bin/env python3
"""Secondary measurements (not the driver's headline line; see bench.py for that):

  python tools/bench_extra.py sweep      # configs[3]: radius sweep r=3..15 on 3840x2160 frames
  python tools/bench_extra.py buffer2d   # Buffer2D kernels vs the HBM roofline
  python tools/bench_extra.py upsample   # configs[2]: 512x424 ToF depth -> 1920x1080 guide
  python tools/bench_extra.py single     # one 640x480 frame per call (latency of Process)
  python tools/bench_extra.py guided     # guided cross-bilateral fill at 1920x1080
  python tools/bench_extra.py next_rows  # f1-f4 (MRF, cloud bilateral, projectiveToReal, mean 3-D error) per call
  torchrun ... tools/bench_extra.py band # configs[4]: 16384x16384 r=9 row bands + NVLink halo exchange

Each prints one JSON object; copies are kept under profiles/.
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from kinectdepthmapenhancement_b200 import Buffer2D, JointBilateralFilter, shard, synth  # noqa: E402
from tools import workloads  # noqa: E402

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}


def ev_time(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def flush_l2():
    if not hasattr(flush_l2, "buf"):
        flush_l2.buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flush_l2.buf.fill_(1)


def sweep():
    print(json.dumps(workloads.sweep(hbm_gbs=PEAKS["hbm_gbs"])))


def buffer2d():
    print(json.dumps(workloads.buffer2d(hbm_gbs=PEAKS["hbm_gbs"])))


def upsample():
    print(json.dumps(workloads.upsample()))


def guided():
    print(json.dumps(workloads.guided()))


def single():
    print(json.dumps(workloads.single()))


def next_rows():
    print(json.dumps(workloads.next_rows()))


def band():
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = workloads.bands(int(os.environ.get("KDME_BAND_SIZE", "16384")), 9, peer=os.environ.get("KDME_BAND_PEER", "0") == "1")
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    {"sweep": sweep, "buffer2d": buffer2d, "upsample": upsample, "single": single, "band": band, "guided": guided, "next_rows": next_rows}[sys.argv[1]]()
