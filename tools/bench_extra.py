#!/usr/bin/env python3
"""Secondary measurements (not the driver's headline line; see bench.py for that):

  python tools/bench_extra.py sweep      # configs[3]: radius sweep r=3..15 on 3840x2160 frames
  python tools/bench_extra.py buffer2d   # Buffer2D kernels vs the HBM roofline
  python tools/bench_extra.py upsample   # configs[2]: 512x424 ToF depth -> 1920x1080 guide
  python tools/bench_extra.py single     # one 640x480 frame per call (latency of Process)
  python tools/bench_extra.py guided     # guided cross-bilateral fill at 1920x1080
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
    res = []
    for (w, h, nf) in [(640, 480, 1), (3840, 2160, 1), (640, 480, 256)]:
        b = Buffer2D(w, h)
        frames = torch.rand((max(nf, 16), h, w), device="cuda") * 3000 + 500
        outp = torch.empty((h, w), device="cuda")
        px = w * h
        it = iter(range(10 ** 9))

        def upd():
            b.updateData(frames[next(it) % frames.shape[0]])
        ms_u = ev_time(upd, 50)
        ms_i = ev_time(lambda: b.insertData(frames[0]), 50)
        ms_g = ev_time(lambda: b.getDepthMap(outp), 50)
        row = {"size": f"{w}x{h}", "update_us": ms_u * 1e3, "insert_us": ms_i * 1e3, "get_depth_us": ms_g * 1e3,
               "update_gbs": px * 20 / ms_u / 1e6, "insert_gbs": px * 12 / ms_i / 1e6, "get_gbs": px * 12 / ms_g / 1e6}
        if nf > 1:
            ms_b = ev_time(lambda: b.updateData(frames[:nf]), 10)
            # fused N-frame update: buffer read+written once (16 B/px) + N input planes (4 B/px each)
            row.update({"fused_frames": nf, "fused_update_us": ms_b * 1e3,
                        "fused_gbs": px * (16 + 4 * nf) / ms_b / 1e6,
                        "fused_vs_per_frame_speedup": ms_u * nf / ms_b})
        row["hbm_frac_update"] = row["update_gbs"] / PEAKS["hbm_gbs"]
        res.append(row)
        b.close()
    print(json.dumps({"workload": "Buffer2D (ArrayBuffer/Buffer2D.cu) kernels; algorithmic bytes 20/12/12 B per pixel",
                      "hbm_peak_gbs": PEAKS["hbm_gbs"], "rows": res}))


def upsample():
    print(json.dumps(workloads.upsample()))


def guided():
    from kinectdepthmapenhancement_b200 import guided_fill
    w, h = 1920, 1080
    d, c = synth.rgbd_frame(w, h, seed=6, frame=1, device="cuda")
    lab = ((torch.arange(h, device="cuda")[:, None] // 40) * 64 + (torch.arange(w, device="cuda")[None, :] // 40)).int().contiguous()
    out = torch.empty_like(d)
    ms_l = ev_time(lambda: guided_fill(d, c, lab, 3, out=out), 20)
    ms_n = ev_time(lambda: guided_fill(d, c, None, 3, out=out), 20)
    print(json.dumps({"workload": "guided cross-bilateral fill (depthmap_enhancement) 1920x1080, window 7, sigmas 30/50/70",
                      "ms_with_labels": ms_l, "ms_no_labels": ms_n, "mpixel_s_with_labels": w * h / ms_l / 1e3,
                      "algorithmic_bytes": w * h * (4 + 3 + 4 + 4), "hbm_gbs_algorithmic": w * h * 15 / ms_l / 1e6}))


def single():
    print(json.dumps(workloads.single()))


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
    {"sweep": sweep, "buffer2d": buffer2d, "upsample": upsample, "single": single, "band": band, "guided": guided}[sys.argv[1]]()
