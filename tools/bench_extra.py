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
    w, h, nf = 3840, 2160, 8
    depth, bgr = synth.rgbd_stream(nf, w, h, seed=7, device="cuda", distinct=8)
    out = torch.empty_like(depth)
    rows = []
    ffma_tf = None
    for r in range(3, 16):
        f = JointBilateralFilter(w, h, window_radius=r, max_batch=nf)
        g4 = f.presmooth(bgr)
        ms_filter = ev_time(lambda: f.filter_guide4(depth, g4, out), 5)
        ms_total = ev_time(lambda: f.process_batch(depth, bgr, out), 5)
        px = nf * w * h
        taps = (2 * r + 1) ** 2
        flop = 33 * taps + 4
        rows.append({"radius": r, "window": 2 * r + 1, "filter_ms": ms_filter, "process_ms": ms_total,
                     "filter_mpixel_s": px / ms_filter / 1e3, "process_mpixel_s": px / ms_total / 1e3,
                     "filter_tflops_algorithmic": px * flop / ms_filter / 1e9,
                     "filter_gtaps_s": px * taps / ms_filter / 1e6,
                     "hbm_gbs_algorithmic": px * 11 / ms_filter / 1e6,
                     "hbm_frac": px * 11 / ms_filter / 1e6 / PEAKS["hbm_gbs"]})
        f.close()
    print(json.dumps({"workload": "configs[3]: 3840x2160 synthetic RGB-D, 8 distinct frames per launch (265 MB of inputs > L2), "
                                  "r=3..15, sigmas 70/50/20", "rows": rows}))


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
    wl, hl, wh, hh, r = 512, 424, 1920, 1080, 7
    lo, _ = synth.rgbd_frame(wl, hl, seed=6, frame=0, noise_rel=0.01, device="cuda")
    _, hi = synth.rgbd_frame(wh, hh, seed=6, frame=0, device="cuda")
    f = JointBilateralFilter(wh, hh, window_radius=r)
    out = torch.empty((hh, wh), device="cuda")
    ms = ev_time(lambda: f.Upsampling(lo, hi, out), 20)
    byt = wh * hh * (3 + 4) + wl * hl * 4
    print(json.dumps({"workload": "configs[2]: 512x424 ToF depth -> 1920x1080 guide, r=7, pre-smooth + gather-form fill",
                      "ms": ms, "mpixel_s_out": wh * hh / ms / 1e3, "algorithmic_bytes": byt,
                      "hbm_gbs_algorithmic": byt / ms / 1e6, "filled_fraction": float((out > 0).float().mean())}))


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
    w, h, r = 640, 480, 7
    d, c = synth.rgbd_frame(w, h, seed=1, frame=0, device="cuda")
    f = JointBilateralFilter(w, h, window_radius=r)
    ms = ev_time(lambda: f.Process(d, c), 200, warm=20)
    f2 = JointBilateralFilter(w, h, window_radius=2)
    ms2 = ev_time(lambda: f2.Process(d, c), 200, warm=20)
    print(json.dumps({"workload": "one 640x480 frame per Process() call (launch latency included, L2-warm)",
                      "r7_us": ms * 1e3, "r7_mpixel_s": w * h / ms / 1e3, "r2_us": ms2 * 1e3, "r2_mpixel_s": w * h / ms2 / 1e3}))


def band():
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(os.environ.get("KDME_BAND_SIZE", "16384"))
    w = h = n
    r = 9
    peer = os.environ.get("KDME_BAND_PEER", "0") == "1"
    rb = shard.RowBandJBF(w, h, r, rank, world, device=local, peer_memory=peer)
    p = rb.plan
    # generate the band in slabs (position-keyed generator: any band on any rank)
    for y in range(p.y0, p.y1, 512):
        rows = min(512, p.y1 - y)
        d, c = synth.rgbd_frame(w, h, seed=16384, frame=0, y0=y, rows=rows, device=f"cuda:{local}")
        rb.depth_band[y - p.y0:y - p.y0 + rows].copy_(d)
        rb.bgr_band[y - p.y0:y - p.y0 + rows].copy_(c)
    torch.cuda.synchronize()

    def step():
        rb.process(exchange=True)
    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 3
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
    # exchange alone
    x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x0.record()
    for _ in range(10):
        rb.halo.barrier() if rb.peer_memory else rb.halo.exchange()
    x1.record()
    torch.cuda.synchronize()
    xms = torch.tensor([x0.elapsed_time(x1) / 10], device="cuda", dtype=torch.float64)
    checksum = rb.out.double().sum()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(xms, op=dist.ReduceOp.MAX)
        dist.all_reduce(checksum)
    if rank == 0:
        print(json.dumps({"workload": f"configs[4]: {w}x{h} synthetic RGB-D mosaic, r=9, row bands over {world} GPU(s), "
                                      + ("halo rows read from peer memory inside the kernels (no exchange)" if rb.peer_memory else
                                         "halo exchange (r+2 rows of depth + BGR per direction) inside the timed step"),
                          "halo_mode": "peer_memory" if rb.peer_memory else "nccl_send_recv",
                          "n_gpus": world, "ms_per_frame": float(ms.item()), "mpixel_s": w * h / float(ms.item()) / 1e3,
                          "halo_exchange_ms": float(xms.item()), "halo_bytes_per_direction": p.halo_bytes_per_direction(),
                          "checksum": float(checksum.item()), "timing": "CUDA events, max over ranks"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    {"sweep": sweep, "buffer2d": buffer2d, "upsample": upsample, "single": single, "band": band, "guided": guided}[sys.argv[1]]()
