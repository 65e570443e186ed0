#!/usr/bin/env python3
"""Bare pinned-memory copy ceiling of the host at N ranks.

Every rank does what jbf_process_host does to the host and nothing else: H2D copies of 7 B/pixel (depth f32 +
BGR u8) and D2H copies of 4 B/pixel (filtered f32) in 16-frame chunks on two streams, one cudaMemcpyAsync per
copy, no kernels.  The aggregate GB/s over all ranks is the ceiling the end-to-end Mpixel/s can reach on this
host: e2e_ceiling_mpixel_s = aggregate_bytes_per_s / 11.

    python tools/host_copy_ceiling.py                       # 1 rank
    torchrun --nproc-per-node 8 tools/host_copy_ceiling.py  # 8 ranks, one per GPU

Prints one JSON object (rank 0).  Variants: torch.pin_memory buffers, library-owned cudaHostAlloc buffers,
write-combined inputs, and the u16-depth byte mix (5 up + 4 down).
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

W, H = 640, 480


def run(kind: str, frames: int, reps: int, world: int, dev, depth_bytes: int = 4):
    from kinectdepthmapenhancement_b200.jbf import host_buffer
    px = frames * W * H
    ddt = torch.float32 if depth_bytes == 4 else torch.int16
    if kind == "torch_pinned":
        hd = torch.empty((px,), dtype=ddt).pin_memory()
        hc = torch.empty((px * 3,), dtype=torch.uint8).pin_memory()
        ho = torch.empty((px,), dtype=torch.float32).pin_memory()
    else:
        wc = kind == "write_combined"
        hd = host_buffer((px,), ddt, write_combined=wc)
        hc = host_buffer((px * 3,), torch.uint8, write_combined=wc)
        ho = host_buffer((px,), torch.float32, write_combined=False)
    hd.fill_(1)
    hc.fill_(2)
    dd = torch.empty((px,), dtype=ddt, device=dev)
    dc = torch.empty((px * 3,), dtype=torch.uint8, device=dev)
    do = torch.zeros((px,), dtype=torch.float32, device=dev)
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    chunk = 16 * W * H

    def once():
        for p0 in range(0, px, chunk):
            p1 = min(px, p0 + chunk)
            with torch.cuda.stream(s_up):
                dd[p0:p1].copy_(hd[p0:p1], non_blocking=True)
                dc[3 * p0:3 * p1].copy_(hc[3 * p0:3 * p1], non_blocking=True)
            with torch.cuda.stream(s_dn):
                ho[p0:p1].copy_(do[p0:p1], non_blocking=True)
        s_up.synchronize()
        s_dn.synchronize()

    once()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    up, dn = px * (depth_bytes + 3) * reps, px * 4 * reps
    return {"h2d_gbs_total": world * up / dt / 1e9, "d2h_gbs_total": world * dn / dt / 1e9,
            "gbs_total": world * (up + dn) / dt / 1e9,
            "e2e_ceiling_mpixel_s": world * px * reps / dt / 1e6, "seconds": dt}


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    frames = int(os.environ.get("KDME_CEIL_FRAMES", "1024"))
    reps = int(os.environ.get("KDME_CEIL_REPS", "4"))
    res = {"n_ranks": world, "frames_per_rank": frames, "reps": reps,
           "what": "bare cudaMemcpyAsync H2D (depth + BGR) and D2H (f32 result) in 16-frame chunks on two streams per rank, "
                   "no kernels; wall clock, max over ranks"}
    for kind in ("torch_pinned", "host_alloc", "write_combined"):
        res[kind + "_f32"] = run(kind, frames, reps, world, dev, 4)
    res["host_alloc_u16"] = run("host_alloc", frames, reps, world, dev, 2)
    res["write_combined_u16"] = run("write_combined", frames, reps, world, dev, 2)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
