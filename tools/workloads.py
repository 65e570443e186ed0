#!/usr/bin/env python3
"""The secondary workloads of BASELINE.json (configs[2..4]) and the single-frame call, as functions that
return one dict each.  bench.py embeds them in the driver-run JSON line (`extra`), tools/bench_extra.py
prints them one at a time, tests/test_shard.py runs the band workload on real ranks.

Timing: CUDA events on the current stream after warm-up; max over ranks where ranks exist.
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from kinectdepthmapenhancement_b200 import JointBilateralFilter, shard, synth  # noqa: E402

FP32_NOMINAL_TFLOPS = 74.4   # 148 SM x 128 FMA/clk x 2 x 1.965 GHz


def flop_per_pixel(radius: int) -> int:
    """33 flop + 3 exp per tap per pixel (reference kernel text) + 4 flop/pixel epilogue."""
    return 33 * (2 * radius + 1) ** 2 + 4


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


# ------------------------------------------------------------------ single frame (the drop-in's real call)
def single(iters: int = 200):
    """One JointBilateralFilter::Process per 640x480 frame (main.cpp:178-180), launch latency included."""
    w, h = 640, 480
    d, c = synth.rgbd_frame(w, h, seed=1, frame=0, device="cuda")
    res = {"workload": "one 640x480 frame per Process() call (pre-smooth + filter + fp64 refinement launches, "
                       "launch latency included, inputs L2-warm)"}
    for r in (7, 2):
        f = JointBilateralFilter(w, h, window_radius=r)
        ms = ev_time(lambda: f.Process(d, c), iters, warm=20)
        res[f"r{r}_us"] = ms * 1e3
        res[f"r{r}_mpixel_s"] = w * h / ms / 1e3
        res[f"r{r}_frac_of_fp32_nominal"] = w * h * flop_per_pixel(r) / (ms * 1e-3) / 1e12 / FP32_NOMINAL_TFLOPS
        f.refine_stats()
        f.Process(d, c)
        res[f"r{r}_refined_px_per_frame"] = f.refine_stats()[0]
        f.close()
    xyz = torch.empty((h, w, 3), device="cuda")
    f = JointBilateralFilter(w, h, window_radius=7)
    ms = ev_time(lambda: f.process_xyz(d, c, 525.0, 525.0, w // 2, h // 2, out=xyz), iters, warm=20)
    res["r7_fused_xyz_us"] = ms * 1e3
    f.close()
    return res


# ------------------------------------------------------------------ configs[2]: upsampling
def upsample(iters: int = 20):
    wl, hl, wh, hh, r = 512, 424, 1920, 1080, 7
    lo, _ = synth.rgbd_frame(wl, hl, seed=6, frame=0, noise_rel=0.01, device="cuda")
    _, hi = synth.rgbd_frame(wh, hh, seed=6, frame=0, device="cuda")
    f = JointBilateralFilter(wh, hh, window_radius=r)
    out = torch.empty((hh, wh), device="cuda")
    ms = ev_time(lambda: f.Upsampling(lo, hi, out), iters)
    g4 = f.presmooth(hi[None])
    ms_pre = ev_time(lambda: f.presmooth(hi[None]), iters)
    byt = wh * hh * (3 + 4) + wl * hl * 4
    res = {"workload": "configs[2]: 512x424 ToF depth -> 1920x1080 guide, r=7, pre-smooth + fill",
           "ms": ms, "presmooth_ms": ms_pre, "mpixel_s_out": wh * hh / ms / 1e3, "algorithmic_bytes": byt,
           "hbm_gbs_algorithmic": byt / ms / 1e6, "filled_fraction": float((out > 0).float().mean())}
    del g4
    f.close()
    return res


# ------------------------------------------------------------------ a5: guided cross-bilateral fill
def guided(iters: int = 20):
    """depthmap_enhancement (EdgeRefinedSuperpixel.cu:104-205) at 1920x1080, window 7, the reference's sigmas."""
    from kinectdepthmapenhancement_b200 import guided_fill
    w, h = 1920, 1080
    d, c = synth.rgbd_frame(w, h, seed=6, frame=1, device="cuda")
    lab = ((torch.arange(h, device="cuda")[:, None] // 40) * 64 + (torch.arange(w, device="cuda")[None, :] // 40)).int().contiguous()
    out = torch.empty_like(d)
    ms_l = ev_time(lambda: guided_fill(d, c, lab, 3, out=out), iters)
    ms_n = ev_time(lambda: guided_fill(d, c, None, 3, out=out), iters)
    return {"workload": "guided cross-bilateral fill (depthmap_enhancement) 1920x1080, window 7, sigmas 30/50/70",
            "ms_with_labels": ms_l, "ms_no_labels": ms_n, "mpixel_s_with_labels": w * h / ms_l / 1e3,
            "algorithmic_bytes": w * h * (4 + 3 + 4 + 4), "hbm_gbs_algorithmic": w * h * 15 / ms_l / 1e6}


# ------------------------------------------------------------------ a7/a8: Buffer2D against the HBM roofline
def buffer2d(hbm_gbs: float = 6551.0):
    from kinectdepthmapenhancement_b200 import Buffer2D
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
        row["hbm_frac_update"] = row["update_gbs"] / hbm_gbs
        res.append(row)
        b.close()
        del frames, outp
    return {"workload": "Buffer2D (ArrayBuffer/Buffer2D.cu) kernels; algorithmic bytes 20/12/12 B per pixel",
            "hbm_peak_gbs": hbm_gbs, "rows": res}


# ------------------------------------------------------------------ the "next" rows f1-f4 at Kinect size
def next_rows(iters: int = 50):
    """MarkovRandomField (f1), Projection_GPU::bilateralfilter on clouds (f2), projectiveToReal alone and fused into the
    filter epilogue (f3), the mean 3-D error reduction (f4) on one 640x480 frame: microseconds per call."""
    from kinectdepthmapenhancement_b200 import evalio
    from kinectdepthmapenhancement_b200.jbf import projective_to_real
    w, h = 640, 480
    d, c = synth.rgbd_frame(w, h, seed=1, frame=0, device="cuda")
    f = JointBilateralFilter(w, h, window_radius=2)
    pts = projective_to_real(d, 525.0, 525.0, w // 2, h // 2)
    z = torch.where(pts[..., 2] > 0, pts[..., 2], torch.ones_like(pts[..., 2]))
    norm = pts.clone()
    norm[..., 0] /= z
    norm[..., 1] /= z
    norm = norm.contiguous()
    xyz = torch.empty((h, w, 3), device="cuda")
    res = {"workload": "next rows on one 640x480 frame (window 5): microseconds per call, launch latency included",
           "f1_mrf_us": ev_time(lambda: f.mrf(d, c), iters) * 1e3,
           "f2_depth_bilateral_xyz_us": ev_time(lambda: evalio.depth_bilateral_xyz(norm, pts), iters) * 1e3,
           "f3_projective_to_real_us": ev_time(lambda: projective_to_real(d, 525.0, 525.0, w // 2, h // 2), iters) * 1e3,
           "f3_process_then_project_us": ev_time(lambda: (f.Process(d, c), projective_to_real(f.getFiltered_Device(), 525.0, 525.0, w // 2, h // 2)), iters) * 1e3,
           "f3_process_xyz_fused_us": ev_time(lambda: f.process_xyz(d, c, 525.0, 525.0, w // 2, h // 2, out=xyz), iters) * 1e3,
           "f4_mean_3d_error_us": ev_time(lambda: evalio.mean_3d_error(xyz, pts), iters) * 1e3}
    f.close()
    return res


# ------------------------------------------------------------------ configs[3]: radius sweep
def sweep(radii=range(3, 16), hbm_gbs: float = 6551.0, iters: int = 5):
    w, h, nf = 3840, 2160, 8
    depth, bgr = synth.rgbd_stream(nf, w, h, seed=7, device="cuda", distinct=8)
    out = torch.empty_like(depth)
    rows = []
    for r in radii:
        f = JointBilateralFilter(w, h, window_radius=r, max_batch=nf)
        g4 = f.presmooth(bgr)
        ms_filter = ev_time(lambda: f.filter_guide4(depth, g4, out), iters)
        ms_total = ev_time(lambda: f.process_batch(depth, bgr, out), iters)
        px = nf * w * h
        taps = (2 * r + 1) ** 2
        rows.append({"radius": r, "window": 2 * r + 1, "filter_ms": ms_filter, "process_ms": ms_total,
                     "filter_mpixel_s": px / ms_filter / 1e3, "process_mpixel_s": px / ms_total / 1e3,
                     "filter_tflops_algorithmic": px * flop_per_pixel(r) / ms_filter / 1e9,
                     "frac_of_fp32_nominal": px * flop_per_pixel(r) / ms_filter / 1e9 / FP32_NOMINAL_TFLOPS,
                     "filter_gtaps_s": px * taps / ms_filter / 1e6,
                     "hbm_gbs_algorithmic": px * 11 / ms_filter / 1e6,
                     "hbm_frac": px * 11 / ms_filter / 1e6 / hbm_gbs})
        f.close()
        del g4
    return {"workload": "configs[3]: 3840x2160 synthetic RGB-D, 8 distinct frames per launch (265 MB of inputs > L2), "
                        "sigmas 70/50/20; filter = two-pass kernel + fp64 refinement, process = + pre-smooth", "rows": rows}


# ------------------------------------------------------------------ configs[4]: row bands
def _digest(t: torch.Tensor) -> int:
    """Order-sensitive 62-bit digest of a float32 tensor's bit patterns (position-weighted sum mod 2^62)."""
    v = t.contiguous().view(torch.int32).to(torch.int64).flatten() & 0xFFFFFFFF
    idx = torch.arange(v.numel(), device=v.device, dtype=torch.int64)
    mult = ((idx * 2654435761) & 0x3FFFFFFF) | 1
    return int(((v * mult) & 0x3FFFFFFFFFFFFFFF).sum().item() & 0x3FFFFFFFFFFFFFFF)


def _strip_single_gpu(w, h, r, seed, y_a, y_b, device):
    """Rows [y_a, y_b) of the frame filtered WITHOUT any band logic: the rows and their halo are generated
    locally from the position-keyed generator and go through the ordinary single-GPU row path."""
    from kinectdepthmapenhancement_b200 import _lib
    halo = r + shard.PRESMOOTH_RADIUS
    s0, s1 = max(0, y_a - halo), min(h, y_b + halo)
    d, c = synth.rgbd_frame(w, h, seed=seed, frame=0, y0=s0, rows=s1 - s0, device=device)
    f = JointBilateralFilter(w, s1 - s0, window_radius=r, device=torch.device(device).index)
    pitch = (w + 3) & ~3
    g4 = torch.empty((s1 - s0, pitch), dtype=torch.int32, device=device)
    out = torch.empty((y_b - y_a, w), dtype=torch.float32, device=device)
    L = _lib.lib()
    _lib.check(L.jbf_presmooth_rows(f._h, c.data_ptr(), 3 * w, g4.data_ptr(), pitch * 4, s1 - s0))
    _lib.check(L.jbf_filter_rows(f._h, d.data_ptr(), g4.data_ptr(), pitch * 4, out.data_ptr(), s1 - s0, y_a - s0, y_b - y_a))
    torch.cuda.synchronize()
    f.close()
    return out


def _oracle_rows(w, h, r, seed, y_a, y_b, x0, ncol):
    """fp64 oracle of rows [y_a, y_b), columns [x0, x0 + ncol): (o64, guard-active mask)."""
    import oracle
    halo = r + shard.PRESMOOTH_RADIUS
    s0, s1 = max(0, y_a - halo), min(h, y_b + halo)
    c0, c1 = max(0, x0 - halo), min(w, x0 + ncol + halo)
    d, c = synth.rgbd_frame(w, h, seed=seed, frame=0, y0=s0, rows=s1 - s0)
    d, c = d[:, c0:c1].contiguous().numpy(), c[:, c0:c1].contiguous().numpy()
    g = oracle.presmooth(c)       # reflect-contaminated only in the outermost 2 rows/cols, which the window never reaches
    o64, mean = oracle.jbf(d, g, 2 * r + 1, precision="f64", return_mean=True)
    act = oracle.guard_active_mask(d, mean, 2 * r + 1, oracle.JBF_SIGMA_D)
    ys = slice(y_a - s0, y_b - s0)
    xs = slice(x0 - c0, x0 - c0 + ncol)
    return o64[ys, xs], act[ys, xs]


def bands(size: int = 16384, radius: int = 9, peer: bool = False, steps: int = 3, verify: bool = True,
          oracle_cols: int = 1024, seed: int = 16384):
    """configs[4]: one size x size frame split into row bands over the ranks of the default process group
    (1 rank: the whole frame), halo rows exchanged over NVLink (NCCL send/recv) or read from peer memory
    inside the kernels.  Verification: (1) every rank re-filters the 32 rows on each side of its seams
    WITHOUT band logic and compares bit for bit, (2) a 62-bit digest of every band is gathered, (3) rank 0
    checks the two rows at every seam (oracle_cols columns) against the fp64 CPU oracle."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    dev = torch.device("cuda", torch.cuda.current_device())
    w = h = size
    rb = shard.RowBandJBF(w, h, radius, rank, world, device=dev.index, peer_memory=peer)
    p = rb.plan
    for y in range(p.y0, p.y1, 512):   # position-keyed generator: any band on any rank
        rows = min(512, p.y1 - y)
        d, c = synth.rgbd_frame(w, h, seed=seed, frame=0, y0=y, rows=rows, device=dev)
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
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    x0e, x1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x0e.record()
    for _ in range(10):
        rb.halo.barrier() if rb.peer_memory else rb.halo.exchange()
    x1e.record()
    torch.cuda.synchronize()
    xms = torch.tensor([x0e.elapsed_time(x1e) / 10], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(xms, op=dist.ReduceOp.MAX)
    ms_f = float(ms.item())
    res = {"workload": f"configs[4]: {w}x{h} synthetic RGB-D mosaic, r={radius}, row bands over {world} GPU(s), "
                       + ("halo rows read from peer memory inside the kernels (no exchange)" if rb.peer_memory else
                          "halo exchange (r+2 rows of depth + BGR per direction, NCCL send/recv) inside the timed step"),
           "halo_mode": "peer_memory" if rb.peer_memory else "nccl_send_recv", "n_gpus": world,
           "ms_per_frame": ms_f, "mpixel_s": w * h / ms_f / 1e3,
           "tflops_algorithmic": w * h * flop_per_pixel(radius) / ms_f / 1e9,
           "frac_of_fp32_nominal": w * h * flop_per_pixel(radius) / ms_f / 1e9 / (FP32_NOMINAL_TFLOPS * world),
           "halo_exchange_ms": float(xms.item()), "halo_bytes_per_direction": p.halo_bytes_per_direction(),
           "band_rows": [b - a for a, b in shard.band_partition(h, world)], "timing": "CUDA events, max over ranks"}
    if not verify:
        return res
    # (1) seam rows re-filtered without band logic, bit for bit
    edge = 32
    ok = True
    for (a, b) in ((p.y0, min(p.y0 + edge, p.y1)), (max(p.y1 - edge, p.y0), p.y1)):
        ref = _strip_single_gpu(w, h, radius, seed, a, b, dev)
        ok = ok and bool(torch.equal(ref.view(torch.int32), rb.out[a - p.y0:b - p.y0].view(torch.int32)))
    okt = torch.tensor([1 if ok else 0], device=dev, dtype=torch.int32)
    # (2) digests of all bands, (3) the two rows at every seam for the oracle
    dg = torch.tensor([_digest(rb.out)], device=dev, dtype=torch.int64)
    ncol = min(oracle_cols, w)
    seam_cols = [((k + 1) * 2654435761 % max(1, w - ncol)) & ~3 for k in range(max(world - 1, 1))]
    first = torch.zeros((max(world - 1, 1), ncol), device=dev)
    last = torch.zeros((max(world - 1, 1), ncol), device=dev)
    if world > 1:
        if rank > 0:
            first[rank - 1] = rb.out[0, seam_cols[rank - 1]:seam_cols[rank - 1] + ncol]
        if rank < world - 1:
            last[rank] = rb.out[-1, seam_cols[rank]:seam_cols[rank] + ncol]
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        dgs = [torch.zeros_like(dg) for _ in range(world)]
        dist.all_gather(dgs, dg)
        dist.all_reduce(first)   # disjoint rows: the sum assembles them
        dist.all_reduce(last)
        seams = [b for (_, b) in shard.band_partition(h, world)][:-1]
    else:
        dgs = [dg]
        # one GPU: check the rows where the 8-way split would put its seams (same code path, whole frame)
        seams = [b for (_, b) in shard.band_partition(h, min(8, max(2, h // 64)))][:-1][:1]
        first[0] = rb.out[seams[0], seam_cols[0]:seam_cols[0] + ncol]
        last[0] = rb.out[seams[0] - 1, seam_cols[0]:seam_cols[0] + ncol]
    res["seam_rows_bitwise_equal_single_gpu_path"] = bool(okt.item())
    res["band_digests"] = [int(x.item()) for x in dgs]
    if rank == 0:
        import numpy as np
        worst_reg, worst_act, n_act, mism = 0.0, 0.0, 0, 0
        for k, ys in enumerate(seams):
            o64, act = _oracle_rows(w, h, radius, seed, ys - 1, ys + 1, seam_cols[k], ncol)
            got = np.stack([last[k].cpu().numpy(), first[k].cpu().numpy()])
            err = np.abs(got.astype(np.float64) - o64)
            mism += int(np.count_nonzero((got > 0) != (o64 > 0)))
            worst_reg = max(worst_reg, float(np.where(act, 0, err).max()))
            worst_act = max(worst_act, float(np.where(act, err, 0).max()))
            n_act += int(act.sum())
        res["oracle_seam_check"] = {"seams": len(seams), "rows_per_seam": 2, "cols": ncol, "mask_mismatches": mism,
                                    "max_abs_regular_mm": worst_reg, "max_abs_active_mm": worst_act, "n_active": n_act,
                                    "oracle": "fp64 CPU oracle on the position-keyed frame"}
    return res


# ------------------------------------------------------------------ configs[1], strong-scaling form
def strong(jbf, depth, bgr, out, total_frames: int, steps: int = 3):
    """4096 frames in total, frame-sharded over the ranks (shard.frame_shard); value = total pixels / max time."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    a, b = shard.frame_shard(total_frames, world, rank)
    n = b - a
    h, w = depth.shape[1:]

    def step():
        jbf.process_batch(depth[:n], bgr[:n], out[:n])
    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=depth.device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return {"scaling": "strong", "total_frames": total_frames, "frames_this_rank": n, "n_gpus": world,
            "ms_per_step": float(ms.item()), "value": total_frames * w * h / float(ms.item()) / 1e3, "unit": "Mpixel/s"}
