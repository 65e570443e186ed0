"""Seeded synthetic RGB-D generator (stands in for the Kinect / ToF sensor, Kinect/Kinect.cpp).

SURVEY.md 8(d): piecewise-planar scene (3-6 planes, z in [500, 4500] mm) with step
edges; depth noise uniform in +-sigma(z) with the reference's own recipe
sigma(z) = 0.45*2.85*(z/10)^2/10000 mm (commented generator, main.cpp:127-130);
holes (zero depth) on ~8 % of pixels: blobs plus 2-px bands along depth edges, and
a sprinkle of values in (0, 50] that the filter must treat as holes
(JointBilateralFilter.cu:21); guide = per-plane albedo + texture noise, with
edges offset 1-3 px from the depth edges.

Everything is a pure function of (seed, frame, absolute pixel position): a
position-keyed integer hash, so any row band of any frame can be produced on any
device (CPU oracle or any GPU rank) and agrees bit-for-bit -- only integer ops
and IEEE +,-,*,/ are used.
"""
from __future__ import annotations

import torch

_M32 = 0xFFFFFFFF


def _hash32(x: torch.Tensor) -> torch.Tensor:
    """lowbias32-style avalanche on int64 tensors holding 32-bit values."""
    x = x & _M32
    x = ((x ^ (x >> 16)) * 0x7FEB352D) & _M32
    x = ((x ^ (x >> 15)) * 0x846CA68B) & _M32
    return x ^ (x >> 16)


def _hash_scalar(*keys: int) -> int:
    h = 0x9E3779B9
    for k in keys:
        h = (h ^ (k & _M32)) & _M32
        h = ((h ^ (h >> 16)) * 0x7FEB352D) & _M32
        h = ((h ^ (h >> 15)) * 0x846CA68B) & _M32
        h = h ^ (h >> 16)
    return h


def _unit(h: torch.Tensor) -> torch.Tensor:
    """uint32 hash -> fp32 in [0, 1) using 24 bits (exact in fp32)."""
    return (h >> 8).to(torch.float32) * (1.0 / 16777216.0)


def rgbd_frame(width: int, height: int, seed: int = 1234, frame: int = 0, *, y0: int = 0,
               rows: int | None = None, device="cpu", hole_frac: float = 0.08,
               depth_scale: float = 1.0, noise_rel: float | None = None):
    """Rows [y0, y0+rows) of one synthetic frame.

    Returns (depth f32 [rows, W] in mm, bgr u8 [rows, W, 3]).
    noise_rel: if given, noise amplitude = noise_rel * z (ToF-like, ~1 % z) instead of the
    Kinect quadratic recipe.
    """
    rows = height - y0 if rows is None else rows
    dev = torch.device(device)
    ys = torch.arange(y0, y0 + rows, device=dev, dtype=torch.int64).view(-1, 1)
    xs = torch.arange(0, width, device=dev, dtype=torch.int64).view(1, -1)
    fkey = _hash_scalar(seed, frame)
    n_planes = 3 + _hash_scalar(fkey, 1) % 4
    span = max(width, height)

    # Voronoi cells in a normalised metric -> straight step edges
    best_d = None
    region = torch.zeros((rows, width), device=dev, dtype=torch.int64)
    region_g = torch.zeros((rows, width), device=dev, dtype=torch.int64)
    jx = 1 + _hash_scalar(fkey, 2) % 3
    jy = 1 + _hash_scalar(fkey, 3) % 3
    best_g = None
    planes = []
    for k in range(n_planes):
        cx = _hash_scalar(fkey, 10, k) % width
        cy = _hash_scalar(fkey, 11, k) % height
        z0 = 500.0 + (_hash_scalar(fkey, 12, k) % 4000)
        ax = ((_hash_scalar(fkey, 13, k) % 2001) - 1000) * (600.0 / 1000.0) / span
        ay = ((_hash_scalar(fkey, 14, k) % 2001) - 1000) * (600.0 / 1000.0) / span
        alb = [40 + _hash_scalar(fkey, 15 + c, k) % 176 for c in range(3)]
        planes.append((cx, cy, z0, ax, ay, alb))
        d2 = (xs - cx) ** 2 + (ys - cy) ** 2
        d2g = (xs + jx - cx) ** 2 + (ys + jy - cy) ** 2
        if best_d is None:
            best_d, best_g = d2, d2g
        else:
            upd = d2 < best_d
            region = torch.where(upd, torch.full_like(region, k), region)
            best_d = torch.minimum(best_d, d2)
            updg = d2g < best_g
            region_g = torch.where(updg, torch.full_like(region_g, k), region_g)
            best_g = torch.minimum(best_g, d2g)

    xf = xs.to(torch.float32)
    yf = ys.to(torch.float32)
    z = torch.zeros((rows, width), device=dev, dtype=torch.float32)
    bgr = torch.zeros((rows, width, 3), device=dev, dtype=torch.float32)
    for k, (cx, cy, z0, ax, ay, alb) in enumerate(planes):
        zk = z0 + ax * (xf - cx) + ay * (yf - cy)
        z = torch.where(region == k, zk.expand(rows, width), z)
        for c in range(3):
            bgr[..., c] = torch.where(region_g == k, torch.full_like(z, float(alb[c])), bgr[..., c])
    z = torch.clamp(z, 400.0, 6000.0) * depth_scale

    pix = (ys * width + xs)  # absolute pixel key
    h_noise = _hash32(pix ^ _hash_scalar(fkey, 100))
    if noise_rel is None:
        amp = 0.45 * 2.85 * (z / 10.0) * (z / 10.0) / 10000.0
    else:
        amp = noise_rel * z
    z = z + (2.0 * _unit(h_noise) - 1.0) * amp

    # texture noise on the guide (+-12 levels), per channel
    for c in range(3):
        t = _unit(_hash32(pix ^ _hash_scalar(fkey, 200 + c)))
        bgr[..., c] = bgr[..., c] + (t * 24.0 - 12.0)
    bgr8 = torch.clamp(torch.floor(bgr + 0.5), 0, 255).to(torch.uint8)

    # holes: 8x8-cell blobs, 2-px bands along depth edges, sparse sub-threshold values
    cell = (ys // 8) * ((width + 7) // 8) + (xs // 8)
    blob = _unit(_hash32(cell ^ _hash_scalar(fkey, 300))) < (hole_frac * 0.6)
    speck = _unit(_hash32(pix ^ _hash_scalar(fkey, 301))) < (hole_frac * 0.25)

    def _region_at(dx, dy):
        # nearest-seed id at (x+dx, y+dy) recomputed analytically (band-independent)
        bd, rg = None, None
        for k, (cx, cy, *_r) in enumerate(planes):
            d2 = (xs + dx - cx) ** 2 + (ys + dy - cy) ** 2
            if bd is None:
                bd, rg = d2, torch.zeros_like(region)
            else:
                rg = torch.where(d2 < bd, torch.full_like(rg, k), rg)
                bd = torch.minimum(bd, d2)
        return rg

    edge = torch.zeros_like(blob)
    for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
        edge |= _region_at(dx, dy) != region
    edge &= _unit(_hash32(pix ^ _hash_scalar(fkey, 302))) < 0.7
    hole = blob | speck | edge
    low = _unit(_hash32(pix ^ _hash_scalar(fkey, 303))) < 0.004
    lowval = torch.floor(_unit(_hash32(pix ^ _hash_scalar(fkey, 304))) * 51.0)  # 0..50 inclusive
    depth = torch.where(hole, torch.zeros_like(z), z)
    depth = torch.where(low, lowval, depth)
    return depth.contiguous(), bgr8.contiguous()


def rgbd_stream(n_frames: int, width: int, height: int, seed: int = 1234, first_frame: int = 0,
                device="cpu", distinct: int | None = None, **kw):
    """Frames [first_frame, first_frame+n_frames): depth [N,H,W] f32, bgr [N,H,W,3] u8.

    distinct: generate only this many distinct frames and tile them (fast fill for
    throughput benches; every frame of the stream is still a full-size independent unit).
    """
    dev = torch.device(device)
    depth = torch.empty((n_frames, height, width), device=dev, dtype=torch.float32)
    bgr = torch.empty((n_frames, height, width, 3), device=dev, dtype=torch.uint8)
    gen = n_frames if distinct is None else min(distinct, n_frames)
    for i in range(gen):
        d, c = rgbd_frame(width, height, seed, first_frame + i, device=dev, **kw)
        depth[i], bgr[i] = d, c
    for i in range(gen, n_frames):
        depth[i], bgr[i] = depth[i % gen], bgr[i % gen]
    return depth, bgr
