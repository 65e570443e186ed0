#!/usr/bin/env python3
"""Broad parity sweep of the GPU filter against the fp64 CPU oracle (the checker; never the thing measured):
many seeds, sizes, radii and hole densities, with the flat north_star tolerance (|gpu - f64| <= 1e-3 mm on every
pixel where the reference's skip-if-zero guard is inactive; mask bit-exact).  Prints one JSON object.

    python tools/parity_sweep.py [--frames 24] [--fuzz 150]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import oracle  # noqa: E402
from kinectdepthmapenhancement_b200 import JointBilateralFilter, synth  # noqa: E402


def run_case(w, h, r, seed, frame, hole, sig=(70.0, 50.0, 20.0)):
    d, c = synth.rgbd_frame(w, h, seed, frame, hole_frac=hole)
    f = JointBilateralFilter(w, h, *sig, window_radius=r)
    f.Process(d.cuda(), c.cuda())
    out = f.getFiltered_Device().cpu().numpy()
    guide = f.getSmoothImage_Device().cpu().numpy()
    par = oracle.parity_block(out, d.numpy(), guide, 2 * r + 1, *sig)
    par.update({"w": w, "h": h, "radius": r, "seed": seed, "frame": frame, "holes": hole, "refined": f.refine_stats()[0],
                "presmooth_bit_exact": bool(np.array_equal(guide, oracle.presmooth(c.numpy())))})
    f.close()
    return par


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=24)
    ap.add_argument("--fuzz", type=int, default=150)
    args = ap.parse_args()
    oracle.build()
    rows = []
    for i in range(args.frames):     # full Kinect frames, the configs[1] generator and others
        r = (7, 2, 9, 15, 4, 7)[i % 6]
        rows.append(run_case(640, 480, r, 1234 + 17 * i, i, (0.08, 0.0, 0.3, 0.08)[i % 4]))
    rng = np.random.default_rng(2024)
    for i in range(args.fuzz):       # small random frames
        w, h = int(rng.integers(1, 200)), int(rng.integers(1, 150))
        rows.append(run_case(w, h, int(rng.integers(1, 16)), 5000 + i, i, float(rng.choice([0.0, 0.08, 0.5, 0.9]))))
    worst = max(rows, key=lambda q: q["max_abs_regular_mm"])
    res = {
        "cases": len(rows), "pixels": sum(q["pixels"] for q in rows),
        "mask_mismatches": sum(q["mask_mismatches"] for q in rows), "nan": sum(q["nan"] for q in rows),
        "presmooth_bit_exact": all(q["presmooth_bit_exact"] for q in rows),
        "max_abs_regular_mm": worst["max_abs_regular_mm"],
        "worst_regular_case": {k: worst[k] for k in ("w", "h", "radius", "seed", "frame", "holes", "worst_regular_yx")},
        "cases_with_regular_above_1e-3": sum(q["max_abs_regular_mm"] > 1e-3 for q in rows),
        "n_active": sum(q["n_active"] for q in rows), "n_active_beyond_1e-3": sum(q["n_active_beyond_1e-3"] for q in rows),
        "max_abs_active_mm": max(q["max_abs_active_mm"] for q in rows),
        "refined_fp64_pixels": sum(q["refined"] for q in rows),
        "histogram_regular_max_mm": {str(k): int(sum(abs(q["max_abs_regular_mm"] - k) < 1e-9 for q in rows))
                                     for k in sorted({q["max_abs_regular_mm"] for q in rows})},
    }
    print(json.dumps(res))


if __name__ == "__main__":
    main()
