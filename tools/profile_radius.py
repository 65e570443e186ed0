#!/usr/bin/env python3
"""A few launches of the filter kernel at one radius on 3840x2160 frames, for ncu (configs[3]).
usage: python tools/profile_radius.py R [width height frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from kinectdepthmapenhancement_b200 import JointBilateralFilter, synth  # noqa: E402

r = int(sys.argv[1])
w, h, nf = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (3840, 2160, 2)
depth, bgr = synth.rgbd_stream(nf, w, h, seed=7, device="cuda", distinct=nf)
f = JointBilateralFilter(w, h, window_radius=r, max_batch=nf)
out = torch.empty_like(depth)
for _ in range(3):
    f.process_batch(depth, bgr, out)
torch.cuda.synchronize()
print("ok", r, float(out.sum()))
