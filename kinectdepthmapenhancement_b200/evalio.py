"""The callers either side of the path: the reference's on-disk depth format and
its evaluation metric.

* depth.xml -- OpenCV FileStorage XML with `opencv-matrix` nodes `averaged_depth` and `depth`
  (cv::Mat_<float> 480x640, dt "f"), written by main.cpp:112-114 and read by main.cpp:146-149.
  Read/written here without OpenCV (plain XML), interoperable with cv::FileStorage.
* ground truth -- the 1000-frame running average kept by Buffer2D::updateData (main.cpp:86-105).
* error metric -- mean 3-D distance between a method's back-projected cloud and the averaged cloud over
  pixels whose z are both in (50, 15000) mm (main.cpp:217-308), computed on the GPU.
"""
from __future__ import annotations

import ctypes as C
import xml.etree.ElementTree as ET

import numpy as np
import torch

from . import _lib


# ------------------------------------------------------------------ depth.xml (host-side file format)
def write_depth_xml(path: str, mats: dict[str, np.ndarray]) -> None:
    """cv::FileStorage(path, WRITE); cv::write(fs, name, mat) for float32 matrices (main.cpp:112-114)."""
    parts = ['<?xml version="1.0"?>\n<opencv_storage>\n']
    for name, m in mats.items():
        m = np.ascontiguousarray(m, dtype=np.float32)
        if m.ndim != 2:
            raise ValueError("matrices must be 2-D")
        parts.append(f'<{name} type_id="opencv-matrix">\n  <rows>{m.shape[0]}</rows>\n  <cols>{m.shape[1]}</cols>\n'
                     f'  <dt>f</dt>\n  <data>\n')
        flat = m.reshape(-1)
        for i in range(0, flat.size, 4):
            parts.append("    " + " ".join("%.8e" % v for v in flat[i:i + 4]) + "\n")
        parts.append(f'  </data></{name}>\n')
    parts.append('</opencv_storage>\n')
    with open(path, "w") as f:
        f.write("".join(parts))


def read_depth_xml(path: str) -> dict[str, np.ndarray]:
    """cv::FileStorage(path, READ); cv::read(node[name], mat) (main.cpp:146-149)."""
    root = ET.parse(path).getroot()
    if root.tag != "opencv_storage":
        raise ValueError("not an OpenCV FileStorage XML file")
    out = {}
    for node in root:
        if node.get("type_id") != "opencv-matrix":
            continue
        rows, cols = int(node.findtext("rows")), int(node.findtext("cols"))
        dt = (node.findtext("dt") or "").strip()
        if dt not in ("f", "1f"):
            raise ValueError(f"{node.tag}: unsupported element type {dt!r} (expected 'f')")
        data = np.array((node.findtext("data") or "").split(), dtype=np.float32)
        if data.size != rows * cols:
            raise ValueError(f"{node.tag}: {data.size} values for a {rows}x{cols} matrix")
        out[node.tag] = data.reshape(rows, cols)
    return out


# ------------------------------------------------------------------ device-side pieces
def mean_3d_error(points: torch.Tensor, truth: torch.Tensor) -> tuple[float, int]:
    """main.cpp:217-308.  points/truth: [H,W,3] float32 CUDA clouds (mm).  Returns (mean distance, count)."""
    for t in (points, truth):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.shape[-1] != 3:
            raise TypeError("clouds must be contiguous float32 CUDA tensors [...,3]")
    if points.shape != truth.shape:
        raise ValueError("clouds must have the same shape")
    mean, cnt = C.c_double(), C.c_longlong()
    with torch.cuda.device(points.device):
        _lib.check(_lib.lib().kdme_mean_3d_error(points.data_ptr(), truth.data_ptr(), points.numel() // 3,
                                                 C.byref(mean), C.byref(cnt), torch.cuda.current_stream().cuda_stream))
    return mean.value, cnt.value


def depth_bilateral_xyz(normalized: torch.Tensor, points: torch.Tensor, window_radius: int = 3,
                        spatial_sigma: float = 20.0, depth_sigma: float = 100.0) -> torch.Tensor:
    """Projection_GPU::bilateralfilter (Projection_GPU.cu:213-246; consts Projection_GPU.cpp:3-5), race-free."""
    for t in (normalized, points):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.shape[-1] != 3:
            raise TypeError("clouds must be contiguous float32 CUDA tensors [H,W,3]")
    h, w, _ = points.shape
    out = torch.empty_like(points)
    with torch.cuda.device(points.device):
        _lib.check(_lib.lib().kdme_depth_bilateral_xyz(normalized.data_ptr(), points.data_ptr(), out.data_ptr(), w, h,
                                                       window_radius, spatial_sigma, depth_sigma,
                                                       torch.cuda.current_stream().cuda_stream))
    return out


def average_depth(frames: torch.Tensor) -> torch.Tensor:
    """The capture loop of main.cpp:86-105: Buffer2D::updateData over N frames, then getDepthMap."""
    from .buffer2d import Buffer2D
    n, h, w = frames.shape
    b = Buffer2D(w, h, device=frames.device.index)
    if (w * h) % 4 == 0:
        b.updateData(frames)
    else:
        for i in range(n):
            b.updateData(frames[i])
    return b.getDepthMap()


def evaluate(depth: torch.Tensor, averaged_depth: torch.Tensor, color: torch.Tensor, fx: float, fy: float, cx: int,
             cy: int, **jbf_kw) -> dict:
    """The JBF leg of the reference's evaluation (main.cpp:160-183, 246-258, 303-305): filter, back-project,
    mean 3-D error of the input and of the filtered cloud against the averaged cloud."""
    from .jbf import JointBilateralFilter, projective_to_real
    h, w = depth.shape
    f = JointBilateralFilter(w, h, device=depth.device.index, **jbf_kw)
    f.Process(depth, color)
    truth = projective_to_real(averaged_depth, fx, fy, cx, cy)
    e_in, n_in = mean_3d_error(projective_to_real(depth, fx, fy, cx, cy), truth)
    e_jbf, n_jbf = mean_3d_error(projective_to_real(f.getFiltered_Device().clone(), fx, fy, cx, cy), truth)
    return {"input": e_in, "input_count": n_in, "jbf": e_jbf, "jbf_count": n_jbf}
