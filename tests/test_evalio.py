"""Row f4: the reference's depth.xml format (interoperable with cv::FileStorage) and its error metric."""
import os

import numpy as np
import pytest
import torch

from conftest import synth_np

import oracle
from kinectdepthmapenhancement_b200 import evalio


def test_depth_xml_roundtrip_and_opencv_interop(tmp_path):
    cv2 = pytest.importorskip("cv2")
    d, _ = synth_np(64, 48, seed=2, frame=0)
    avg = (d * 0.999).astype(np.float32)
    p = str(tmp_path / "depth.xml")
    evalio.write_depth_xml(p, {"averaged_depth": avg, "depth": d})
    back = evalio.read_depth_xml(p)
    assert np.array_equal(back["depth"], d) and np.array_equal(back["averaged_depth"], avg)
    fs = cv2.FileStorage(p, cv2.FILE_STORAGE_READ)          # what main.cpp:146-149 does
    assert np.array_equal(fs.getNode("depth").mat(), d)
    assert np.array_equal(fs.getNode("averaged_depth").mat(), avg)
    fs.release()
    p2 = str(tmp_path / "depth_cv.xml")
    fs = cv2.FileStorage(p2, cv2.FILE_STORAGE_WRITE)        # what main.cpp:112-114 does
    fs.write("averaged_depth", avg)
    fs.write("depth", d)
    fs.release()
    back = evalio.read_depth_xml(p2)
    assert np.array_equal(back["depth"], d) and np.array_equal(back["averaged_depth"], avg)
    with pytest.raises(ValueError):
        open(p2, "w").write('<?xml version="1.0"?><other/>')
        evalio.read_depth_xml(p2)


def test_oracle_metric_and_f2_pinned_to_reference():
    d, _ = synth_np(96, 64, seed=5, frame=0)
    pts = oracle.projective_to_real(d, 525.0, 525.0, 48, 32)
    z = np.where(pts[..., 2] > 0, pts[..., 2], 1)
    norm = pts.copy()
    norm[..., 0] /= z
    norm[..., 1] /= z
    a = oracle.depth_bilateral_xyz(norm, pts)
    if oracle.ref_available():
        b = oracle.depth_bilateral_xyz(norm, pts, impl="ref")
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    a64 = oracle.depth_bilateral_xyz(norm, pts, precision="f64")
    assert np.array_equal(a[..., 2] > 0, a64[..., 2] > 0) and np.abs(a - a64).max() < 0.02
    m, n = oracle.mean_3d_error(a, pts)
    assert n == int(((a[..., 2] > 50) & (pts[..., 2] > 50)).sum()) and 0 < m < 50


@pytest.mark.gpu
def test_gpu_f2_f4_match_oracle():
    d, _ = synth_np(160, 120, seed=8, frame=1)
    pts = oracle.projective_to_real(d, 525.0, 525.0, 80, 60)
    z = np.where(pts[..., 2] > 0, pts[..., 2], 1)
    norm = pts.copy()
    norm[..., 0] /= z
    norm[..., 1] /= z
    want64 = oracle.depth_bilateral_xyz(norm, pts, precision="f64")
    got = evalio.depth_bilateral_xyz(torch.from_numpy(norm).cuda(), torch.from_numpy(pts).cuda()).cpu().numpy()
    assert np.array_equal(got[..., 2] > 0, want64[..., 2] > 0)
    assert np.abs(got[..., 2] - want64[..., 2]).max() <= 1e-3
    assert np.abs(got - want64).max() <= 2e-3
    m_o, n_o = oracle.mean_3d_error(want64, pts)
    m_g, n_g = evalio.mean_3d_error(torch.from_numpy(want64).cuda(), torch.from_numpy(pts).cuda())
    assert n_g == n_o and abs(m_g - m_o) <= 1e-4 * m_o


@pytest.mark.gpu
def test_gpu_evaluation_leg_matches_oracle_and_reduces_noise():
    """The reference's experiment in miniature (main.cpp:86-105, 160-183, 246-258, 303-305): 64 noisy frames
    averaged by Buffer2D::updateData are the truth; one noisy frame is filtered; the mean 3-D errors of the
    input and of the JBF cloud are reported.  Checked against the same leg run with the oracle, and -- on a
    smooth surface, where the filter is meant to help -- the JBF error must be the smaller one."""
    w, h = 320, 240
    ys, xs = torch.meshgrid(torch.arange(h, device="cuda"), torch.arange(w, device="cuda"), indexing="ij")
    base = (1200.0 + 0.9 * xs + 0.6 * ys).float()
    color = torch.full((h, w, 3), 120, dtype=torch.uint8, device="cuda")
    color[:, : w // 2, 1] = 160
    g = torch.Generator(device="cuda").manual_seed(1)
    # +-0.4 % of z: inside Buffer2D's 1 % acceptance gate (Buffer2D.cu:20)
    noisy = torch.stack([base * (1 + (torch.rand(base.shape, device="cuda", generator=g) - 0.5) * 0.008)
                         for _ in range(64)]).contiguous()
    truth = evalio.average_depth(noisy)
    fx = fy = 525.0
    res = evalio.evaluate(noisy[0].contiguous(), truth, color, fx, fy, w // 2, h // 2, window_radius=2)
    assert res["jbf_count"] == res["input_count"] == w * h
    assert res["jbf"] < 0.6 * res["input"], res
    # the same leg with the oracle
    ob = oracle.Buffer2D(w, h)
    for f in noisy.cpu().numpy():
        ob.update(f)
    assert np.array_equal(ob.depth_map().view(np.uint32), truth.cpu().numpy().view(np.uint32))
    d0, c0 = noisy[0].cpu().numpy(), color.cpu().numpy()
    filt = oracle.jbf_process(d0, c0, 5, precision="f64")
    t3 = oracle.projective_to_real(ob.depth_map(), fx, fy, w // 2, h // 2)
    e_in, _ = oracle.mean_3d_error(oracle.projective_to_real(d0, fx, fy, w // 2, h // 2), t3)
    e_jbf, _ = oracle.mean_3d_error(oracle.projective_to_real(filt, fx, fy, w // 2, h // 2), t3)
    assert abs(res["input"] - e_in) <= 1e-4 * e_in and abs(res["jbf"] - e_jbf) <= 1e-3 * e_jbf
