"""The header-only C++ classes keep the reference's call sites compiling (main.cpp:62-67,179-183)."""
import os
import subprocess
import tempfile

from conftest import ROOT

SNIPPET = r'''
#include "JointBilateralFilter.h"
#include "Buffer2D.h"
#include <cstdio>
// the reference's call pattern, verbatim in shape (main.cpp:62-67, 99, 104, 179-183)
int run(float* inputDepth_Device, cv::gpu::GpuMat Color_Device, float* bufferDepth_Device, int W, int H) {
    Buffer2D Buffer(W, H);
    JointBilateralFilter JBF(W, H);
    Buffer.updateData(inputDepth_Device);
    Buffer.getDepthMap(bufferDepth_Device);
    JBF.Process(inputDepth_Device, Color_Device);
    float* filtered = JBF.getFiltered_Device();
    cv::gpu::GpuMat smooth = JBF.getSmoothImage_Device();
    return filtered != 0 && smooth.data != 0;
}
int main() {
    try { JointBilateralFilter JBF(640, 480); }
    catch (const std::exception& e) { std::printf("expected without a GPU: %s\n", e.what()); return 0; }
    return 0;
}
'''


def test_reference_call_sites_compile_and_link():
    from kinectdepthmapenhancement_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "dropin.cpp")
        open(src, "w").write(SNIPPET)
        exe = os.path.join(td, "dropin")
        libdir = os.path.dirname(_lib.LIB_PATH)
        subprocess.run(["g++", "-std=c++11", "-DKDME_NO_OPENCV", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                        "-L", libdir, "-lkdme_b200", "-Wl,-rpath," + libdir], check=True)
        res = subprocess.run([exe], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
