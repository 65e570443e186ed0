"""Host-side mirror of ArrayBuffer / Buffer2D (ArrayBuffer.h:9-45, Buffer2D.h:9-35)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .jbf import _check_cuda, _ptr, _tensor_view


class Buffer2D:
    """Buffer2D(width, height) -- Buffer2D.cpp:4-10: device AoS {d, w} running weighted depth."""

    def __init__(self, width: int, height: int, device=None, stream: torch.cuda.Stream | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("Buffer2D needs a CUDA device; there is no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else device.index or 0)
        self.width, self.height = int(width), int(height)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            sptr = stream.cuda_stream if stream is not None else torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().buf2d_create(C.byref(self._h), self.width, self.height, self.device.index, sptr))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().buf2d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _plane(self, t, name):
        _check_cuda(t, torch.float32, name, self.device)
        if t.numel() != self.width * self.height:
            raise ValueError(f"{name} must hold width*height floats")

    def initDeviceMemoryElements(self):
        _lib.check(_lib.lib().buf2d_init(self._h))

    def insertData(self, data: torch.Tensor):
        """insertData(float*) [H,W]; insertData(weighted_d*) [H,W,2] with as_weighted=True via insertWeighted;
        insertData(float2*) via insertData2."""
        self._plane(data, "data")
        _lib.check(_lib.lib().buf2d_insert_f32(self._h, _ptr(data)))

    def insertWeighted(self, dw: torch.Tensor):
        """insertData(weighted_d*) -- Buffer2D.cpp:13-15."""
        _check_cuda(dw, torch.float32, "dw", self.device)
        if dw.numel() != 2 * self.width * self.height:
            raise ValueError("dw must hold width*height {d,w} pairs")
        _lib.check(_lib.lib().buf2d_insert_dw(self._h, _ptr(dw)))

    def insertData2(self, xy: torch.Tensor):
        """insertData(float2*) -- Buffer2D.cu:123-147 (w = row index, as the reference does)."""
        _check_cuda(xy, torch.float32, "xy", self.device)
        if xy.numel() != 2 * self.width * self.height:
            raise ValueError("xy must hold width*height float2")
        _lib.check(_lib.lib().buf2d_insert_f32x2(self._h, _ptr(xy)))

    def updateData(self, data: torch.Tensor):
        """updateData(float*) -- Buffer2D.cu:97-120; [N,H,W] input applies N frames in one pass."""
        _check_cuda(data, torch.float32, "data", self.device)
        n = data.numel() // (self.width * self.height)
        if n < 1 or data.numel() != n * self.width * self.height:
            raise ValueError("data must hold a whole number of frames")
        _lib.check(_lib.lib().buf2d_update_batch_f32(self._h, _ptr(data), n))

    def updateDataU16Host(self, depth_u16: torch.Tensor):
        """insertData(xn::DepthMetaData*) without OpenNI -- Buffer2D.cpp:18-32."""
        if depth_u16.is_cuda or depth_u16.dtype not in (torch.uint16, torch.int16) or not depth_u16.is_contiguous():
            raise TypeError("depth_u16 must be a contiguous CPU uint16 tensor")
        _lib.check(_lib.lib().buf2d_update_u16_host(self._h, _ptr(depth_u16)))

    def getDepthMap(self, out: torch.Tensor | None = None) -> torch.Tensor:
        if out is None:
            out = torch.empty((self.height, self.width), dtype=torch.float32, device=self.device)
        self._plane(out, "out")
        _lib.check(_lib.lib().buf2d_get_depth(self._h, _ptr(out)))
        return out

    def getWeightMap(self, out: torch.Tensor | None = None) -> torch.Tensor:
        if out is None:
            out = torch.empty((self.height, self.width), dtype=torch.float32, device=self.device)
        self._plane(out, "out")
        _lib.check(_lib.lib().buf2d_get_weight(self._h, _ptr(out)))
        return out

    def getRawPointer(self) -> torch.Tensor:
        """weighted_d* getRawPointer() -- ArrayBuffer.cpp:19-21, as a borrowed [H,W,2] view."""
        return _tensor_view(_lib.lib().buf2d_raw(self._h), (self.height, self.width, 2), torch.float32,
                            self.device, self)
