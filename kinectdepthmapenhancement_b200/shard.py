"""Multi-GPU partitioning of the joint-bilateral path on one NVLink/NVSwitch box.

The reference is single-GPU (no cudaSetDevice / streams / NCCL anywhere); this is what the path
needs to use 2/4/8 B200s, one process per GPU over torch.distributed:

* frame streams -- fully independent units (JointBilateralFilter::Process carries no state between
  frames, JointBilateralFilter.cu:283-290): contiguous frame ranges per rank, NO data-path collective.
* one very large frame -- row bands whose boundaries are multiples of the kernel's tile height; each
  rank needs `radius` rows of depth and smoothed guide above and below its band.  Raw inputs are
  exchanged with a halo of radius + 2 rows (2 = pre-smooth radius) so the pre-smooth is recomputed
  locally on the halo and ONE neighbour exchange per frame suffices: a batched isend/irecv
  (ncclSend/ncclRecv inside one group over NVLink with the NCCL backend).  Edge ranks skip the
  missing neighbour; the image border is the kernel's zero fill.  The band result equals the
  single-GPU result bit for bit (same tiles, same summation order).

Everything in HaloExchanger works on CPU tensors with the gloo backend too (used by the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

TILE_H = 16          # jbf_fast_kernel tile height; band boundaries are multiples of it
PRESMOOTH_RADIUS = 2  # cv::gpu::bilateralFilter(.., 5, ..) at JointBilateralFilter.cu:285


def frame_shard(n_frames: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous frame range [start, stop) of `rank`; the remainder goes to the first ranks."""
    if world < 1 or not (0 <= rank < world) or n_frames < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def band_partition(height: int, world: int, tile_h: int = TILE_H) -> list[tuple[int, int]]:
    """Row bands [y0, y1) per rank, boundaries on multiples of tile_h, sizes differing by <= tile_h."""
    if world < 1 or height < 1:
        raise ValueError("bad partition arguments")
    n_tiles = (height + tile_h - 1) // tile_h
    if n_tiles < world:
        raise ValueError(f"{height} rows give only {n_tiles} tile rows for {world} ranks")
    bands = []
    for r in range(world):
        t0, t1 = frame_shard(n_tiles, world, r)
        bands.append((t0 * tile_h, min(t1 * tile_h, height)))
    return bands


@dataclass
class BandPlan:
    rank: int
    world: int
    width: int
    height: int
    radius: int
    y0: int
    y1: int
    up: int     # halo rows held above the band (0 at the image top)
    down: int   # halo rows held below the band (0 at the image bottom)

    @property
    def band_rows(self) -> int:
        return self.y1 - self.y0

    @property
    def ext_rows(self) -> int:
        return self.up + self.band_rows + self.down

    @property
    def halo(self) -> int:
        return self.radius + PRESMOOTH_RADIUS

    def halo_bytes_per_direction(self) -> int:
        return self.halo * self.width * (4 + 3)


def make_plan(width: int, height: int, radius: int, rank: int, world: int) -> BandPlan:
    bands = band_partition(height, world)
    y0, y1 = bands[rank]
    halo = radius + PRESMOOTH_RADIUS
    for (a, b) in bands:
        if b - a < halo and world > 1:
            raise ValueError(f"band of {b - a} rows is thinner than the halo ({halo} rows); use fewer ranks")
    return BandPlan(rank, world, width, height, radius, y0, y1, min(halo, y0), min(halo, height - y1))


class HaloExchanger:
    """Owns the extended (band + halo) depth and BGR arrays of one rank and fills the halos from the
    neighbouring ranks with one batched isend/irecv."""

    def __init__(self, plan: BandPlan, device="cpu", group=None):
        self.plan, self.group = plan, group
        p = plan
        self.depth_ext = torch.zeros((p.ext_rows, p.width), dtype=torch.float32, device=device)
        self.bgr_ext = torch.zeros((p.ext_rows, p.width, 3), dtype=torch.uint8, device=device)
        self.depth_band = self.depth_ext[p.up:p.up + p.band_rows]
        self.bgr_band = self.bgr_ext[p.up:p.up + p.band_rows]

    def start(self):
        """Post the neighbour exchange; returns the work handles (empty for world == 1)."""
        p = self.plan
        if p.world == 1:
            return []
        ops = []
        h = p.halo
        top, bot = p.up, p.up + p.band_rows
        if p.rank > 0:  # upper neighbour: my first h band rows are its lower halo; its last h rows are my upper halo
            ops += [dist.P2POp(dist.isend, self.depth_ext[top:top + h], p.rank - 1, self.group),
                    dist.P2POp(dist.isend, self.bgr_ext[top:top + h], p.rank - 1, self.group),
                    dist.P2POp(dist.irecv, self.depth_ext[0:p.up], p.rank - 1, self.group),
                    dist.P2POp(dist.irecv, self.bgr_ext[0:p.up], p.rank - 1, self.group)]
        if p.rank < p.world - 1:
            ops += [dist.P2POp(dist.isend, self.depth_ext[bot - h:bot], p.rank + 1, self.group),
                    dist.P2POp(dist.isend, self.bgr_ext[bot - h:bot], p.rank + 1, self.group),
                    dist.P2POp(dist.irecv, self.depth_ext[bot:bot + p.down], p.rank + 1, self.group),
                    dist.P2POp(dist.irecv, self.bgr_ext[bot:bot + p.down], p.rank + 1, self.group)]
        return dist.batch_isend_irecv(ops)

    def exchange(self):
        for w in self.start():
            w.wait()


class RowBandJBF:
    """One rank's share of JointBilateralFilter::Process on a frame split into row bands."""

    def __init__(self, width, height, radius, rank, world, spatial_sigma=70.0, color_sigma=50.0, depth_sigma=20.0,
                 device=None, group=None):
        from .jbf import JointBilateralFilter
        self.plan = make_plan(width, height, radius, rank, world)
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device) if isinstance(device, int) else device
        self.halo = HaloExchanger(self.plan, self.device, group)
        p = self.plan
        self.jbf = JointBilateralFilter(width, p.ext_rows, spatial_sigma, color_sigma, depth_sigma, radius,
                                        max_batch=1, device=self.device.index)
        self.pitch = (width + 3) & ~3
        self.guide4 = torch.empty((p.ext_rows, self.pitch), dtype=torch.int32, device=self.device)
        self.out = torch.empty((p.band_rows, width), dtype=torch.float32, device=self.device)

    @property
    def depth_band(self):
        return self.halo.depth_band

    @property
    def bgr_band(self):
        return self.halo.bgr_band

    def process(self, exchange: bool = True) -> torch.Tensor:
        from . import _lib
        p = self.plan
        if exchange:
            self.halo.exchange()
        L = _lib.lib()
        h = self.jbf._h
        _lib.check(L.jbf_presmooth_rows(h, self.halo.bgr_ext.data_ptr(), 3 * p.width, self.guide4.data_ptr(),
                                        self.pitch * 4, p.ext_rows))
        _lib.check(L.jbf_filter_rows(h, self.halo.depth_ext.data_ptr(), self.guide4.data_ptr(), self.pitch * 4,
                                     self.out.data_ptr(), p.ext_rows, p.up, p.band_rows))
        return self.out
