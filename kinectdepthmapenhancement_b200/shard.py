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
  single-GPU result bit for bit: a pixel's arithmetic depends only on the image around it, not on the
  tile or band it falls in (band boundaries are kept on multiples of 16 rows only for even tiling).

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


class PeerHalo:
    """Band arrays in NVLink-mapped symmetric memory: no exchange step at all.  Every rank allocates its
    extended arrays with torch symmetric memory, the ranks rendezvous once, and the kernels read the halo
    rows directly from the neighbours' bands (peer loads in the staging code of the seam tiles).  The local
    halo rows of the arrays are never written or read."""

    def __init__(self, plan: BandPlan, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.plan = p = plan
        group = group if group is not None else dist.group.WORLD
        plans = [make_plan(p.width, p.height, p.radius, r, p.world) for r in range(p.world)]
        max_rows = max(q.ext_rows for q in plans)
        self._depth_flat = symm_mem.empty(max_rows * p.width, dtype=torch.float32, device=device)
        self._bgr_flat = symm_mem.empty(max_rows * p.width * 3, dtype=torch.uint8, device=device)
        self._hd = symm_mem.rendezvous(self._depth_flat, group=group)
        self._hb = symm_mem.rendezvous(self._bgr_flat, group=group)
        self.depth_ext = self._depth_flat[:p.ext_rows * p.width].view(p.ext_rows, p.width)
        self.bgr_ext = self._bgr_flat[:p.ext_rows * p.width * 3].view(p.ext_rows, p.width, 3)
        self.depth_band = self.depth_ext[p.up:p.up + p.band_rows]
        self.bgr_band = self.bgr_ext[p.up:p.up + p.band_rows]
        # peer-mapped views of the neighbours' whole symmetric buffers (kept alive here)
        self._peers = {}
        self.depth_up = self.depth_dn = self.bgr_up = self.bgr_dn = 0
        if p.rank > 0:
            q = plans[p.rank - 1]
            d = self._hd.get_buffer(p.rank - 1, (max_rows * p.width,), torch.float32)
            b = self._hb.get_buffer(p.rank - 1, (max_rows * p.width * 3,), torch.uint8)
            self._peers["up"] = (d, b)
            first = q.up + q.band_rows - p.up          # neighbour row that is my ext row 0
            self.depth_up = d.data_ptr() + first * p.width * 4
            self.bgr_up = b.data_ptr() + first * p.width * 3
        if p.rank < p.world - 1:
            q = plans[p.rank + 1]
            d = self._hd.get_buffer(p.rank + 1, (max_rows * p.width,), torch.float32)
            b = self._hb.get_buffer(p.rank + 1, (max_rows * p.width * 3,), torch.uint8)
            self._peers["dn"] = (d, b)
            self.depth_dn = d.data_ptr() + q.up * p.width * 4   # neighbour's first band row = my ext row band1
            self.bgr_dn = b.data_ptr() + q.up * p.width * 3

    def barrier(self):
        """Device-side barrier across the ranks on the current stream: bands written / halo reads finished."""
        self._hd.barrier()


class RowBandJBF:
    """One rank's share of JointBilateralFilter::Process on a frame split into row bands."""

    def __init__(self, width, height, radius, rank, world, spatial_sigma=70.0, color_sigma=50.0, depth_sigma=20.0,
                 device=None, group=None, peer_memory: bool = False):
        """peer_memory=False: halos filled by one batched NCCL isend/irecv per frame (HaloExchanger).
        peer_memory=True: bands live in NVLink-mapped symmetric memory and the kernels read the halo rows from
        the neighbour GPUs themselves (PeerHalo) -- no exchange step, only a device-side barrier."""
        from .jbf import JointBilateralFilter
        self.plan = make_plan(width, height, radius, rank, world)
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device) if isinstance(device, int) else device
        self.peer_memory = bool(peer_memory) and world > 1
        self.halo = PeerHalo(self.plan, self.device, group) if self.peer_memory else HaloExchanger(self.plan, self.device, group)
        p = self.plan
        self.jbf = JointBilateralFilter(width, p.ext_rows, spatial_sigma, color_sigma, depth_sigma, radius,
                                        max_batch=1, device=self.device.index)
        self.pitch = (width + 3) & ~3
        self.guide4 = torch.empty((p.ext_rows, self.pitch), dtype=torch.int32, device=self.device)
        self.out = torch.empty((p.band_rows, width), dtype=torch.float32, device=self.device)
        self._strip = None

    @property
    def depth_band(self):
        return self.halo.depth_band

    @property
    def bgr_band(self):
        return self.halo.bgr_band

    def _presmooth(self, row0: int, rows: int, keep0: int | None = None, keep1: int | None = None) -> None:
        """Pre-smooth ext rows [row0, row0 + rows) (reflect-101 at the slice ends).  Without keep0/keep1 all of
        them are written into guide4; with them the slice is smoothed into a scratch strip and only rows
        [keep0, keep1) are copied over (the slice's own two end rows are reflect-contaminated)."""
        from . import _lib
        p = self.plan
        if keep0 is None:
            dst = self.guide4[row0:]
        else:
            if self._strip is None or self._strip.shape[0] < rows:
                self._strip = torch.empty((rows, self.pitch), dtype=torch.int32, device=self.device)
            dst = self._strip
        _lib.check(_lib.lib().jbf_presmooth_rows(self.jbf._h, self.halo.bgr_ext[row0:].data_ptr(), 3 * p.width,
                                                 dst.data_ptr(), self.pitch * 4, rows))
        if keep0 is not None:
            self.guide4[keep0:keep1].copy_(dst[keep0 - row0:keep1 - row0])

    def _filter(self, y_off: int, out_rows: int) -> None:
        from . import _lib
        p = self.plan
        _lib.check(_lib.lib().jbf_filter_rows(self.jbf._h, self.halo.depth_ext.data_ptr(), self.guide4.data_ptr(),
                                              self.pitch * 4, self.out[y_off - p.up:].data_ptr(), p.ext_rows, y_off,
                                              out_rows))

    def process(self, exchange: bool = True, overlap: bool = True) -> torch.Tensor:
        """Filter this rank's band.  With overlap=True the neighbour exchange runs while the band's interior
        rows (those whose window and pre-smooth footprint stay inside the band) are filtered; the two seam
        strips follow once the halos have landed.  The result is bit-identical to the unsplit launch and to the
        single-GPU frame (tile-independent arithmetic)."""
        p = self.plan
        if self.peer_memory:
            return self.process_peer(barrier=exchange)
        edge = TILE_H * ((p.halo + TILE_H - 1) // TILE_H)          # seam strip height (tile multiple >= r + 2)
        top = edge if p.up else 0
        bot = edge if p.down else 0
        if not (overlap and p.world > 1) or p.band_rows < top + bot + TILE_H:
            if exchange:
                self.halo.exchange()
            self._presmooth(0, p.ext_rows)
            self._filter(p.up, p.band_rows)
            return self.out
        works = self.halo.start() if exchange else []
        # interior: pre-smooth the band alone (its two outermost rows are reflect-contaminated and are redone
        # below); rows [up + top, up + band - bot_rows) only look at rows >= 2 inside the band
        self._presmooth(p.up, p.band_rows)
        bot_rows = bot + ((p.band_rows - top - bot) % TILE_H if bot else 0)   # keep the interior a tile multiple
        interior = p.band_rows - top - bot_rows
        self._filter(p.up + top, interior)
        for w in works:
            w.wait()
        R2 = PRESMOOTH_RADIUS
        if top:      # guide rows [0, up + 2) need the upper halo; rows from up + 2 on are already right
            self._presmooth(0, p.up + 2 * R2, keep0=0, keep1=p.up + R2)
            self._filter(p.up, top)
        if bot:
            r0 = p.up + p.band_rows - 2 * R2
            self._presmooth(r0, p.ext_rows - r0, keep0=r0 + R2, keep1=p.ext_rows)
            self._filter(p.up + p.band_rows - bot_rows, bot_rows)
        return self.out

    def process_peer(self, barrier: bool = True, pointers=None) -> torch.Tensor:
        """Peer-memory schedule: [barrier] -> pre-smooth (halo rows of BGR read from the neighbours) ->
        filter (halo rows of depth read from the neighbours) -> [barrier].  `pointers` overrides the peer
        pointers (single-GPU emulation in the tests): (depth_up, depth_dn, bgr_up, bgr_dn)."""
        from . import _lib
        p = self.plan
        if pointers is None:
            pointers = (self.halo.depth_up, self.halo.depth_dn, self.halo.bgr_up, self.halo.bgr_dn)
        d_up, d_dn, b_up, b_dn = [None if not x else x for x in pointers]
        if barrier:
            self.halo.barrier()
        L = _lib.lib()
        h = self.jbf._h
        band0, band1 = p.up, p.up + p.band_rows
        _lib.check(L.jbf_presmooth_rows_p2p(h, self.halo.bgr_ext.data_ptr(), 3 * p.width, self.guide4.data_ptr(),
                                            self.pitch * 4, p.ext_rows, band0, band1, b_up, b_dn))
        _lib.check(L.jbf_filter_rows_p2p(h, self.halo.depth_ext.data_ptr(), self.guide4.data_ptr(), self.pitch * 4,
                                         self.out.data_ptr(), p.ext_rows, p.up, p.band_rows, band0, band1, d_up, d_dn))
        if barrier:
            self.halo.barrier()
        return self.out
