"""Row-band sharding of ONE very large burst over the GPUs of a box (BASELINE config 4).

Every rank owns a contiguous band of raw rows of all N frames.  The only exchange step of the path is the halo: the
rows a band needs from its neighbours so that every stage (coarsest-level tile matching has the largest footprint)
sees true data around the rows it keeps.  Halo rows are copied out of the neighbours' memory over NVLink (`PeerHaloExchange`,
torch symmetric memory) or travel by NCCL send/recv (`exchange_halos`, one grouped batch per direction: also the gloo / CPU test
path); after that each rank runs the unmodified chain on band + halo in row-band mode
(`mfsr_params.band_*`), which maps tile rows through the full frame's coordinates and merges only the kept rows.
No other collective touches the data path.

Band origins are multiples of `tile_size << (levels - 1)` (128 rows by default) so that the tile grids and the 2x2
pyramid of a band coincide with the full-frame ones; flow-from-tiles and the LK warp evaluate their normalised texture
coordinates in full-frame coordinates; the merge window grows by one row at interior seams.  The stitched result is
bit-identical to the single-GPU full-frame result (tests/test_rowband_gpu.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import torch

DEFAULT_HALO = 256     # rows; two coarsest-level tiles (2 x 128 rows) cover the matcher's footprint at 4 levels
DEFAULT_MARGIN = 64    # legacy margin (covers vertical flows up to ~35 rows); default_margin(params) derives the safe one


def default_margin(params) -> int:
    """Rows beyond the kept rows on which flow / kernel parameters / robustness must be evaluated (mfsr_params.band_margin) so that
    the kept rows cannot see the band's edge: stencil footprints (LK 3 sweeps x 5 rows + robustness 12 = 27 rows) plus the largest
    vertical flow the aligner can produce per measured pair, max_shift * (2^levels - 1), plus the global base shift; rounded up to
    the LK tile height (32).  With a smaller margin large vertical motion silently clamps at the band's edge."""
    reach = 27 + int(params.max_shift) * ((1 << int(params.levels)) - 1) + int(abs(params.base_shift[1]) + 0.999)
    return (reach + 31) // 32 * 32


@dataclass(frozen=True)
class Band:
    rank: int
    row0: int        # first owned raw row (global)
    row1: int        # one past the last owned raw row
    top: int         # first row of band + halo (global, multiple of the grid)
    bottom: int      # one past the last row of band + halo

    @property
    def rows(self) -> int:
        return self.row1 - self.row0

    @property
    def halo_up(self) -> int:
        return self.row0 - self.top

    @property
    def halo_down(self) -> int:
        return self.bottom - self.row1


def plan_bands(height: int, world: int, grid: int = 128, halo: int = DEFAULT_HALO) -> List[Band]:
    """Contiguous bands whose origins are multiples of `grid`; the last band takes the remainder."""
    if height < grid * world:
        raise ValueError(f"{height} rows cannot be split into {world} bands aligned to {grid}")
    if halo % grid:
        raise ValueError("halo must be a multiple of the grid (band + halo origins stay aligned)")
    units = height // grid
    base, extra = divmod(units, world)
    bands, r = [], 0
    for k in range(world):
        n = (base + (1 if k < extra else 0)) * grid
        r1 = height if k == world - 1 else r + n
        bands.append(Band(k, r, r1, max(0, r - halo), min(height, r1 + halo)))
        r = r1
    for b in bands:
        if world > 1 and (b.halo_up > bands[b.rank - 1].rows if b.rank > 0 else False):
            raise ValueError("halo larger than a neighbouring band: use fewer ranks or a smaller halo")
        if world > 1 and b.rank < world - 1 and b.halo_down > bands[b.rank + 1].rows:
            raise ValueError("halo larger than a neighbouring band: use fewer ranks or a smaller halo")
    return bands


def exchange_halos(own: torch.Tensor, bands: List[Band], rank: int, group=None) -> torch.Tensor:
    """own: [N, band.rows, W] (the rank's rows of every frame).  Returns [N, band + halo rows, W].

    One send and one receive per neighbour, all posted as a single batch (torch.distributed.batch_isend_irecv ->
    grouped ncclSend/ncclRecv on NCCL, plain sockets on gloo)."""
    import torch.distributed as dist
    b = bands[rank]
    dtype = own.dtype
    own = own.contiguous().view(torch.uint8)        # bytes on the wire: every backend moves uint8
    n, rows, w = own.shape
    assert rows == b.rows
    out = torch.empty((n, b.bottom - b.top, w), dtype=own.dtype, device=own.device)
    out[:, b.halo_up:b.halo_up + rows] = own
    ops, keep = [], []
    if rank > 0:                                   # upper neighbour: I need its last halo_up rows, it needs my first rows
        up = bands[rank - 1]
        recv = torch.empty((n, b.halo_up, w), dtype=own.dtype, device=own.device)
        send = own[:, :up.halo_down].contiguous()
        ops += [dist.P2POp(dist.isend, send, rank - 1, group), dist.P2POp(dist.irecv, recv, rank - 1, group)]
        keep.append(("up", recv, send))
    if rank < len(bands) - 1:
        dn = bands[rank + 1]
        recv = torch.empty((n, b.halo_down, w), dtype=own.dtype, device=own.device)
        send = own[:, rows - dn.halo_up:].contiguous()
        ops += [dist.P2POp(dist.isend, send, rank + 1, group), dist.P2POp(dist.irecv, recv, rank + 1, group)]
        keep.append(("down", recv, send))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    for side, recv, _ in keep:
        if side == "up":
            out[:, :b.halo_up] = recv
        else:
            out[:, b.halo_up + rows:] = recv
    return out.view(dtype)


class PeerHaloExchange:
    """The same exchange without a collective library on the data path.  The band buffer of every rank (own rows + halo rows of all
    frames) is allocated in torch's symmetric memory, i.e. mapped into every peer of the node; the rank's own rows are written
    straight into it (own_view()), and exchange() copies the halo rows out of the neighbours' buffers over NVLink / NVSwitch — one
    contiguous device-to-device copy per frame and neighbour, between two device-side barriers — and returns the band.  No staging
    copy of the own rows, no send / receive buffers.  The band buffer is reused by the next burst: the consumer of exchange()'s
    result (the handle reads it in place) has to be finished before own_view() is written again.
    Raises whatever torch raises where symmetric memory is not available (single GPU, gloo, no P2P): callers fall back to
    exchange_halos()."""

    def __init__(self, n_frames: int, bands: List[Band], rank: int, width_bytes: int, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.bands, self.rank, self.n, self.wb = bands, rank, n_frames, width_bytes
        self.rows_max = max(b.bottom - b.top for b in bands)
        self.buf = symm.empty((n_frames, self.rows_max, width_bytes), dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)

    @classmethod
    def from_handle(cls, buf: torch.Tensor, hdl, bands: List[Band], rank: int):
        """An exchange over an existing buffer + peer handle (anything with get_buffer(rank, sizes, dtype) and barrier(channel=...)):
        what __init__ builds from torch's symmetric memory; the CPU tests pass a stand-in that hands out the other ranks' buffers."""
        self = cls.__new__(cls)
        self.bands, self.rank, self.n, self.wb = bands, rank, buf.shape[0], buf.shape[2]
        self.rows_max = buf.shape[1]
        self.buf, self.hdl = buf, hdl
        return self

    def own_view(self, dtype=torch.int16) -> torch.Tensor:
        """[N, band.rows, W]: where the rank's own rows go (frames are rows_max rows apart)."""
        b = self.bands[self.rank]
        return self.buf[:, b.halo_up:b.halo_up + b.rows].view(dtype)

    def band_view(self, dtype=torch.int16) -> torch.Tensor:
        b = self.bands[self.rank]
        return self.buf[:, :b.bottom - b.top].view(dtype)

    def exchange(self, dtype=torch.int16) -> torch.Tensor:
        """Own rows are in own_view().  Returns [N, band + halo rows, W] (a view: each frame dense, frames rows_max rows apart)."""
        b = self.bands[self.rank]
        shape = (self.n, self.rows_max, self.wb)
        self.hdl.barrier(channel=0)                  # every rank's own rows are in place
        if self.rank > 0:
            up = self.bands[self.rank - 1]
            peer = self.hdl.get_buffer(self.rank - 1, shape, torch.uint8)
            lo = up.halo_up + up.rows - b.halo_up
            for f in range(self.n):
                self.buf[f, :b.halo_up].copy_(peer[f, lo:lo + b.halo_up], non_blocking=True)
        if self.rank < len(self.bands) - 1:
            dn = self.bands[self.rank + 1]
            peer = self.hdl.get_buffer(self.rank + 1, shape, torch.uint8)
            for f in range(self.n):
                self.buf[f, b.halo_up + b.rows:b.bottom - b.top].copy_(peer[f, dn.halo_up:dn.halo_up + b.halo_down], non_blocking=True)
        self.hdl.barrier(channel=1)                  # every rank has read: own rows may take the next burst
        return self.band_view(dtype)


def band_params(params, band: Band, global_height: int, margin=None):
    """Copy of `params` switched to row-band mode for `band`.  margin = 0 evaluates every stage on the whole band + halo; None derives
    the margin from the params (default_margin)."""
    p = type(params).from_buffer_copy(params)
    p.band_margin = int(default_margin(params) if margin is None else margin)
    p.full_frame = 1
    p.band_global_h = int(global_height)
    p.band_row0 = int(band.top)
    p.band_keep_row0 = int(band.halo_up)
    p.band_keep_rows = int(band.rows)
    return p


def run_band(params, frames_with_halo: torch.Tensor, band: Band, global_height: int, device: int = 0, ref_idx: int = 0):
    """Runs the chain on band + halo (CUDA tensor [N, rows, W], 16-bit) and returns the band's output rows [s*rows, s*W, 3]."""
    from .pipeline import BurstSuperResolution
    n, h, w = frames_with_halo.shape
    sr = BurstSuperResolution(band_params(params, band, global_height), device=device, max_width=w, max_height=h, max_frames=n)
    sr.set_input(frames_with_halo.contiguous(), ref_idx=ref_idx)
    out = sr.next_frame()
    sr.synchronize()
    res = out.clone()
    sr.close()
    return res
