"""Row-band planning and the halo exchange (the path's only exchange step) on CPU with gloo, world sizes 2 and 3."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_frame_super_resolution_b200 import rowband


def test_plan_bands_cover_align_halo():
    for height, world in ((6048, 8), (6048, 2), (3024, 4), (1024, 3)):
        bands = rowband.plan_bands(height, world, 128, 256 if height // world >= 256 else 128)
        assert bands[0].row0 == 0 and bands[-1].row1 == height
        for b, nxt in zip(bands, bands[1:] + [None]):
            assert b.row0 % 128 == 0 and b.top % 128 == 0 and b.rows > 0
            assert b.top == max(0, b.row0 - (b.row0 - b.top)) and b.bottom <= height
            if nxt:
                assert b.row1 == nxt.row0
    with pytest.raises(ValueError):
        rowband.plan_bands(300, 4)            # not enough rows for aligned bands
    with pytest.raises(ValueError):
        rowband.plan_bands(6048, 2, 128, 100)  # halo must keep band + halo origins aligned
    with pytest.raises(ValueError):
        rowband.plan_bands(1024, 8, 128, 256)  # halo would span more than the neighbouring band


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, h, w = 3, 1024 + 64, 40
    full = (torch.arange(n * h * w, dtype=torch.int64).reshape(n, h, w) % 30011).to(torch.int16)   # every rank can rebuild the truth
    bands = rowband.plan_bands(h, world, 128, 128)
    b = bands[rank]
    got = rowband.exchange_halos(full[:, b.row0:b.row1].clone(), bands, rank)
    ok = got.dtype == torch.int16 and got.shape == (n, b.bottom - b.top, w) and torch.equal(got, full[:, b.top:b.bottom])
    q.put((rank, bool(ok), (b.row0, b.row1, b.top, b.bottom)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() + world * 7) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == list(range(world))
    assert all(r[1] for r in res), res


def test_peer_halo_exchange_indexing():
    """rowband.PeerHaloExchange: every rank reads its halo rows out of the neighbours' band buffers (own rows sit behind the neighbour's
    own upper halo).  The peer mapping is replaced by a stand-in that hands out the other ranks' buffers; 3 bands, uneven last band."""
    import torch
    from multi_frame_super_resolution_b200 import rowband

    n, H, W = 2, 128 * 7, 24
    full = torch.arange(n * H * W, dtype=torch.int32).remainder(30011).to(torch.int16).reshape(n, H, W)
    bands = rowband.plan_bands(H, 3, 128, 256)
    rows_max = max(b.bottom - b.top for b in bands)
    bufs = [torch.zeros((n, rows_max, W * 2), dtype=torch.uint8) for _ in bands]

    class Hdl:
        def get_buffer(self, rank, sizes, dtype):
            assert tuple(sizes) == tuple(bufs[rank].shape) and dtype == torch.uint8
            return bufs[rank]

        def barrier(self, channel=0):
            pass

    ex = [rowband.PeerHaloExchange.from_handle(bufs[r], Hdl(), bands, r) for r in range(3)]
    for r, b in enumerate(bands):
        ex[r].own_view().copy_(full[:, b.row0:b.row1])
    for r, b in enumerate(bands):
        got = ex[r].exchange()
        assert got.shape == (n, b.bottom - b.top, W) and got.dtype == torch.int16
        assert torch.equal(got, full[:, b.top:b.bottom]), r
        assert all(got[f].is_contiguous() for f in range(n))            # dense frames, a larger frame stride: what set_input accepts
