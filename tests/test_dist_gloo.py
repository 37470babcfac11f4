"""Host logic of the multi-GPU path on CPU: world_size-2 (and 3) gloo process groups exercise the sharding layout
and the one all-to-all exchange with NumPy standing in for the CUDA gather/scatter kernels.

Checks: (1) every zonal wavenumber has exactly one owner, bands tile the latitude pairs; (2) what rank a packs
for rank b is exactly what b expects from a; (3) after pack -> all_to_all -> unpack every rank holds, for its
latitude band, the rows of ALL zonal wavenumbers, bit for bit (rows are tagged with their global identity)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, gridname, T, nf, q):
    import torch
    import torch.distributed as dist

    import atlas_b200
    from atlas_b200 import dist as spdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        grid = atlas_b200.Grid(gridname)
        owner, band, m_rows, b_rows = spdist.shard_layout(grid, T, rank, world)
        seg_m = spdist.shard_segments(grid, T, rank, world, 0)
        seg_b = spdist.shard_segments(grid, T, rank, world, 1)
        nleg = (grid.ny() + 1) // 2
        # global row identity: row r of the exchange buffer <-> (m, parity, latitude); tag = r (same layout on all ranks)
        from atlas_b200 import _lib

        # total rows = fb_rowoff[T+1]; derive from the segments' extent on the band side of a 1-rank layout
        seg_all = spdist.shard_segments(grid, T, 0, 1, 0)
        total_rows = int((seg_all[:, 0] + seg_all[:, 2]).max())
        fb = np.full((total_rows, nf), -1.0)
        # this rank computed (Legendre stage) the rows of its own zonal wavenumbers: tag them with row id * 10 + field
        for fb_row, buf_row, n in seg_m:
            fb[fb_row:fb_row + n] = (np.arange(fb_row, fb_row + n)[:, None] * 16.0 + np.arange(nf)[None, :])
        send = np.zeros((int(m_rows.sum()), nf))
        for fb_row, buf_row, n in seg_m:
            send[buf_row:buf_row + n] = fb[fb_row:fb_row + n]
        recv = np.zeros((int(b_rows.sum()), nf))
        ts, tr = torch.from_numpy(send.reshape(-1)), torch.from_numpy(recv.reshape(-1))
        dist.all_to_all_single(tr, ts, [int(r) * nf for r in b_rows], [int(r) * nf for r in m_rows])
        recv = tr.numpy().reshape(-1, nf)
        fb2 = np.full((total_rows, nf), -1.0)
        for fb_row, buf_row, n in seg_b:
            fb2[fb_row:fb_row + n] = recv[buf_row:buf_row + n]
        # expectation: for every m (any owner) and both parities, the rows of my band are present and tagged right
        ok = True
        covered = 0
        for fb_row, buf_row, n in seg_b:
            want = (np.arange(fb_row, fb_row + n)[:, None] * 16.0 + np.arange(nf)[None, :])
            ok &= np.array_equal(fb2[fb_row:fb_row + n], want)
            covered += n
        # counts agree pairwise
        allm = [None] * world
        allb = [None] * world
        dist.all_gather_object(allm, m_rows.tolist())
        dist.all_gather_object(allb, b_rows.tolist())
        for a in range(world):
            for b in range(world):
                ok &= allm[a][b] == allb[b][a]
        q.put((rank, bool(ok), owner.tolist(), band.tolist(), covered, total_rows))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,gridname,T", [(2, "O32", 31), (3, "O48", 47), (2, "F16", 15)])
def test_exchange_layout_and_all_to_all(world, gridname, T):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, gridname, T, 2, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    res.sort()
    owners = res[0][2]
    band = res[0][3]
    assert all(r[1] for r in res), res
    assert all(r[2] == owners and r[3] == band for r in res)
    assert sorted(set(owners)) == list(range(world))          # every rank owns some zonal wavenumbers
    assert band[0] == 0 and band[-1] == (len(band) and band[-1]) and all(b1 >= b0 for b0, b1 in zip(band, band[1:]))
    assert sum(r[4] for r in res) == res[0][5]                  # bands tile all rows of the exchange buffer
