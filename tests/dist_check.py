"""Run under torch.distributed.run with N >= 2 GPUs: sharded inverse + direct transform vs the CPU oracle.
Used by tests/test_gpu_dist.py and by hand:  python -m torch.distributed.run --nproc-per-node 2 tests/dist_check.py O48 47 5"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import helpers as H  # noqa: E402

import atlas_b200  # noqa: E402
from atlas_b200.dist import ShardedTrans  # noqa: E402


def main():
    gridname, T, nf = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    exchange = sys.argv[4] if len(sys.argv) > 4 else "peer"
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    grid = atlas_b200.Grid(gridname)
    st = ShardedTrans(grid, T, local, exchange=exchange)
    sp = H.synthetic_spectra(T, nf)
    d_sp = torch.from_numpy(sp).cuda()
    # three back-to-back round trips without host synchronisation in between: exercises the alternation of the
    # two exchange buffers and the barrier epochs of the peer-memory path; the last one is checked
    for it in range(3):
        d_gp = torch.zeros(nf * grid.size(), dtype=torch.float64, device="cuda")
        st.invtrans(nf, d_sp, d_gp)
        d_tmp = torch.zeros_like(d_sp)
        st.dirtrans(nf, d_gp, d_tmp)
    torch.cuda.synchronize()
    d_gp = torch.zeros(nf * grid.size(), dtype=torch.float64, device="cuda")
    st.invtrans(nf, d_sp, d_gp)
    d_gp[: d_gp.numel() // 7] = float("nan")   # whatever the other bands' rows held is overwritten by the all-gather
    st.invtrans(nf, d_sp, d_gp)
    st.gather_grid(nf, d_gp)
    d_sp2 = torch.zeros_like(d_sp)
    st.dirtrans(nf, d_gp, d_sp2)
    dist.all_reduce(d_sp2)  # every coefficient is produced by exactly one rank
    ok = True
    if rank == 0:
        from oracle import pyoracle as po

        plan = po.OraclePlan(grid.nx(), grid.y(), T, regular=grid.regular, weights=grid.weights())
        want = plan.invtrans(nf, sp, mode=2)
        e1 = H.rel_max(d_gp.cpu().numpy(), want)
        want_sp = plan.dirtrans(nf, want)
        e2 = H.rel_max(d_sp2.cpu().numpy(), want_sp)
        ok = e1 < 1e-12 and e2 < 1e-12
        print(f"DIST_CHECK world={dist.get_world_size()} exchange={exchange} {gridname} T{T} nf={nf}: invtrans rel err {e1:.2e}, dirtrans rel err {e2:.2e} -> {'OK' if ok else 'FAIL'}")
    # the same with every rank holding only its own share (SPTRANS_SHARD_LOCAL_IO), gathered into a global array
    st2 = ShardedTrans(grid, T, local, exchange=exchange, local_io=True)
    nsp, stride = st2.trans.local_sizes()
    l_sp = torch.from_numpy(st2.local_spectra(sp, nf)).cuda()
    l_gp = torch.full((nf * stride,), float("nan"), dtype=torch.float64, device="cuda")
    st2.invtrans(nf, l_sp, l_gp)
    g_gp = torch.full((nf * grid.size(),), float("nan"), dtype=torch.float64, device="cuda")
    st2.gather_grid(nf, l_gp, g_gp)
    l_sp2 = torch.full_like(l_sp, float("nan"))
    st2.dirtrans(nf, l_gp, l_sp2)
    g_sp2 = np.zeros_like(sp)
    st2.scatter_local_spectra(l_sp2.cpu().numpy(), g_sp2, nf)
    t_sp2 = torch.from_numpy(g_sp2).cuda()
    dist.all_reduce(t_sp2)
    if rank == 0:
        e3 = H.rel_max(g_gp.cpu().numpy(), want)
        e4 = H.rel_max(t_sp2.cpu().numpy(), want_sp)
        ok = ok and e3 < 1e-12 and e4 < 1e-12
        print(f"DIST_CHECK local_io: invtrans rel err {e3:.2e}, dirtrans rel err {e4:.2e} -> {'OK' if ok else 'FAIL'}")
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.broadcast(flag, 0)
    ok = bool(flag.item() > 0.5)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
