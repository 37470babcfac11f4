"""Stage times of the 8-rank sharded transform with all 8 ranks emulated on ONE device (peer regions = local memory):
what the per-rank kernels cost without NVLink in the way.  Compare with stage_ms_per_rank of the real 8-GPU run."""
import ctypes as C, sys, os, json
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import helpers as H, atlas_b200
from atlas_b200 import _lib
from atlas_b200.trans import _ptr
lib = _lib.lib
R, T, nf = 8, 1279, 137
grid = atlas_b200.Grid("O1280")
plans = [atlas_b200.Trans(grid, T, rank=r, nranks=R, local_io=True) for r in range(R)]
stream = torch.cuda.current_stream().cuda_stream
regions = (C.c_void_p * R)()
for r, t in enumerate(plans):
    t.set_stream(stream)
    _lib.check(lib.sptrans_peer_alloc(t._h, nf, None))
    reg = C.c_void_p(); _lib.check(lib.sptrans_peer_region(t._h, C.byref(reg), None)); regions[r] = reg
for t in plans:
    _lib.check(lib.sptrans_peer_attach_ptrs(t._h, R, regions))
d_sp, d_gp = [], []
for t in plans:
    nsp, stride = t.local_sizes()
    d_sp.append(torch.randn(nsp * nf, dtype=torch.float64, device="cuda"))
    d_gp.append(torch.zeros(nf * stride, dtype=torch.float64, device="cuda"))
def buf(t):
    b = C.c_void_p(); _lib.check(lib.sptrans_peer_buffer(t._h, C.byref(b))); return b
def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
res = {"inv_legendre": [], "inv_fourier": [], "dir_fourier_push": [], "dir_legendre": []}
NIT = int(os.environ.get('NIT', '3'))
acc = {k: [] for k in res}
for it in range(NIT):
    il = [timed(lambda t=t, a=a: _lib.check(lib.sptrans_invtrans_legendre_peers(t._h, nf, _ptr(a)))) for t, a in zip(plans, d_sp)]
    iff = []
    for t, g in zip(plans, d_gp):
        iff.append(timed(lambda t=t, g=g: _lib.check(lib.sptrans_invtrans_fourier(t._h, nf, T - 1, buf(t), _ptr(g), 0))))
        _lib.check(lib.sptrans_peer_advance(t._h))
    pull = os.environ.get("SPTRANS_DIR_PULL", "0") != "0"
    if pull:
        df = [timed(lambda t=t, g=g: _lib.check(lib.sptrans_dirtrans_fourier_local(t._h, nf, _ptr(g)))) for t, g in zip(plans, d_gp)]
        dl = [timed(lambda t=t, a=a: _lib.check(lib.sptrans_dirtrans_legendre_pull(t._h, nf, _ptr(a)))) for t, a in zip(plans, d_sp)]
        for t in plans:
            _lib.check(lib.sptrans_peer_advance(t._h))
    else:
        df = [timed(lambda t=t, g=g: _lib.check(lib.sptrans_dirtrans_fourier_peers(t._h, nf, _ptr(g)))) for t, g in zip(plans, d_gp)]
        dl = []
        for t, a in zip(plans, d_sp):
            dl.append(timed(lambda t=t, a=a: _lib.check(lib.sptrans_dirtrans_legendre(t._h, nf, buf(t), _ptr(a)))))
            _lib.check(lib.sptrans_peer_advance(t._h))
    if it >= 1:
        for k, v in (("inv_legendre", il), ("inv_fourier", iff), ("dir_fourier_push", df), ("dir_legendre", dl)):
            acc[k].append(v)
res = {k: list(np.median(np.array(v), axis=0)) for k, v in acc.items()}   # median over the iterations after the first
print(json.dumps({"tag": os.environ.get("TAG", ""), **{k: [round(float(x), 3) for x in v] for k, v in res.items()}}))
