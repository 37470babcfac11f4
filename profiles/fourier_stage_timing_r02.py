"""Fourier-stage timing at TCo1279 L137 (stage-level API), for A/B experiments: prints ms inverse / direct (median of 7)."""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import atlas_b200
from atlas_b200 import _lib
T, nf = 1279, 137
grid = atlas_b200.Grid("O1280")
tr = atlas_b200.Trans(grid, T)
tr.set_stream(torch.cuda.current_stream().cuda_stream)
fb = torch.randn(tr.fourier_elems_per_field() * 2 * nf, dtype=torch.float64, device="cuda")
gp = torch.empty(nf * grid.size(), dtype=torch.float64, device="cuda")
def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
inv = [timed(lambda: tr.invtrans_fourier(nf, T - 1, fb, gp)) for _ in range(9)][2:]
dr = [timed(lambda: tr.dirtrans_fourier(nf, gp, fb)) for _ in range(9)][2:]
print(json.dumps({"tag": os.environ.get("TAG", ""), "inv_ms": round(float(np.median(inv)), 3), "dir_ms": round(float(np.median(dr)), 3)}))
