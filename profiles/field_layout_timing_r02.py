"""Field-layout entry points on device-resident Fields at TCo1279 L137 against the raw-row entry points (ms, median of 5):
what the level-fastest <-> row repack costs on top of the transform.  python profiles/field_layout_timing_r02.py"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import atlas_b200
T, nlev = 1279, 137
grid = atlas_b200.Grid("O1280")
tr = atlas_b200.Trans(grid, T)
tr.set_stream(torch.cuda.current_stream().cuda_stream)
npts, nspec2 = grid.size(), tr.nb_spectral_coefficients()
kw = dict(dtype=torch.float64, device="cuda")
sp = torch.randn(nspec2, nlev, **kw)
rows = torch.empty(nlev, npts, **kw)
fld = torch.empty(npts, nlev, **kw)
sp2 = torch.empty(nspec2, nlev, **kw)
def timed(fn):
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return round(float(np.median(ts[2:])), 3)
out = {"invtrans_rows": timed(lambda: tr.invtrans(nlev, sp, rows)), "invtrans_field": timed(lambda: tr.invtrans_field(sp, fld)),
       "dirtrans_rows": timed(lambda: tr.dirtrans(nlev, rows, sp2)), "dirtrans_field": timed(lambda: tr.dirtrans_field(fld, sp2))}
chk = torch.empty(nlev, npts, **kw)
tr.invtrans(nlev, sp, chk)
tr.invtrans_field(sp, fld)
out["field_equals_rows_transposed"] = bool(torch.equal(fld, chk.t().contiguous()))
print(json.dumps(out))
