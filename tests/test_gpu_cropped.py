"""Cropped (regional) structured grids on the FFT path (TransLocal.cc:371-531, :1034-1079, :1180-1187).

The reference transforms a regional grid that is a cropping of a global grid with the GLOBAL grid's Legendre functions,
per-latitude zonal truncation and row FFTs, and copies out the crop's longitudes (jlonMin with wrap-around).  These tests
reproduce the reference's own fixtures -- test_trans_domain (O64 cropped to lon [-5, 5] x lat [-2.5, 0], T63,
src/tests/trans/test_transgeneral.cc:751-954) and test_trans_southpole (L9 cropped to the southern / northern hemisphere
incl. the pole, T8, :1144-1333) -- on the CUDA path:
  * against the CPU oracle's structured path: its global inverse transform restricted to the crop's points by the
    reference's copy-out rule (1e-13 rms scalar, 1e-12 wind),
  * against the closed-form harmonics the reference checks (1e-13 scalar, 2e-6 wind), every (m, n) with a closed form.
"""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

CASES = [("O64", 63, (-5.0, 5.0, -2.5, 0.0)), ("L9", 8, (0.0, 10.0, -90.0, -10.0)), ("L9", 8, (0.0, 10.0, 10.0, 90.0)),
         ("O48", 47, (300.0, 420.0, -35.0, 62.0)), ("F24", 23, (10.0, 200.0, -80.0, 10.0)),
         ("L9", 17, (350.0, 380.0, -30.0, 30.0)),    # across the equator row of a grid with an odd number of rows, wrapping in longitude
         ("O160", 159, (0.0, 359.9, 40.0, 90.0))]    # a polar cap: whole rows, northern hemisphere only


def make(gridname, T, box):
    import atlas_b200
    from oracle import pyoracle as po

    g = atlas_b200.Grid(gridname)
    crop = atlas_b200.CroppedGrid(g, *box)
    trans = atlas_b200.Trans(crop, T)
    plan = po.OraclePlan(g.nx(), g.y(), T, regular=g.regular, weights=g.weights())
    return g, crop, trans, plan


@pytest.mark.parametrize("gridname,T,box", CASES)
def test_cropped_invtrans_matches_oracle(gridname, T, box):
    g, crop, trans, plan = make(gridname, T, box)
    idx = crop.global_indices()
    assert trans.nb_gridpoints() == crop.size() == idx.size
    nf = 5
    sp = H.synthetic_spectra(T, nf)
    gp = np.full(nf * crop.size(), np.nan)
    trans.invtrans(nf, sp, gp)
    want = plan.invtrans(nf, sp, mode=2).reshape(nf, -1)[:, idx]
    assert H.compute_rms(gp, want.reshape(-1)) < 1e-13 and H.rel_max(gp, want.reshape(-1)) < 1e-12
    # vor/div + scalars at T+1 (u | v | scalars), wind scaled by 1 / cos(lat)
    nvd, nsc = 2, 1
    vor, div, sc = H.synthetic_spectra(T, nvd, seed=3), H.synthetic_spectra(T, nvd, seed=4), H.synthetic_spectra(T, nsc, seed=5)
    gw = np.full((2 * nvd + nsc) * crop.size(), np.nan)
    trans.invtrans(nsc, sc, nvd, vor, div, gw)
    ww = plan.invtrans(nsc, sc, nvd, vor, div, mode=2).reshape(2 * nvd + nsc, -1)[:, idx]
    # a crop with a pole row: u, v = U, V / cos(89.9999999 deg) (TransLocal.cc:1447-1458) carries the factor 5.7e8 and so does
    # the rounding error of both sides (the reference's own wind tolerance is 2e-6)
    has_pole = np.abs(crop.y()).max() > 89.0
    assert H.compute_rms(gw, ww.reshape(-1)) < (1e-7 if has_pole else 1e-12)
    # multi-level Field layout (npts, nlev) through the same plan
    gpf = np.full((crop.size(), nf), np.nan)
    trans.invtrans_field(sp.reshape(-1, nf), gpf)
    assert np.array_equal(gpf.T.reshape(-1), gp)
    # device arrays in place
    import torch

    d_gp = torch.full((nf * crop.size(),), float("nan"), dtype=torch.float64, device="cuda")
    trans.invtrans(nf, torch.from_numpy(sp).cuda(), d_gp)
    assert np.array_equal(d_gp.cpu().numpy(), gp)
    # like TransLocal, no direct transform on a regional grid
    from atlas_b200 import _lib

    with pytest.raises(_lib.NotImplementedInBackend):
        trans.dirtrans(nf, gp, np.zeros_like(sp))


@pytest.mark.parametrize("gridname,T,box", CASES[:3])
def test_cropped_invtrans_vs_closed_form(gridname, T, box):
    """The acceptance loop of test_trans_domain / test_trans_southpole: unit coefficients, analytic fields on the crop."""
    from oracle import pyoracle as po

    g, crop, trans, plan = make(gridname, T, box)
    idx = crop.global_indices()
    lon_g, lat_g = H.grid_lonlat(g.nx(), np.clip(g.y(), -89.9999999, 89.9999999))
    lon, lat = lon_g[idx], lat_g[idx]
    cases = [(m, n, im) for m in range(min(T, 45) + 1) for n in range(m, min(T, 45) + 1) if H.has_closed_form(n, m)
             for im in (0, 1) if not (m == 0 and im == 1) and m < T]
    nf = len(cases)
    sp = np.zeros((T + 1) * (T + 2) * nf)
    for f, (m, n, im) in enumerate(cases):
        sp[H.spec_index(T, m, n, im, nf, f)] = 1.0
    gp = np.full(nf * crop.size(), np.nan)
    trans.invtrans(nf, sp, gp)
    gp = gp.reshape(nf, -1)
    worst = 0.0
    for f, (m, n, im) in enumerate(cases):
        want = H.analytic_harmonic(n, m, im, lon, lat)
        mask = H.expected_zonal_mask(T, g.nx(), g.y(), g.regular, m, po.lib().orc_fourier_truncation)[idx]
        worst = max(worst, H.compute_rms(gp[f], np.where(mask, want, 0.0)))
    assert worst < 1e-13, worst
    # wind from unit vorticity / divergence coefficients (n, m) in {(1,0), (1,1)}: 2e-6 like the reference
    poleless = np.abs(g.y()[crop.jlat_min:crop.jlat_min + crop.ny()]).max() < 89.0
    for var_in in (0, 1):
        for (n, m, im) in ((1, 0, 0), (1, 1, 0), (1, 1, 1)):
            vor = np.zeros((T + 1) * (T + 2))
            div = np.zeros_like(vor)
            (vor if var_in == 0 else div)[H.spec_index(T, m, n, im)] = 1.0
            uv = np.full(2 * crop.size(), np.nan)
            trans.invtrans(1, vor, div, uv)
            uv = uv.reshape(2, -1)
            for var_out in (0, 1):
                want = H.analytic_wind(n, m, im, lon, lat, var_in, var_out)
                if poleless:
                    assert H.compute_rms(uv[var_out], want) < 2e-6
