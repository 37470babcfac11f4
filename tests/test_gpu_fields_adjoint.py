"""GPU tests of the atlas Field-layout entry points and of the adjoint transforms.

Field layouts: a multi-level grid-point Field is (node, level[, component]) with the last index fastest, the transpose
of the IFS-style [field][node] rows; TransIFS packs field = component * nlev + level (trans/ifs/TransIFS.cc:610-667,
:1392-1437, :2113-2137).  The *_field entry points must give exactly (bit for bit) the transposed result of the
raw-pointer entry points, which the parity tests pin against the oracle.

Adjoints: the reference's own adjoint tests (src/tests/trans/test_transgeneral.cc:1591-1818) check the dot-product
identity <A x, y>_grid == <x, A* y>_spec where the grid inner product is the Euclidean sum over points and the SPECTRAL
inner product counts every m > 0 coefficient twice (`adj_value += (m1 > 0 ? 2 * temp : temp)`, :1683-1686, :1790-1793;
TransIFS passes it at 1e-12).  TransLocal itself implements no adjoint (parity unpinned), so this identity -- against the
forward operators the other tests pin -- is the definition (helpers.spectral_dot).  Tolerance: 1e-12 relative to |Ax||y|.
"""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def make(gridname, T):
    import atlas_b200

    grid = atlas_b200.Grid(gridname)
    return grid, atlas_b200.Trans(grid, T, atlas_b200.option.type("b200"))


def rows_to_field(rows, npts, nlev, ncomp):
    """[ncomp * nlev][npts] -> (npts, nlev, ncomp) / (npts, nlev)"""
    a = rows.reshape(ncomp, nlev, npts).transpose(2, 1, 0)
    return np.ascontiguousarray(a if ncomp > 1 else a[:, :, 0])


def field_to_rows(field, npts, nlev, ncomp):
    a = field.reshape(npts, nlev, ncomp).transpose(2, 1, 0)
    return np.ascontiguousarray(a).reshape(-1)


@pytest.mark.parametrize("gridname,T,nlev", [("O48", 47, 5), ("F24", 23, 1), ("O80", 79, 37)])
def test_field_layout_scalar(gridname, T, nlev):
    grid, trans = make(gridname, T)
    npts, nspec2 = grid.size(), trans.nb_spectral_coefficients()
    sp = H.synthetic_spectra(T, nlev).reshape(nspec2, nlev)
    rows = np.full(nlev * npts, np.nan)
    trans.invtrans(nlev, sp.reshape(-1), rows)
    gpf = np.full((npts, nlev), np.nan)
    trans.invtrans_field(sp, gpf)
    assert np.array_equal(gpf, rows_to_field(rows, npts, nlev, 1))
    # direct transform and adjoint of the inverse from the Field layout
    back_rows = np.full(nspec2 * nlev, np.nan)
    trans.dirtrans(nlev, rows, back_rows)
    back = np.full((nspec2, nlev), np.nan)
    trans.dirtrans_field(gpf, back)
    assert np.array_equal(back.reshape(-1), back_rows)
    adj_rows = np.full(nspec2 * nlev, np.nan)
    trans.invtrans_adj(nlev, rows, adj_rows)
    adj = np.full((nspec2, nlev), np.nan)
    trans.invtrans_adj_field(gpf, adj)
    assert np.array_equal(adj.reshape(-1), adj_rows)
    assert trans.last_timings()["repack"] > 0.0


@pytest.mark.parametrize("gridname,T,nlev", [("O48", 47, 3), ("F24", 23, 2)])
def test_field_layout_wind_and_gradient(gridname, T, nlev):
    grid, trans = make(gridname, T)
    npts, nspec2 = grid.size(), trans.nb_spectral_coefficients()
    vor = H.synthetic_spectra(T, nlev, seed=3).reshape(nspec2, nlev)
    div = H.synthetic_spectra(T, nlev, seed=4).reshape(nspec2, nlev)
    rows = np.full(2 * nlev * npts, np.nan)
    trans.invtrans(nlev, vor.reshape(-1), div.reshape(-1), rows)
    wind = np.full((npts, nlev, 2), np.nan)
    trans.invtrans_vordiv2wind_field(vor, div, wind)
    assert np.array_equal(wind, rows_to_field(rows, npts, nlev, 2))
    v_rows, d_rows = np.full(nspec2 * nlev, np.nan), np.full(nspec2 * nlev, np.nan)
    trans.dirtrans(nlev, rows, v_rows, d_rows)
    v_f, d_f = np.full((nspec2, nlev), np.nan), np.full((nspec2, nlev), np.nan)
    trans.dirtrans_wind2vordiv_field(wind, v_f, d_f)
    assert np.array_equal(v_f.reshape(-1), v_rows) and np.array_equal(d_f.reshape(-1), d_rows)
    g_rows = np.full(2 * nlev * npts, np.nan)
    trans.invtrans_grad(nlev, vor.reshape(-1), g_rows)
    grad = np.full((npts, nlev, 2), np.nan)
    trans.invtrans_grad_field(vor, grad)
    assert np.array_equal(grad, rows_to_field(g_rows, npts, nlev, 2))


@pytest.mark.parametrize("gridname,T,nlev", [("O48", 47, 5), ("F24", 23, 3), ("O160", 159, 11)])
def test_field_layout_against_oracle(gridname, T, nlev):
    """Field-layout entry points straight against the CPU oracle (not against the raw-pointer entry points of the same
    library): spectral Field (nspec2, nlev), grid-point Field (npts, nlev), wind / gradient Field (npts, nlev, 2) with
    field = component * nlev + level as TransIFS packs them (trans/ifs/TransIFS.cc:610-667, :1392-1437, :2113-2137)."""
    from oracle import pyoracle as po

    grid, trans = make(gridname, T)
    plan = po.OraclePlan(grid.nx(), grid.y(), T, regular=grid.regular, weights=grid.weights())
    npts, nspec2 = grid.size(), trans.nb_spectral_coefficients()
    sp = H.synthetic_spectra(T, nlev)
    gpf = np.full((npts, nlev), np.nan)
    trans.invtrans_field(sp.reshape(nspec2, nlev), gpf)
    want = plan.invtrans(nlev, sp, mode=2)
    assert H.compute_rms(field_to_rows(gpf, npts, nlev, 1), want) < 1e-13
    spf = np.full((nspec2, nlev), np.nan)
    trans.dirtrans_field(rows_to_field(want, npts, nlev, 1), spf)
    assert H.rel_max(spf.reshape(-1), plan.dirtrans(nlev, want)) < 1e-12
    vor, div = H.synthetic_spectra(T, nlev, seed=7), H.synthetic_spectra(T, nlev, seed=8)
    wind = np.full((npts, nlev, 2), np.nan)
    trans.invtrans_vordiv2wind_field(vor.reshape(nspec2, nlev), div.reshape(nspec2, nlev), wind)
    want_w = plan.invtrans(0, None, nlev, vor, div, mode=2)
    assert H.compute_rms(field_to_rows(wind, npts, nlev, 2), want_w) < 1e-12
    grad = np.full((npts, nlev, 2), np.nan)
    trans.invtrans_grad_field(sp.reshape(nspec2, nlev), grad)
    assert H.compute_rms(field_to_rows(grad, npts, nlev, 2), plan.invtrans_grad(nlev, sp)) < 1e-12


def test_field_layout_device_pointers_and_errors():
    import torch

    from atlas_b200 import _lib

    grid, trans = make("O48", 47)
    T, nlev = 47, 4
    npts, nspec2 = grid.size(), trans.nb_spectral_coefficients()
    sp = H.synthetic_spectra(T, nlev).reshape(nspec2, nlev)
    want = np.full((npts, nlev), np.nan)
    trans.invtrans_field(sp, want)
    d_sp = torch.from_numpy(sp).cuda()
    d_gp = torch.full((npts, nlev), float("nan"), dtype=torch.float64, device="cuda")
    trans.invtrans_field(d_sp, d_gp)
    assert np.array_equal(d_gp.cpu().numpy(), want)
    d_back = torch.full((nspec2, nlev), float("nan"), dtype=torch.float64, device="cuda")
    trans.dirtrans_field(d_gp, d_back)
    back = np.full((nspec2, nlev), np.nan)
    trans.dirtrans_field(want, back)
    assert np.array_equal(d_back.cpu().numpy(), back)
    # device-resident wind / gradient / adjoint Fields: the repack runs in field chunks next to the Fourier kernels
    # (fourier_inv_to_field / fourier_dir_from_field) -- bit-identical to the host-array path (separate transpose pass)
    for gname, TT, nl in (("O48", 47, 5), ("O160", 159, 9), ("F24", 23, 1)):
        grid2, trans2 = make(gname, TT)
        np2, ns2 = grid2.size(), trans2.nb_spectral_coefficients()
        vor = H.synthetic_spectra(TT, nl, seed=3).reshape(ns2, nl)
        div = H.synthetic_spectra(TT, nl, seed=4).reshape(ns2, nl)
        wind_h = np.full((np2, nl, 2), np.nan)
        trans2.invtrans_vordiv2wind_field(vor, div, wind_h)
        d_wind = torch.full((np2, nl, 2), float("nan"), dtype=torch.float64, device="cuda")
        trans2.invtrans_vordiv2wind_field(torch.from_numpy(vor).cuda(), torch.from_numpy(div).cuda(), d_wind)
        assert np.array_equal(d_wind.cpu().numpy(), wind_h), gname
        grad_h = np.full((np2, nl, 2), np.nan)
        trans2.invtrans_grad_field(vor, grad_h)
        d_grad = torch.full((np2, nl, 2), float("nan"), dtype=torch.float64, device="cuda")
        trans2.invtrans_grad_field(torch.from_numpy(vor).cuda(), d_grad)
        assert np.array_equal(d_grad.cpu().numpy(), grad_h), gname
        gp_h = np.full((np2, nl), np.nan)
        trans2.invtrans_field(vor, gp_h)
        d_gp2 = torch.full((np2, nl), float("nan"), dtype=torch.float64, device="cuda")
        trans2.invtrans_field(torch.from_numpy(vor).cuda(), d_gp2)
        assert np.array_equal(d_gp2.cpu().numpy(), gp_h), gname
        if grid2.weights() is not None:
            sp_h = np.full((ns2, nl), np.nan)
            trans2.dirtrans_field(gp_h, sp_h)
            d_sp2 = torch.full((ns2, nl), float("nan"), dtype=torch.float64, device="cuda")
            trans2.dirtrans_field(d_gp2, d_sp2)
            assert np.array_equal(d_sp2.cpu().numpy(), sp_h), gname
        adj_h = np.full((ns2, nl), np.nan)
        trans2.invtrans_adj_field(gp_h, adj_h)
        d_adj = torch.full((ns2, nl), float("nan"), dtype=torch.float64, device="cuda")
        trans2.invtrans_adj_field(d_gp2, d_adj)
        assert np.array_equal(d_adj.cpu().numpy(), adj_h), gname
    with pytest.raises(ValueError):
        trans.invtrans_field(sp, np.zeros((npts, nlev + 1)))
    with pytest.raises(_lib.SptransError):
        _lib.check(_lib.lib.sptrans_invtrans_field(trans._h, 2, None, None))
    _lib.check(_lib.lib.sptrans_invtrans_field(trans._h, 0, None, None))  # zero levels: nothing to do


def _close(lhs, rhs, scale, tol=1e-12):
    return abs(lhs - rhs) <= tol * max(abs(lhs), abs(rhs), scale * 1e-3)


@pytest.mark.parametrize("gridname,T,nvd,nsc", [("F24", 23, 2, 0), ("O48", 47, 3, 2), ("L9", 17, 1, 1), ("O400", 399, 2, 1)])
def test_invtrans_wind_adjoint_identity(gridname, T, nvd, nsc):
    """<invtrans(scalars, vor, div), y> == <(scalars, vor, div), invtrans_adj y>  (TransImpl.h:147-149, :165-166)."""
    grid, trans = make(gridname, T)
    npts = grid.size()
    nspec2 = trans.nb_spectral_coefficients()
    rng = np.random.default_rng(11)
    vor = rng.standard_normal(nspec2 * nvd)
    div = rng.standard_normal(nspec2 * nvd)
    sc = rng.standard_normal(nspec2 * max(nsc, 1))
    y = rng.standard_normal((2 * nvd + nsc) * npts)
    ax = np.full((2 * nvd + nsc) * npts, np.nan)
    if nsc:
        trans.invtrans(nsc, sc, nvd, vor, div, ax)
    else:
        trans.invtrans(nvd, vor, div, ax)
    av, ad, asc = np.full_like(vor, np.nan), np.full_like(div, np.nan), np.full_like(sc, np.nan)
    if nsc:
        trans.invtrans_adj(nsc, y, nvd, av, ad, asc)
    else:
        trans.invtrans_adj(nvd, y, av, ad)
    lhs = float(ax @ y)
    rhs = H.spectral_dot(T, nvd, vor, av) + H.spectral_dot(T, nvd, div, ad) + (H.spectral_dot(T, nsc, sc, asc) if nsc else 0.0)
    # L9 has rows at the poles, where u, v = U, V / cos(89.9999999 deg) (TransLocal.cc:1447-1458): the operator norm
    # carries the factor 5.7e8 and so does the rounding error of both sides (the reference's own wind tolerance is 2e-6)
    tol = 1e-7 if gridname == "L9" else 1e-12
    assert _close(lhs, rhs, np.linalg.norm(ax) * np.linalg.norm(y), tol), (lhs, rhs)


@pytest.mark.parametrize("gridname,T,nf", [("F24", 23, 2), ("O48", 47, 3), ("O400", 399, 2)])
def test_invtrans_grad_adjoint_identity(gridname, T, nf):
    grid, trans = make(gridname, T)
    npts, nspec2 = grid.size(), trans.nb_spectral_coefficients()
    rng = np.random.default_rng(12)
    x = rng.standard_normal(nspec2 * nf)
    y = rng.standard_normal(2 * nf * npts)
    ax = np.full(2 * nf * npts, np.nan)
    trans.invtrans_grad(nf, x, ax)
    ay = np.full_like(x, np.nan)
    trans.invtrans_grad_adj(nf, y, ay)
    lhs, rhs = float(ax @ y), H.spectral_dot(T, nf, x, ay)
    assert _close(lhs, rhs, np.linalg.norm(ax) * np.linalg.norm(y)), (lhs, rhs)


@pytest.mark.parametrize("gridname,T,nf", [("F24", 23, 2), ("O48", 47, 3), ("O160", 159, 5), ("O400", 399, 2)])
def test_dirtrans_adjoint_identity(gridname, T, nf):
    """<dirtrans g, s>_spec == <g, dirtrans_adj s>_grid  (TransImpl::dirtrans_adj, TransImpl.h:63-67; the "direct" leg of
    test_transgeneral.cc:1650-1722)."""
    grid, trans = make(gridname, T)
    npts, nspec2 = grid.size(), trans.nb_spectral_coefficients()
    rng = np.random.default_rng(13)
    g = rng.standard_normal(nf * npts)
    s = rng.standard_normal(nspec2 * nf)
    dg = np.full(nspec2 * nf, np.nan)
    trans.dirtrans(nf, g, dg)
    ds = np.full(nf * npts, np.nan)
    trans.dirtrans_adj(nf, s, ds)
    lhs, rhs = H.spectral_dot(T, nf, dg, s), float(g @ ds)
    assert _close(lhs, rhs, np.linalg.norm(dg) * np.linalg.norm(s)), (lhs, rhs)


@pytest.mark.parametrize("gridname,T,nf", [("F24", 23, 2), ("O48", 47, 3), ("O160", 159, 2)])
def test_dirtrans_wind2vordiv_adjoint_identity(gridname, T, nf):
    """<wind2vordiv(u, v), (z, d)>_spec == <(u, v), wind2vordiv_adj(z, d)>_grid  (TransImpl::dirtrans_wind2vordiv_adj,
    TransImpl.h:69-70; the reference's test_2level_adjoint_test_with_vortdiv, test_transgeneral.cc:1725-1818, with the
    spectral inner product that counts m > 0 twice), raw-pointer and Field layouts."""
    grid, trans = make(gridname, T)
    npts, nspec2 = grid.size(), trans.nb_spectral_coefficients()
    rng = np.random.default_rng(17)
    wind = rng.standard_normal(2 * nf * npts)
    z = rng.standard_normal(nspec2 * nf)
    d = rng.standard_normal(nspec2 * nf)
    vor, div = np.full(nspec2 * nf, np.nan), np.full(nspec2 * nf, np.nan)
    trans.dirtrans(nf, wind, vor, div)
    back = np.full(2 * nf * npts, np.nan)
    trans.dirtrans_wind2vordiv_adj(nf, z, d, back)
    lhs = H.spectral_dot(T, nf, vor, z) + H.spectral_dot(T, nf, div, d)
    rhs = float(wind @ back)
    scale = np.sqrt(np.linalg.norm(vor) ** 2 + np.linalg.norm(div) ** 2) * np.sqrt(np.linalg.norm(z) ** 2 + np.linalg.norm(d) ** 2)
    assert _close(lhs, rhs, scale), (lhs, rhs)
    # Field layout: (nspec2, nlev) spectra, (npts, nlev, 2) wind -- bit-identical to the raw-pointer result
    gpw = np.full((npts, nf, 2), np.nan)
    trans.dirtrans_wind2vordiv_adj_field(z.reshape(nspec2, nf), d.reshape(nspec2, nf), gpw)
    assert np.array_equal(field_to_rows(gpw, npts, nf, 2), back)


@pytest.mark.parametrize("gridname,T", [("O32", 31), ("F24", 23), ("O80", 79)])
def test_legendre_cache_import(gridname, T):
    """Reference-layout Legendre cache (TransLocal.cc:592-647): export -> import round trip leaves the transform bit
    identical; a modified blob is what the transform then uses; a blob of the wrong size is refused like the reference's
    size assertion (TransLocal.cc:612)."""
    from atlas_b200 import _lib

    grid, trans = make(gridname, T)
    nf = 3
    sp = H.synthetic_spectra(T, nf)
    want = np.full(nf * grid.size(), np.nan)
    trans.invtrans(nf, sp, want)
    blob = trans.export_legendre_cache()
    _, other = make(gridname, T)
    other.import_legendre_cache(blob)
    assert np.array_equal(other.export_legendre_cache(), blob)  # export regenerates: independent of the import
    got = np.full_like(want, np.nan)
    other.invtrans(nf, sp, got)
    assert np.array_equal(got, want)
    other.import_legendre_cache(2.0 * blob)
    other.invtrans(nf, sp, got)
    assert np.array_equal(got, 2.0 * want)
    back = np.full_like(sp, np.nan)
    other.dirtrans(nf, want, back)   # the direct transform reads the same table
    ref = np.full_like(sp, np.nan)
    trans.dirtrans(nf, want, ref)
    assert np.array_equal(back, 2.0 * ref)
    with pytest.raises(_lib.SptransError):
        other.import_legendre_cache(blob[:-8])


# ---- point sets: TransLocal's unstructured path (TransLocal.cc:1289-1392) -------------------------------------------
def _points_oracle(T, trunc, nf, spec, lon_deg, lat_deg, nb_uv=0):
    """Literal restatement of invtrans_unstructured (:1289-1392) with the oracle's pinned Legendre functions:
    per point Pnm(lat) -> sum over n for every m <= trunc -> dot product with (1, 0, 2cos, -2sin, ...) -> 1/cos for wind.
    `spec` is [m][n][re/im][fld] at truncation `trunc`."""
    from oracle import pyoracle as po

    sp = spec.reshape(-1, 2, nf)
    out = np.zeros((nf, lon_deg.size))
    for ip, (lo, la) in enumerate(zip(lon_deg, lat_deg)):
        lon, lat = np.deg2rad(lo), np.deg2rad(la)
        leg = po.legendre_lat(trunc, lat)
        acc = np.zeros(nf)
        for m in range(trunc + 1):
            off = (2 * trunc + 3 - m) * m // 2
            ns = trunc - m + 1
            c = leg[off:off + ns] @ sp[off:off + ns, 0, :] + 1j * (leg[off:off + ns] @ sp[off:off + ns, 1, :])
            if m == 0:
                acc += c.real
            else:
                acc += 2.0 * np.cos(m * lon) * c.real - 2.0 * np.sin(m * lon) * c.imag
        acc[:nb_uv] /= np.cos(lat)
        out[:, ip] = acc
    return out.reshape(-1)


def _some_points(rng, n):
    lon = rng.uniform(-180.0, 360.0, n)
    lat = rng.uniform(-89.0, 89.0, n)
    # shared latitudes, mirrored latitudes, the equator, a pole
    lat[1] = lat[0]
    lat[2] = -lat[0]
    lat[3] = 0.0
    lat[4] = 90.0
    lon[4] = 12.5
    return lon, lat


@pytest.mark.parametrize("T,nf,npt", [(21, 3, 40), (63, 5, 33), (159, 2, 17)])
def test_unstructured_points_scalar(T, nf, npt):
    import atlas_b200

    rng = np.random.default_rng(3)
    lon, lat = _some_points(rng, npt)
    trans = atlas_b200.Trans(atlas_b200.UnstructuredGrid(lon, lat), T)
    assert trans.nb_gridpoints() == npt
    sp = H.synthetic_spectra(T, nf)
    gp = np.full(nf * npt, np.nan)
    trans.invtrans(nf, sp, gp)
    want = _points_oracle(T, T, nf, sp, lon, lat)
    assert H.rel_max(gp, want) < 1e-12 and H.compute_rms(gp, want) < 1e-13
    # the same values as the structured transform on the points of a regular grid (no zonal truncation towards the
    # poles there), except that the point path keeps m == T
    grid = atlas_b200.Grid("F16")
    glon, glat = grid.lonlat()
    tp = atlas_b200.Trans(atlas_b200.UnstructuredGrid(glon, glat), 15)
    tg = atlas_b200.Trans(grid, 15)
    s15 = H.synthetic_spectra(15, 2)
    s15.reshape(-1, 2, 2)[-1] = 0.0   # zero the m == T coefficient: both paths then agree
    a, b = np.full(2 * grid.size(), np.nan), np.full(2 * grid.size(), np.nan)
    tp.invtrans(2, s15, a)
    tg.invtrans(2, s15, b)
    assert H.rel_max(a, b) < 1e-12


def test_unstructured_points_wind_and_errors():
    import atlas_b200
    from atlas_b200 import _lib
    from oracle import pyoracle as po

    T, nvd, nsc, npt = 31, 2, 1, 25
    rng = np.random.default_rng(4)
    lon = rng.uniform(0.0, 360.0, npt)
    lat = rng.uniform(-80.0, 80.0, npt)
    lat[1] = -lat[0]
    trans = atlas_b200.Trans(atlas_b200.UnstructuredGrid(lon, lat), T)
    vor, div, sc = H.synthetic_spectra(T, nvd, seed=5), H.synthetic_spectra(T, nvd, seed=6), H.synthetic_spectra(T, nsc, seed=7)
    gp = np.full((2 * nvd + nsc) * npt, np.nan)
    trans.invtrans(nsc, sc, nvd, vor, div, gp)
    # reference :1523-1597: extend to T+1, vd2uv at T+1, fields [U.. | V.. | scalars] -> unstructured transform at T+1
    Te = T + 1
    ext = lambda a, n: np.ascontiguousarray(_extend(a, T, n))
    U, V = po.vd2uv(Te, nvd, ext(vor, nvd), ext(div, nvd))
    allsp = np.concatenate([U.reshape(-1, 2, nvd), V.reshape(-1, 2, nvd), ext(sc, nsc).reshape(-1, 2, nsc)], axis=2)
    want = _points_oracle(T, Te, 2 * nvd + nsc, np.ascontiguousarray(allsp).reshape(-1), lon, lat, nb_uv=2 * nvd)
    assert H.compute_rms(gp, want) < 1e-12
    with pytest.raises(_lib.NotImplementedInBackend):
        trans.dirtrans(1, np.zeros(npt), np.zeros(trans.nb_spectral_coefficients()))
    with pytest.raises(_lib.NotImplementedInBackend):
        trans.invtrans_adj(1, np.zeros(npt), np.zeros(trans.nb_spectral_coefficients()))


def _extend(sp, T, nf):
    """extend_truncation (TransLocal.cc:1496-1519): [m][n] at T -> T+1 with zeros at n == T+1 and m == T+1."""
    a = sp.reshape(-1, 2, nf)
    Te = T + 1
    out = np.zeros(((Te + 1) * (Te + 2) // 2, 2, nf))
    k = ke = 0
    for m in range(T + 1):
        cnt = T - m + 1
        out[ke:ke + cnt] = a[k:k + cnt]
        k += cnt
        ke += cnt + 1
    return out.reshape(-1)
