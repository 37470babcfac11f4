"""CPU tests (no GPU): the oracle against every golden vector the reference offers for this path.

Pins the oracle before it is trusted as the checker of the CUDA path:
  * Gaussian latitudes vs the reference's tabulated values (tests/golden/gaussian_latitudes_N*.npy)
  * Legendre polynomials bit-for-bit vs the unmodified reference source (oracle/_ref or the golden file)
  * FFT vs the literal DFT
  * inverse transform vs the reference's closed-form harmonics with the reference's own tolerances
    (test_transgeneral.cc:534-538: 1e-13 scalar; :833-838: 2e-6 wind)
  * literal ("as written") vs fast (blocked/OpenMP/FFT) oracle paths
"""
import os

import numpy as np
import pytest

import helpers as H
from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def octahedral(N):
    lat, w = po.gaussian_quadrature(N)
    nx = np.array([20 + 4 * j for j in range(N)] + [20 + 4 * j for j in range(N - 1, -1, -1)], dtype=np.int32)
    return nx, lat, w


def regular_gaussian(N):
    lat, w = po.gaussian_quadrature(N)
    return np.full(2 * N, 4 * N, dtype=np.int32), lat, w


def lonlat_grid(N):
    lat = 90.0 - 180.0 * np.arange(2 * N + 1) / (2 * N)
    lat[N] = 0.0
    return np.full(2 * N + 1, 4 * N, dtype=np.int32), lat, None


@pytest.mark.parametrize("N", [32, 400, 1280])
def test_gaussian_latitudes_match_reference_tables(N):
    gold = np.load(os.path.join(GOLD, f"gaussian_latitudes_N{N}.npy"))
    lat, w = po.gaussian_quadrature(N)
    assert np.abs(lat[:N] - gold).max() < 5e-12  # tables carry 12 decimals (N1280 row 0 is off by 3e-12)
    assert np.allclose(lat[N:], -lat[:N][::-1], rtol=0, atol=0)
    assert abs(w.sum() - 1.0) < 1e-13  # atlas normalisation: weights over all 2N rows sum to 1


def test_legendre_matches_reference_golden_bitwise():
    g = np.load(os.path.join(GOLD, "legendre_ref_T33.npz"))
    trc = int(g["trc"])
    for lat, want in zip(g["lats"], g["legpol"]):
        got = po.legendre_lat(trc, float(lat))
        assert np.array_equal(got, want)


@pytest.mark.skipif(po.load_ref_legendre() is None, reason="oracle/_ref not built (reference tree absent)")
@pytest.mark.parametrize("trc,lat_deg", [(64, 60.0), (161, 3.3), (400, 89.9999999), (400, 0.2)])
def test_legendre_matches_reference_inplace_bitwise(trc, lat_deg):
    lat = float(np.deg2rad(lat_deg))
    assert np.array_equal(po.legendre_lat(trc, lat), po.legendre_lat(trc, lat, ref=True))


@pytest.mark.skipif(po.load_ref_legendre() is None, reason="oracle/_ref not built (reference tree absent)")
def test_legendre_tables_match_reference_inplace_bitwise():
    import ctypes as C

    nx, lat, w = octahedral(16)
    T = 15
    plan = po.OraclePlan(nx, lat, T)
    sym, asym, sb, ab = plan.tables()
    ref = po.load_ref_legendre()
    rs = np.zeros_like(sym)
    ra = np.zeros_like(asym)
    lats = np.deg2rad(lat[:16])
    ref.ref_legendre_tables(T + 1, 16, lats.ctypes.data_as(C.POINTER(C.c_double)), rs.ctypes.data_as(C.POINTER(C.c_double)),
                            ra.ctypes.data_as(C.POINTER(C.c_double)), sb.ctypes.data_as(C.POINTER(C.c_size_t)),
                            ab.ctypes.data_as(C.POINTER(C.c_size_t)))
    assert np.array_equal(sym, rs) and np.array_equal(asym, ra)


@pytest.mark.parametrize("n", [4, 20, 24, 36, 100, 144, 332, 1616, 2052, 5132])
def test_fft_c2r_matches_literal_dft(n):
    rng = np.random.default_rng(n)
    spec = rng.standard_normal(n // 2 + 1) + 1j * rng.standard_normal(n // 2 + 1)
    spec[0] = spec[0].real
    a = po.c2r(n, spec, naive=False)
    b = po.c2r(n, spec, naive=True)
    assert H.rel_max(a, b) < 2e-14
    # r2c is the adjoint/inverse pair
    back = po.r2c(n, a) / n
    m = (n - 1) // 2
    assert np.abs(back[1:m + 1] - spec[1:m + 1]).max() < 1e-13


def test_fourier_truncation_octahedral_cubic():
    # O1280 / T1279: cubic rule; every m < = T is resolved at the equator, few at the pole
    nx, lat, w = octahedral(1280)
    plan_nlat0 = []
    for m_probe, j in ((0, 0), (1279, 1279)):
        ft = po.lib().orc_fourier_truncation(1279, int(nx[j]), 5136, 2560, float(np.deg2rad(lat[j])), 0)
        plan_nlat0.append(ft)
    assert plan_nlat0[0] == 8 and plan_nlat0[1] == 1279


CASES = [("F32", regular_gaussian(32), 31, True), ("O32", octahedral(32), 31, False), ("O64", octahedral(64), 63, False),
         ("L3", lonlat_grid(3), 5, True), ("L9", lonlat_grid(9), 17, True)]


@pytest.mark.parametrize("name,grid,T,regular", CASES)
def test_invtrans_scalar_vs_closed_form_harmonics(name, grid, T, regular):
    """test_transgeneral.cc:493-643 (and the pole case :956-1140): unit coefficient -> closed-form harmonic."""
    nx, lat, w = grid
    plan = po.OraclePlan(nx, lat, T, regular=regular, weights=w)
    lon, latp = H.grid_lonlat(nx, np.clip(lat, -89.9999999, 89.9999999) if name.startswith("L") else lat)
    ncoef2 = (T + 1) * (T + 2)
    worst = 0.0
    ncases = 0
    for m in range(0, min(T, 45) + 1):
        for n in range(m, min(T, 45) + 1):
            if not H.has_closed_form(n, m) or m >= T:  # m == T is dropped by the scalar path (TransLocal.cc:982)
                continue
            for imag in (0, 1):
                if m == 0 and imag == 1:
                    continue
                sp = np.zeros(ncoef2)
                sp[H.spec_index(T, m, n, imag)] = 1.0
                got = plan.invtrans(1, sp, mode=2)
                want = H.analytic_harmonic(n, m, imag, lon, latp)
                mask = H.expected_zonal_mask(T, nx, lat, regular, m, po.lib().orc_fourier_truncation)
                want = np.where(mask, want, 0.0)
                worst = max(worst, H.compute_rms(got, want))
                ncases += 1
    assert ncases > 10
    assert worst < 1e-13, worst


def test_invtrans_wind_vs_closed_form():
    """vorticity / divergence unit coefficients -> u, v (test_transgeneral.cc:540-600, tolerance 2e-6)."""
    nx, lat, w = regular_gaussian(32)
    T = 31
    plan = po.OraclePlan(nx, lat, T, regular=True, weights=w)
    lon, latp = H.grid_lonlat(nx, lat)
    npts = plan.npts
    ncoef2 = (T + 1) * (T + 2)
    for (n, m) in ((1, 0), (1, 1)):
        for imag in (0, 1):
            if m == 0 and imag == 1:
                continue
            for var_in in (0, 1):
                vor = np.zeros(ncoef2)
                div = np.zeros(ncoef2)
                (vor if var_in == 0 else div)[H.spec_index(T, m, n, imag)] = 1.0
                gp = plan.invtrans(0, None, 1, vor, div, mode=2)
                for var_out in (0, 1):
                    want = H.analytic_wind(n, m, imag, lon, latp, var_in, var_out)
                    got = gp[var_out * npts:(var_out + 1) * npts]
                    assert H.compute_rms(got, want) < 2e-6 or np.abs(want).max() == 0 and np.abs(got).max() < 1e-3


def test_literal_and_fast_paths_agree():
    nx, lat, w = octahedral(32)
    T = 31
    nf = 4
    plan = po.OraclePlan(nx, lat, T, weights=w)
    sp = H.synthetic_spectra(T, nf)
    a = plan.invtrans(nf, sp, mode=0)   # naive GEMM + naive DFT, single thread: "reference as written"
    b = plan.invtrans(nf, sp, mode=2)   # blocked GEMM + FFT + OpenMP
    assert H.rel_max(b, a) < 1e-14
    vor = H.synthetic_spectra(T, 2, seed=7)
    div = H.synthetic_spectra(T, 2, seed=8)
    a = plan.invtrans(nf, sp, 2, vor, div, mode=0)
    b = plan.invtrans(nf, sp, 2, vor, div, mode=2)
    assert H.rel_max(b, a) < 1e-14


def test_dirtrans_roundtrip_regular_grid():
    """dirtrans(invtrans(x)) == x on a regular Gaussian grid (exact quadrature) except the m == T column
    the scalar inverse drops; semantics of test_transgeneral.cc:1494-1585 (test_trans_levels)."""
    nx, lat, w = regular_gaussian(24)
    T = 23
    nf = 3
    plan = po.OraclePlan(nx, lat, T, regular=True, weights=w)
    sp = H.synthetic_spectra(T, nf)
    gp = plan.invtrans(nf, sp, mode=2)
    back = plan.dirtrans(nf, gp)
    sp_expect = sp.copy().reshape(-1, 2, nf)
    sp_expect[-1] = 0.0  # (m=T, n=T)
    assert np.abs(back - sp_expect.reshape(-1)).max() < 1e-13


def test_vd2uv_against_merged_inverse_definition():
    """U,V from vd2uv at T+1 are what the merged inverse feeds to the Legendre stage: check linearity
    and the (1,0)/(1,1) closed forms through the full inverse instead (above); here only shape/zero rules."""
    T = 10
    nf = 2
    rng = np.random.default_rng(0)
    vor = rng.standard_normal((T + 1) * (T + 2) * nf)
    div = rng.standard_normal((T + 1) * (T + 2) * nf)
    U, V = po.vd2uv(T, nf, vor, div)
    U2, V2 = po.vd2uv(T, nf, 2 * vor, 2 * div)
    assert np.allclose(U2, 2 * U, rtol=1e-14, atol=0) and np.allclose(V2, 2 * V, rtol=1e-14, atol=0)
    # imaginary parts of m = 0 stay zero
    for n in range(T + 1):
        for f in range(nf):
            assert U[H.spec_index(T, 0, n, 1, nf, f)] == 0.0


def test_wind_direct_transform_round_trip_and_gradient_closed_form():
    """dirtrans(wind) and invtrans_grad are not in TransLocal (parity unpinned): the oracle's definitions are checked
    by round trip through the reference's own vd2uv + inverse, and against a closed-form gradient."""
    nx, lat, w = regular_gaussian(24)
    T, nf = 23, 2
    plan = po.OraclePlan(nx, lat, T, regular=True, weights=w)
    vor = H.synthetic_spectra(T, nf, seed=5)
    div = H.synthetic_spectra(T, nf, seed=6)
    vor.reshape(-1, 2, nf)[0] = 0.0
    div.reshape(-1, 2, nf)[0] = 0.0
    wind = plan.invtrans(0, None, nf, vor, div, mode=2)
    v2, d2 = plan.dirtrans_wind(nf, wind)
    assert H.rel_max(v2, vor) < 1e-12 and H.rel_max(d2, div) < 1e-12
    sp1 = np.zeros((T + 1) * (T + 2))
    sp1[H.spec_index(T, 1, 2, 0)] = 1.0
    g1 = plan.invtrans_grad(1, sp1).reshape(2, -1)
    lon, latp = H.grid_lonlat(nx, lat)
    s, c = np.sin(latp), np.cos(latp)
    a = H.EARTH_RADIUS
    assert H.compute_rms(g1[0], -np.sqrt(7.5) * s * 2 * np.sin(lon) / a) < 1e-13
    assert H.compute_rms(g1[1], np.sqrt(7.5) * (c * c - s * s) * 2 * np.cos(lon) / a) < 1e-13


def test_unstructured_path_agrees_with_structured_path_and_closed_forms():
    """orc_invtrans_unstructured (restatement of TransLocal.cc:1289-1392): (a) on the points of a regular Gaussian grid
    -- where the structured path applies no zonal truncation -- it reproduces the structured oracle, provided the
    m == T coefficient is zero (the structured scalar path drops that column, :982, the unstructured one keeps it, :1331);
    (b) the reference's closed-form harmonics at scattered points (test_transgeneral.cc:80-374, used by its own
    unstructured test :1144-1336 with tolerance 1e-13); (c) wind: the merged T+1 spectra of :1556-1589 give the
    structured wind fields."""
    N, T, nf = 16, 15, 3
    lat, w = po.gaussian_quadrature(N)
    nx = np.full(2 * N, 4 * N, dtype=np.int32)
    plan = po.OraclePlan(nx, lat, T, regular=True, weights=w)
    lon_g, lat_g = H.grid_lonlat(nx, lat)
    sp = H.synthetic_spectra(T, nf)
    sp.reshape(-1, 2, nf)[-1] = 0.0
    want = plan.invtrans(nf, sp, mode=0)
    got = po.invtrans_unstructured(T, nf, 0, sp, np.rad2deg(lon_g), np.rad2deg(lat_g))
    assert H.rel_max(got, want) < 1e-13
    # (b) closed forms at scattered points, including the m == T sectoral harmonic the structured path cannot see
    rng = np.random.default_rng(9)
    lon = rng.uniform(-180.0, 360.0, 50)
    latp = rng.uniform(-90.0, 90.0, 50)
    for (n, m, imag) in [(0, 0, 0), (1, 0, 0), (2, 1, 1), (3, 2, 0), (3, 3, 1), (T, T, 0)]:
        s1 = np.zeros((T + 1) * (T + 2))
        s1[H.spec_index(T, m, n, imag)] = 1.0
        g = po.invtrans_unstructured(T, 1, 0, s1, lon, latp)
        ref = H.analytic_harmonic(n, m, imag, np.deg2rad(lon), np.deg2rad(latp))
        assert H.compute_rms(g, ref) < 1e-13, (n, m, imag)
    # (c) wind
    nvd = 2
    vor, div = H.synthetic_spectra(T, nvd, seed=5), H.synthetic_spectra(T, nvd, seed=6)
    wantw = plan.invtrans(0, None, nvd, vor, div, mode=0)
    Te = T + 1
    import ctypes

    dp = ctypes.POINTER(ctypes.c_double)
    ve, de = np.zeros((Te + 1) * (Te + 2) * nvd), np.zeros((Te + 1) * (Te + 2) * nvd)
    po.lib().orc_extend_truncation(T, nvd, vor.ctypes.data_as(dp), ve.ctypes.data_as(dp))
    po.lib().orc_extend_truncation(T, nvd, div.ctypes.data_as(dp), de.ctypes.data_as(dp))
    U, V = po.vd2uv(Te, nvd, ve, de)
    allsp = np.ascontiguousarray(np.concatenate([U.reshape(-1, 2, nvd), V.reshape(-1, 2, nvd)], axis=2)).reshape(-1)
    gotw = po.invtrans_unstructured(Te, 2 * nvd, nvd, allsp, np.rad2deg(lon_g), np.rad2deg(lat_g))
    assert H.compute_rms(gotw, wantw) < 1e-13


def test_gpu_test_point_checker_is_pinned_to_the_oracle():
    """tests/test_gpu_fields_adjoint.py checks the CUDA point-set path against a literal NumPy transcription of
    TransLocal.cc:1289-1392; that transcription is held to the C oracle here (no GPU needed)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("gpu_fields", os.path.join(os.path.dirname(__file__), "test_gpu_fields_adjoint.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(2)
    lon, lat = mod._some_points(rng, 12)
    T, nf = 21, 3
    sp = H.synthetic_spectra(T, nf)
    a = mod._points_oracle(T, T, nf, sp, lon, lat)
    b = po.invtrans_unstructured(T, nf, 0, sp, lon, lat)
    assert H.rel_max(a, b) < 1e-14
    a = mod._points_oracle(T, T, nf, sp, lon[5:], lat[5:], nb_uv=2)   # (without the pole point: 1 / cos(90 deg))
    b = po.invtrans_unstructured(T, nf, 1, sp, lon[5:], lat[5:])
    assert H.rel_max(a, b) < 1e-14
