"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, through the C ABI, against the CPU
oracle on identical inputs, against the reference's closed-form harmonics, and -- at BASELINE sizes --
through size-independent properties.

Stated fp64 tolerances (the reference's own ladder is 1e-13 rel-RMS scalar / 2e-6 wind,
test_transgeneral.cc:534-538,833-838):
    scalar fields  : compute_rms <= 1e-13  and  max-norm relative error <= 1e-12
    wind fields    : compute_rms <= 1e-12 vs the oracle (2e-6 vs closed forms, as in the reference)
    spectra        : max-norm relative error <= 1e-12
Legendre tables are checked BIT-FOR-BIT.
"""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

TOL_RMS = 1e-13
TOL_MAX = 1e-12


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(params=["direct", "chirpz"])
def fft_path(request, monkeypatch):
    """Rows whose length has no prime factor above 23 run on the direct mixed-radix kernels by default; the small test grids consist mostly of
    such rows, so the tests that take this fixture run a second time with every row on the chirp-z kernels."""
    monkeypatch.setenv("SPTRANS_FFT_DIRECT", "1" if request.param == "direct" else "0")
    return request.param


def make(gridname, T):
    import atlas_b200
    from oracle import pyoracle as po

    grid = atlas_b200.Grid(gridname)
    trans = atlas_b200.Trans(grid, T, atlas_b200.option.type("b200"))
    plan = po.OraclePlan(grid.nx(), grid.y(), T, regular=grid.regular, weights=grid.weights())
    return grid, trans, plan


def test_library_is_loaded_and_sees_the_gpu(torch_cuda):
    from atlas_b200 import _lib

    assert _lib.lib.sptrans_device_count() >= 1


@pytest.mark.parametrize("gridname,T", [("O16", 15), ("O32", 31), ("F24", 23), ("L9", 17), ("O80", 79)])
def test_legendre_tables_bit_identical_to_oracle(gridname, T):
    """Device-generated Pnm, exported in the reference's cache layout (TransLocal.cc:592-647), equal the
    oracle's tables bit for bit (which in turn are bit-identical to the unmodified reference source)."""
    grid, trans, plan = make(gridname, T)
    sym, asym, sb, ab = plan.tables()
    blob = trans.export_legendre_cache()
    assert blob.size == sym.size + asym.size
    assert np.array_equal(blob[:sym.size], sym)
    assert np.array_equal(blob[sym.size:], asym)
    assert np.array_equal(trans.nlat0(), plan.nlat0())


def test_legendre_table_vs_reference_golden():
    """Directly against the committed output of the unmodified reference (tests/golden/legendre_ref_T33.npz)."""
    import atlas_b200

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "legendre_ref_T33.npz"))
    trc = int(g["trc"])  # table truncation T+1 = 33 -> T = 32
    lats = np.rad2deg(g["lats"])
    # a 2-row-per-hemisphere "grid" whose northern latitudes are the golden ones (sorted north->south)
    order = np.argsort(-lats)
    lat = np.concatenate([lats[order], -lats[order][::-1]])
    grid = atlas_b200.StructuredGrid("golden", np.full(lat.size, 4 * 40, dtype=np.int32), lat, None, regular=True)
    trans = atlas_b200.Trans(grid, trc - 1)
    blob = trans.export_legendre_cache()
    nleg = lats.size
    pos_s = pos_a = 0
    size_sym = sum(-(-(((trc - m + 2) // 2) * nleg) // 8) * 8 for m in range(trc + 1))
    for m in range(trc + 1):
        Ks, Ka = (trc - m + 2) // 2, (trc - m + 1) // 2
        for jj, j in enumerate(order):
            ks = ka = 0
            for n in range(trc, m - 1, -1):
                want = g["legpol"][j][(2 * trc + 3 - m) * m // 2 + n - m]
                if (n - m) % 2 == 0:
                    got = blob[pos_s + Ks * jj + ks]
                    ks += 1
                else:
                    got = blob[size_sym + pos_a + Ka * jj + ka]
                    ka += 1
                assert got == want, (m, n, j)
        pos_s += -(-(Ks * nleg) // 8) * 8
        pos_a += -(-(Ka * nleg) // 8) * 8


@pytest.mark.parametrize("gridname,T,nf", [("O32", 31, 4), ("O32", 31, 1), ("F24", 23, 3), ("L9", 17, 2),
                                           ("O48", 47, 137), ("O160", 159, 10), ("O80", 79, 75)])
def test_invtrans_scalar_matches_oracle(gridname, T, nf, fft_path):
    grid, trans, plan = make(gridname, T)
    sp = H.synthetic_spectra(T, nf)
    gp = np.full(nf * grid.size(), np.nan)
    trans.invtrans(nf, sp, gp)
    want = plan.invtrans(nf, sp, mode=2)
    assert H.compute_rms(gp, want) < TOL_RMS
    assert H.rel_max(gp, want) < TOL_MAX


def test_invtrans_literal_reference_path_small():
    """BASELINE config 1 (O32 L4) against the literal single-thread restatement (naive GEMM + naive DFT)."""
    grid, trans, plan = make("O32", 31)
    sp = H.synthetic_spectra(31, 4)
    gp = np.full(4 * grid.size(), np.nan)
    trans.invtrans(4, sp, gp)
    want = plan.invtrans(4, sp, mode=0)
    assert H.compute_rms(gp, want) < TOL_RMS and H.rel_max(gp, want) < TOL_MAX


@pytest.mark.parametrize("gridname,T,regular", [("F32", 31, True), ("O32", 31, False), ("O64", 63, False), ("L9", 17, True)])
def test_invtrans_vs_closed_form_harmonics(gridname, T, regular):
    """The reference's own acceptance test (test_transgeneral.cc:493-643) run on the CUDA path."""
    from oracle import pyoracle as po

    grid, trans, plan = make(gridname, T)
    nx, lat = grid.nx(), grid.y()
    lon, latp = H.grid_lonlat(nx, np.clip(lat, -89.9999999, 89.9999999))
    cases = [(m, n, im) for m in range(min(T, 45)) for n in range(m, min(T, 45) + 1) if H.has_closed_form(n, m)
             for im in (0, 1) if not (m == 0 and im == 1)]
    nf = len(cases)
    sp = np.zeros((T + 1) * (T + 2) * nf)
    for f, (m, n, im) in enumerate(cases):
        sp[H.spec_index(T, m, n, im, nf, f)] = 1.0
    gp = np.full(nf * grid.size(), np.nan)
    trans.invtrans(nf, sp, gp)
    gp = gp.reshape(nf, -1)
    worst = 0.0
    for f, (m, n, im) in enumerate(cases):
        want = H.analytic_harmonic(n, m, im, lon, latp)
        # (the expectation is built with the ORACLE's fourier_truncation, not the product's)
        mask = H.expected_zonal_mask(T, nx, lat, regular, m, po.lib().orc_fourier_truncation)
        want = np.where(mask, want, 0.0)
        worst = max(worst, H.compute_rms(gp[f], want))
    assert worst < TOL_RMS, worst


@pytest.mark.parametrize("gridname,T,nsc,nvd", [("O32", 31, 4, 2), ("F24", 23, 0, 1), ("O48", 47, 3, 5), ("L9", 17, 1, 1)])
def test_invtrans_vordiv_matches_oracle(gridname, T, nsc, nvd, fft_path):
    """TransLocal::invtrans(nscal, sp, nvordiv, vor, div, gp) (TransLocal.cc:1523-1597): T+1 path, u/cos scaling."""
    grid, trans, plan = make(gridname, T)
    sp = H.synthetic_spectra(T, nsc) if nsc else None
    vor = H.synthetic_spectra(T, nvd, seed=11)
    div = H.synthetic_spectra(T, nvd, seed=12)
    nall = nsc + 2 * nvd
    gp = np.full(nall * grid.size(), np.nan)
    trans.invtrans(nsc, sp, nvd, vor, div, gp)
    want = plan.invtrans(nsc, sp, nvd, vor, div, mode=2)
    # grids with pole rows divide U,V by cos(89.9999999 deg) = 1.7e-9, amplifying rounding differences by ~1e9;
    # the reference relaxes its own wind tolerance to 2e-5 there (test_transgeneral.cc:1040-1045)
    pole = abs(grid.y(0)) > 89.9999999
    assert H.compute_rms(gp, want) < (1e-8 if pole else 1e-12)
    assert H.rel_max(gp, want) < (1e-7 if pole else 1e-11)


def test_invtrans_wind_vs_closed_form():
    grid, trans, plan = make("F32", 31)
    T = 31
    lon, latp = H.grid_lonlat(grid.nx(), grid.y())
    npts = grid.size()
    for (n, m) in ((1, 0), (1, 1)):
        for imag in (0, 1):
            if m == 0 and imag == 1:
                continue
            for var_in in (0, 1):
                vor = np.zeros((T + 1) * (T + 2))
                div = np.zeros((T + 1) * (T + 2))
                (vor if var_in == 0 else div)[H.spec_index(T, m, n, imag)] = 1.0
                gp = np.full(2 * npts, np.nan)
                trans.invtrans(1, vor, div, gp)
                for var_out in (0, 1):
                    want = H.analytic_wind(n, m, imag, lon, latp, var_in, var_out)
                    got = gp[var_out * npts:(var_out + 1) * npts]
                    if np.abs(want).max() == 0:
                        assert np.abs(got).max() < 1e-3
                    else:
                        assert H.compute_rms(got, want) < 2e-6  # the reference's wind tolerance


def test_vordiv_to_uv_matches_oracle():
    import atlas_b200
    from oracle import pyoracle as po

    for T, nf in ((10, 2), (63, 5), (32, 1)):
        vor = H.synthetic_spectra(T, nf, seed=3)
        div = H.synthetic_spectra(T, nf, seed=4)
        U = np.full_like(vor, np.nan)
        V = np.full_like(vor, np.nan)
        atlas_b200.VorDivToUV(T).execute(vor.size // nf, nf, vor, div, U, V)
        Uo, Vo = po.vd2uv(T, nf, vor, div)
        assert H.rel_max(U, Uo) < 1e-14 and H.rel_max(V, Vo) < 1e-14


@pytest.mark.parametrize("gridname,T,nf", [("O32", 31, 4), ("F24", 23, 3), ("O48", 47, 137), ("O160", 159, 7)])
def test_dirtrans_matches_oracle(gridname, T, nf, fft_path):
    grid, trans, plan = make(gridname, T)
    sp = H.synthetic_spectra(T, nf)
    gp = plan.invtrans(nf, sp, mode=2)
    got = np.full_like(sp, np.nan)
    trans.dirtrans(nf, gp, got)
    want = plan.dirtrans(nf, gp)
    assert H.rel_max(got, want) < TOL_MAX


def test_dirtrans_unit_harmonic_gives_unit_coefficient():
    """test_trans_levels semantics (test_transgeneral.cc:1494-1585): gp = Y(n,m) => coefficient 1, others 0."""
    grid, trans, plan = make("F24", 23)
    T = 23
    lon, latp = H.grid_lonlat(grid.nx(), grid.y())
    for (m, n, im) in ((0, 0, 0), (0, 3, 0), (2, 3, 1), (5, 5, 0), (1, 2, 1)):
        gp = np.ascontiguousarray(H.analytic_harmonic(n, m, im, lon, latp))
        sp = np.full((T + 1) * (T + 2), np.nan)
        trans.dirtrans(1, gp, sp)
        want = np.zeros_like(sp)
        want[H.spec_index(T, m, n, im)] = 1.0
        assert np.abs(sp - want).max() < 1e-13


@pytest.mark.parametrize("gridname,T,nf", [("F24", 23, 2), ("O48", 47, 3), ("O160", 159, 4)])
def test_dirtrans_wind2vordiv_matches_oracle(gridname, T, nf, fft_path):
    """TransImpl::dirtrans(nb_fields, wind, vor, div): parity unpinned in the reference (TransLocal: NotImplemented);
    checked against the oracle's definition and as the inverse of invtrans_vordiv2wind."""
    grid, trans, plan = make(gridname, T)
    vor = H.synthetic_spectra(T, nf, seed=21)
    div = H.synthetic_spectra(T, nf, seed=22)
    vor.reshape(-1, 2, nf)[0] = 0.0  # the global mean of vorticity / divergence is unphysical (lap^-1 undefined)
    div.reshape(-1, 2, nf)[0] = 0.0
    wind = np.full(2 * nf * grid.size(), np.nan)
    trans.invtrans(nf, vor, div, wind)
    v2 = np.full_like(vor, np.nan)
    d2 = np.full_like(div, np.nan)
    trans.dirtrans(nf, wind, v2, d2)
    ov, od = plan.dirtrans_wind(nf, wind)
    assert H.rel_max(v2, ov) < 1e-11 and H.rel_max(d2, od) < 1e-11
    tol = 1e-12 if grid.regular else 1e-9  # octahedral rows are pruned per wavenumber: quadrature no longer exact
    assert H.rel_max(v2, vor) < tol and H.rel_max(d2, div) < tol


@pytest.mark.parametrize("gridname,T,nf", [("F24", 23, 2), ("O48", 47, 5), ("O80", 79, 3)])
def test_invtrans_grad_matches_oracle_and_closed_form(gridname, T, nf, fft_path):
    """invtrans_grad (TransIFS semantics, ifs/TransIFS.cc:2075-2142): [E-W | N-S]; the oracle evaluates it through
    the reference's own vd2uv + inverse (grad f = irrotational wind of the velocity potential f)."""
    grid, trans, plan = make(gridname, T)
    sp = H.synthetic_spectra(T, nf, seed=31)
    g = np.full(2 * nf * grid.size(), np.nan)
    trans.invtrans_grad(nf, sp, g)
    want = plan.invtrans_grad(nf, sp)
    assert H.compute_rms(g, want) < 1e-12 and H.rel_max(g, want) < 1e-11
    # closed form: f = 2 Pbar_2^1 cos(lon)
    sp1 = np.zeros((T + 1) * (T + 2))
    sp1[H.spec_index(T, 1, 2, 0)] = 1.0
    g1 = np.full(2 * grid.size(), np.nan)
    trans.invtrans_grad(1, sp1, g1)
    lon, latp = H.grid_lonlat(grid.nx(), grid.y())
    s, c = np.sin(latp), np.cos(latp)
    a = H.EARTH_RADIUS
    ew = -np.sqrt(7.5) * s * 2 * np.sin(lon) / a
    ns = np.sqrt(7.5) * (c * c - s * s) * 2 * np.cos(lon) / a
    npts = grid.size()
    assert H.compute_rms(g1[:npts], ew) < 1e-13 and H.compute_rms(g1[npts:], ns) < 1e-13


def test_round_trip_regular_grid():
    grid, trans, plan = make("F48", 47)
    T, nf = 47, 6
    sp = H.synthetic_spectra(T, nf)
    gp = np.zeros(nf * grid.size())
    trans.invtrans(nf, sp, gp)
    back = np.zeros_like(sp)
    trans.dirtrans(nf, gp, back)
    want = sp.copy().reshape(-1, 2, nf)
    want[-1] = 0.0  # the scalar inverse drops m == T (TransLocal.cc:982)
    assert np.abs(back - want.reshape(-1)).max() < 1e-13


def test_device_pointers_in_place(torch_cuda):
    torch = torch_cuda
    grid, trans, plan = make("O48", 47)
    T, nf = 47, 9
    sp = H.synthetic_spectra(T, nf)
    d_sp = torch.from_numpy(sp).cuda()
    d_gp = torch.full((nf * grid.size(),), float("nan"), dtype=torch.float64, device="cuda")
    trans.invtrans(nf, d_sp, d_gp)
    want = plan.invtrans(nf, sp, mode=2)
    assert H.rel_max(d_gp.cpu().numpy(), want) < TOL_MAX
    d_back = torch.zeros_like(d_sp)
    trans.dirtrans(nf, d_gp, d_back)
    assert H.rel_max(d_back.cpu().numpy(), plan.dirtrans(nf, want)) < TOL_MAX
    t = trans.last_timings()
    assert t["legendre"] > 0 and t["fourier"] > 0 and t["h2d"] == 0 and t["d2h"] == 0


def test_edge_cases():
    import atlas_b200
    from atlas_b200 import _lib

    grid, trans, plan = make("O16", 15)
    # nb_fields == 0 is a no-op (reference: `if (nb_scalar_fields > 0)`, TransLocal.cc:1412)
    trans.invtrans(0, np.zeros(2), np.zeros(2))
    # truncation 0: only the global mean survives... and the scalar path drops m == T == 0 entirely
    g0, t0, p0 = make("F8", 0)
    sp = np.array([3.0, 0.0])
    gp = np.full(g0.size(), np.nan)
    t0.invtrans(1, sp, gp)
    assert np.array_equal(gp, p0.invtrans(1, sp, mode=2))
    # asymmetric / regional grids are refused like TransLocal refuses non-nested regional grids
    with pytest.raises(_lib.SptransError):
        atlas_b200.Trans(atlas_b200.StructuredGrid("bad", [8, 8, 8], [60.0, 10.0, -50.0]), 3)
    # dirtrans without quadrature weights
    gL, tL, pL = make("L9", 17)
    with pytest.raises(_lib.SptransError):
        tL.dirtrans(1, np.zeros(gL.size()), np.zeros(18 * 19))


def test_config2_tco399_l137_vs_oracle():
    """BASELINE config 2: TCo399 L137 inverse + direct against the (multi-threaded) oracle."""
    grid, trans, plan = make("O400", 399)
    T, nf = 399, 137
    sp = H.synthetic_spectra(T, nf)
    gp = np.zeros(nf * grid.size())
    trans.invtrans(nf, sp, gp)
    want = plan.invtrans(nf, sp, mode=2)
    assert H.compute_rms(gp, want) < TOL_RMS
    assert H.rel_max(gp, want) < TOL_MAX
    back = np.zeros_like(sp)
    trans.dirtrans(nf, want, back)
    want_sp = plan.dirtrans(nf, want)
    assert H.rel_max(back, want_sp) < TOL_MAX


def test_config3_tco1279_l137_properties(torch_cuda):
    """BASELINE config 3/4 size (TCo1279 L137) on one GPU through size-independent properties:
    linearity, field independence, and agreement with the oracle on a field subset (the Legendre/Fourier
    stages treat fields independently, so the oracle only needs to transform the sampled fields)."""
    torch = torch_cuda
    grid, trans, plan = make("O1280", 1279)
    T, nf = 1279, 137
    sp = torch.from_numpy(H.synthetic_spectra(T, nf)).cuda()
    gp = torch.empty(nf * grid.size(), dtype=torch.float64, device="cuda")
    trans.invtrans(nf, sp, gp)
    sample = [0, 68, 136]
    sp_s = np.ascontiguousarray(sp.view(-1, nf)[:, sample].cpu().numpy().reshape(-1))
    want = plan.invtrans(len(sample), sp_s, mode=2).reshape(len(sample), -1)
    got = gp.view(nf, -1)[sample].cpu().numpy()
    assert H.compute_rms(got, want) < TOL_RMS and H.rel_max(got, want) < TOL_MAX
    # linearity: T(2 x) == 2 T(x) bit for bit (scaling by 2 is exact in binary floating point)
    gp2 = torch.empty_like(gp)
    sp2 = sp * 2.0
    trans.invtrans(nf, sp2, gp2)
    assert torch.equal(gp2, gp * 2.0)
    # direct transform of the grid fields: oracle on the same field subset
    back = torch.empty_like(sp)
    trans.dirtrans(nf, gp, back)
    want_sp = plan.dirtrans(len(sample), want.reshape(-1)).reshape(-1, len(sample))
    got_sp = back.view(-1, nf)[:, sample].cpu().numpy()
    assert H.rel_max(got_sp, want_sp) < TOL_MAX
    t = trans.last_timings()
    print("TCo1279 L137 dirtrans stage ms:", t)


@pytest.mark.parametrize("gridname,T,nf,R", [("O48", 47, 5, 2), ("O80", 79, 4, 3), ("F24", 23, 3, 2)])
def test_sharded_path_emulated_on_one_gpu(torch_cuda, gridname, T, nf, R):
    """The multi-GPU algorithm (m-sharded Legendre -> pack -> all-to-all -> unpack -> band-sharded Fourier, and
    the mirror image for the direct transform) with all R ranks living on one device and the all-to-all replaced
    by slice copies.  Exercises the sharded plans, table pruning, tile lists and gather/scatter kernels."""
    import ctypes as C

    import atlas_b200
    from atlas_b200 import _lib
    from atlas_b200.trans import _ptr
    from oracle import pyoracle as po

    torch = torch_cuda
    grid = atlas_b200.Grid(gridname)
    plans = [atlas_b200.Trans(grid, T, rank=r, nranks=R) for r in range(R)]
    LL = C.POINTER(C.c_longlong)
    m_rows, b_rows = [], []
    for t in plans:
        ms, bs = np.zeros(R, dtype=np.int64), np.zeros(R, dtype=np.int64)
        _lib.check(_lib.lib.sptrans_exchange_rows(t._h, ms.ctypes.data_as(LL), bs.ctypes.data_as(LL)))
        m_rows.append(ms)
        b_rows.append(bs)
    k = 2 * nf
    kw = dict(dtype=torch.float64, device="cuda")
    fb = [torch.zeros(t.fourier_elems_per_field() * k, **kw) for t in plans]
    buf_m = [torch.zeros(max(int(m_rows[r].sum()), 1) * k, **kw) for r in range(R)]
    buf_b = [torch.zeros(max(int(b_rows[r].sum()), 1) * k, **kw) for r in range(R)]

    def all_to_all(src, src_rows, dst, dst_rows):
        for s in range(R):
            so = np.concatenate([[0], np.cumsum(src_rows[s])]) * k
            for d in range(R):
                do = np.concatenate([[0], np.cumsum(dst_rows[d])]) * k
                assert src_rows[s][d] == dst_rows[d][s]
                dst[d][do[s]:do[s + 1]] = src[s][so[d]:so[d + 1]]

    sp = H.synthetic_spectra(T, nf)
    d_sp = torch.from_numpy(sp).cuda()
    d_gp = torch.full((nf * grid.size(),), float("nan"), **kw)
    for r, t in enumerate(plans):
        t.invtrans_legendre(nf, T, d_sp, fb[r])
        _lib.check(_lib.lib.sptrans_exchange_pack(t._h, nf, 0, _ptr(fb[r]), _ptr(buf_m[r])))
    all_to_all(buf_m, m_rows, buf_b, b_rows)
    for r, t in enumerate(plans):
        _lib.check(_lib.lib.sptrans_exchange_unpack(t._h, nf, 1, _ptr(buf_b[r]), _ptr(fb[r])))
        t.invtrans_fourier(nf, T - 1, fb[r], d_gp)
    plan = po.OraclePlan(grid.nx(), grid.y(), T, regular=grid.regular, weights=grid.weights())
    want = plan.invtrans(nf, sp, mode=2)
    assert H.rel_max(d_gp.cpu().numpy(), want) < TOL_MAX
    # direct transform, mirror image
    d_back = torch.full_like(d_sp, float("nan"))
    for r, t in enumerate(plans):
        fb[r].zero_()
        t.dirtrans_fourier(nf, d_gp, fb[r])
        _lib.check(_lib.lib.sptrans_exchange_pack(t._h, nf, 1, _ptr(fb[r]), _ptr(buf_b[r])))
    all_to_all(buf_b, b_rows, buf_m, m_rows)
    for r, t in enumerate(plans):
        _lib.check(_lib.lib.sptrans_exchange_unpack(t._h, nf, 0, _ptr(buf_m[r]), _ptr(fb[r])))
        t.dirtrans_legendre(nf, fb[r], d_back)  # every rank writes the coefficients of its own zonal wavenumbers
    assert H.rel_max(d_back.cpu().numpy(), plan.dirtrans(nf, want)) < TOL_MAX


@pytest.mark.parametrize("gridname,T,nf,R", [("O48", 47, 5, 2), ("O80", 79, 4, 3), ("F24", 23, 3, 1), ("O48", 47, 137, 4)])
def test_peer_memory_exchange_emulated_on_one_gpu(torch_cuda, gridname, T, nf, R, fft_path):
    """The peer-memory exchange with all R ranks living on one device: every plan's exchange region is handed to the
    others as plain pointers (sptrans_peer_attach_ptrs), the inverse Legendre kernel stores each row into the buffer of
    the rank that owns its latitude band, the direct transform pushes rows to the owner of their zonal wavenumber.
    The stages of the R ranks are sequenced on one stream here (the cross-GPU barrier is covered by test_gpu_dist)."""
    import ctypes as C

    import atlas_b200
    from atlas_b200 import _lib
    from atlas_b200.trans import _ptr
    from oracle import pyoracle as po

    torch = torch_cuda
    lib = _lib.lib
    grid = atlas_b200.Grid(gridname)
    plans = [atlas_b200.Trans(grid, T, rank=r, nranks=R) for r in range(R)]
    stream = torch.cuda.current_stream().cuda_stream
    regions = (C.c_void_p * R)()
    for r, t in enumerate(plans):
        t.set_stream(stream)
        _lib.check(lib.sptrans_peer_alloc(t._h, nf, None))
        reg = C.c_void_p()
        _lib.check(lib.sptrans_peer_region(t._h, C.byref(reg), None))
        regions[r] = reg
    for t in plans:
        _lib.check(lib.sptrans_peer_attach_ptrs(t._h, R, regions))
    sp = H.synthetic_spectra(T, nf)
    d_sp = torch.from_numpy(sp).cuda()
    kw = dict(dtype=torch.float64, device="cuda")
    plan = po.OraclePlan(grid.nx(), grid.y(), T, regular=grid.regular, weights=grid.weights())
    want = plan.invtrans(nf, sp, mode=2)
    want_sp = plan.dirtrans(nf, want)

    def local_buffer(t):
        b = C.c_void_p()
        _lib.check(lib.sptrans_peer_buffer(t._h, C.byref(b)))
        return b

    for it in range(3):  # both exchange buffers are used, and reused
        d_gp = torch.full((nf * grid.size(),), float("nan"), **kw)
        for t in plans:
            _lib.check(lib.sptrans_invtrans_legendre_peers(t._h, nf, _ptr(d_sp)))
        for t in plans:
            _lib.check(lib.sptrans_invtrans_fourier(t._h, nf, T - 1, local_buffer(t), _ptr(d_gp), 0))
            _lib.check(lib.sptrans_peer_advance(t._h))
        assert H.rel_max(d_gp.cpu().numpy(), want) < TOL_MAX
        d_back = torch.full_like(d_sp, float("nan"))
        for t in plans:
            _lib.check(lib.sptrans_dirtrans_fourier_peers(t._h, nf, _ptr(d_gp)))
        for t in plans:
            _lib.check(lib.sptrans_dirtrans_legendre(t._h, nf, local_buffer(t), _ptr(d_back)))
            _lib.check(lib.sptrans_peer_advance(t._h))
        assert H.rel_max(d_back.cpu().numpy(), want_sp) < TOL_MAX
        # the same exchange by PULL (what sptrans_dirtrans_sharded does): Fourier rows stay in the owner's buffer and the
        # Legendre GEMM's TMA copies fetch them from there
        d_back = torch.full_like(d_sp, float("nan"))
        for t in plans:
            _lib.check(lib.sptrans_dirtrans_fourier_local(t._h, nf, _ptr(d_gp)))
        for t in plans:
            _lib.check(lib.sptrans_dirtrans_legendre_pull(t._h, nf, _ptr(d_back)))
        for t in plans:
            _lib.check(lib.sptrans_peer_advance(t._h))
        assert H.rel_max(d_back.cpu().numpy(), want_sp) < TOL_MAX
    if R == 1:  # the whole stream-ordered call, barrier kernel included (trivial with one rank)
        d_gp = torch.full((nf * grid.size(),), float("nan"), **kw)
        d_back = torch.full_like(d_sp, float("nan"))
        _lib.check(lib.sptrans_invtrans_sharded(plans[0]._h, nf, _ptr(d_sp), _ptr(d_gp)))
        _lib.check(lib.sptrans_dirtrans_sharded(plans[0]._h, nf, _ptr(d_gp), _ptr(d_back)))
        torch.cuda.synchronize()
        assert H.rel_max(d_gp.cpu().numpy(), want) < TOL_MAX
        assert H.rel_max(d_back.cpu().numpy(), want_sp) < TOL_MAX
        tm = plans[0].last_timings()
        assert tm["legendre"] > 0 and tm["fourier"] > 0
    for t in plans:
        _lib.check(lib.sptrans_peer_free(t._h))


@pytest.mark.parametrize("gridname,T,nf,R", [("O48", 47, 5, 2), ("O80", 79, 4, 3), ("L9", 17, 2, 2), ("O160", 159, 9, 8)])
def test_shard_local_io_emulated_on_one_gpu(torch_cuda, gridname, T, nf, R, fft_path):
    """SPTRANS_SHARD_LOCAL_IO: every rank's spectral array holds only its zonal wavenumbers ([m ascending][n][re/im][field])
    and its grid array only the rows of its latitude band ([field][northern rows, then their southern mirrors]) -- what
    bench.py's N > 1 runs and BASELINE config 5 use.  R ranks emulated on one device over the peer-memory exchange;
    the assembled global results against the CPU oracle."""
    import ctypes as C

    import atlas_b200
    from atlas_b200 import _lib
    from atlas_b200.dist import shard_layout
    from atlas_b200.trans import _ptr
    from oracle import pyoracle as po

    torch = torch_cuda
    lib = _lib.lib
    grid = atlas_b200.Grid(gridname)
    npts, nlat = grid.size(), grid.ny()
    plans = [atlas_b200.Trans(grid, T, rank=r, nranks=R, local_io=True) for r in range(R)]
    stream = torch.cuda.current_stream().cuda_stream
    regions = (C.c_void_p * R)()
    for r, t in enumerate(plans):
        t.set_stream(stream)
        _lib.check(lib.sptrans_peer_alloc(t._h, nf, None))
        reg = C.c_void_p()
        _lib.check(lib.sptrans_peer_region(t._h, C.byref(reg), None))
        regions[r] = reg
    for t in plans:
        _lib.check(lib.sptrans_peer_attach_ptrs(t._h, R, regions))
    owner, band, _, _ = shard_layout(grid, T, 0, R)
    rowoff = np.concatenate([[0], np.cumsum(grid.nx(), dtype=np.int64)])

    def spans(r):
        j0, j1 = int(band[r]), int(band[r + 1])
        sp_ = [(int(rowoff[j0]), int(rowoff[j1])), (int(rowoff[nlat - j1]), int(rowoff[nlat - j0]))]
        if sp_[0][1] > sp_[1][0]:
            sp_ = [(sp_[0][0], sp_[1][1])]
        return [(a, b) for a, b in sp_ if b > a]

    def m_slices(r):
        return [((2 * T + 3 - m) * m // 2 * 2 * nf, (T - m + 1) * 2 * nf) for m in range(T + 1) if owner[m] == r]

    sp = H.synthetic_spectra(T, nf)
    plan = po.OraclePlan(grid.nx(), grid.y(), T, regular=grid.regular, weights=grid.weights())
    want = plan.invtrans(nf, sp, mode=2)
    kw = dict(dtype=torch.float64, device="cuda")
    d_sp, d_gp, sizes = [], [], []
    for r, t in enumerate(plans):
        nsp, stride = t.local_sizes()
        parts = [sp[o:o + n] for o, n in m_slices(r)]
        loc = np.concatenate(parts) if parts else np.zeros(0)
        assert loc.size == nsp * nf and stride >= sum(b - a for a, b in spans(r))
        d_sp.append(torch.from_numpy(np.ascontiguousarray(loc)).cuda() if loc.size else torch.zeros(2, **kw))
        d_gp.append(torch.full((max(nf * stride, 2),), float("nan"), **kw))
        sizes.append((nsp, stride))

    def local_buffer(t):
        b = C.c_void_p()
        _lib.check(lib.sptrans_peer_buffer(t._h, C.byref(b)))
        return b

    for t, a in zip(plans, d_sp):
        _lib.check(lib.sptrans_invtrans_legendre_peers(t._h, nf, _ptr(a)))
    for t, g in zip(plans, d_gp):
        _lib.check(lib.sptrans_invtrans_fourier(t._h, nf, T - 1, local_buffer(t), _ptr(g), 0))
        _lib.check(lib.sptrans_peer_advance(t._h))
    got = np.full((nf, npts), np.nan)
    for r in range(R):
        loc = d_gp[r].cpu().numpy()[: nf * sizes[r][1]].reshape(nf, sizes[r][1])
        o = 0
        for a, b in spans(r):
            got[:, a:b] = loc[:, o:o + b - a]
            o += b - a
    assert H.rel_max(got.reshape(-1), want) < TOL_MAX
    if grid.weights() is None:
        return
    d_back = [torch.full_like(a, float("nan")) for a in d_sp]
    for t, g in zip(plans, d_gp):
        _lib.check(lib.sptrans_dirtrans_fourier_local(t._h, nf, _ptr(g)))
    for t, b in zip(plans, d_back):
        _lib.check(lib.sptrans_dirtrans_legendre_pull(t._h, nf, _ptr(b)))
    for t in plans:
        _lib.check(lib.sptrans_peer_advance(t._h))
    back = np.full_like(sp, np.nan)
    for r in range(R):
        loc = d_back[r].cpu().numpy()
        o = 0
        for g0, n in m_slices(r):
            back[g0:g0 + n] = loc[o:o + n]
            o += n
    assert H.rel_max(back, plan.dirtrans(nf, want)) < TOL_MAX
    for t in plans:
        _lib.check(lib.sptrans_peer_free(t._h))


@pytest.mark.parametrize("gridname,T,nf", [("O32", 31, 4), ("O48", 47, 137), ("O160", 159, 20), ("F24", 23, 3)])
def test_tensor_core_split_tf32_legendre(gridname, T, nf):
    """BASELINE config 4: Legendre stage on tcgen05 (kind::tf32, operands split hi+lo, fp32 accumulation in TMEM).
    Stated tolerance vs the fp64 oracle: 2e-6 relative (max norm) -- fp32-level, three orders better than plain TF32."""
    grid, trans, plan = make(gridname, T)
    trans.set_precision("tc")
    sp = H.synthetic_spectra(T, nf)
    gp = np.full(nf * grid.size(), np.nan)
    trans.invtrans(nf, sp, gp)
    want = plan.invtrans(nf, sp, mode=2)
    err = H.rel_max(gp, want)
    assert err < 2e-6, err
    assert err > 1e-12  # it really is the reduced-precision path
    back = np.full_like(sp, np.nan)
    trans.dirtrans(nf, want, back)
    want_sp = plan.dirtrans(nf, want)
    assert H.rel_max(back, want_sp) < 2e-6
    trans.set_precision("fp64")
    trans.invtrans(nf, sp, gp)
    assert H.rel_max(gp, want) < TOL_MAX


def test_row_mode_fourier_kernels_small_grid():
    """Rows too long for the packed north/south transform (O2560: n + 2L > 13824) go through the row-mode kernels
    (even/odd packing, complex length n/2).  SPTRANS_FFT_MAXM lowers the limit so a small grid exercises them;
    run in a subprocess because the limit is read once per process."""
    import subprocess
    import sys

    code = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import helpers as H
import atlas_b200
from oracle import pyoracle as po
gridname, T, nf = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
for _ in range(1):
    grid = atlas_b200.Grid(gridname)
    trans = atlas_b200.Trans(grid, T)
    plan = po.OraclePlan(grid.nx(), grid.y(), T, regular=grid.regular, weights=grid.weights())
    sp = H.synthetic_spectra(T, nf)
    gp = np.full(nf * grid.size(), np.nan)
    trans.invtrans(nf, sp, gp)
    want = plan.invtrans(nf, sp, mode=2)
    e1 = H.rel_max(gp, want)
    e2 = 0.0
    if grid.weights() is not None:
        back = np.full_like(sp, np.nan)
        trans.dirtrans(nf, want, back)
        e2 = H.rel_max(back, plan.dirtrans(nf, want))
    print("ROWMODE", gridname, e1, e2)
    assert e1 < 1e-12 and e2 < 1e-12, (gridname, e1, e2)
print("ROWMODE_OK")
'''
    # (grid, T, fields, limit): the limit is chosen so that the longest rows need row mode (n + 2L > limit >= n/2 + 2L)
    for gridname, T, nf, maxm in (("O48", 47, 5, 256), ("L9", 17, 2, 64), ("F24", 23, 3, 96)):
        env = dict(os.environ, SPTRANS_FFT_MAXM=str(maxm))
        r = subprocess.run([sys.executable, "-c", code, gridname, str(T), str(nf)], cwd=os.path.dirname(os.path.dirname(__file__)),
                           env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert r.returncode == 0 and "ROWMODE_OK" in r.stdout, r.stdout[-3000:]


def test_shard_local_io_with_row_mode_kernels():
    """What BASELINE config 5 runs on 8 GPUs -- sharded plans, shard-local I/O arrays AND the row-mode Fourier kernels for
    the rows beyond the single-CTA limit -- at a small size: the emulated-rank test above in a subprocess whose limit is
    lowered (SPTRANS_FFT_MAXM is read once per process)."""
    import subprocess
    import sys

    env = dict(os.environ, SPTRANS_FFT_MAXM="256")
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", os.path.abspath(__file__), "-k",
           "test_shard_local_io_emulated_on_one_gpu and (O48 or L9)"]   # O48: rows beyond 256 take row mode
    r = subprocess.run(cmd, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=900)
    # (two grids x {every row on the chirp-z / row-mode kernels, smooth rows on the direct kernels next to them})
    assert r.returncode == 0 and "4 passed" in r.stdout, r.stdout[-3000:]


def test_config5_tco2559_inverse_sample():
    """BASELINE config 5 (TCo2559): the plan builds on one GPU (55 GB of pruned Legendre tables) and the inverse of a
    few fields agrees with the oracle on sampled latitude rows computed independently (direct summation of the
    Fourier series with the oracle's Legendre stage would need 67 GB of host tables, so the check is the
    zonal-mean identity and linearity instead)."""
    import torch

    import atlas_b200

    grid = atlas_b200.Grid("O2560")
    T, nf = 2559, 2
    trans = atlas_b200.Trans(grid, T)
    # zonal-mean identity: only the (m=0, n=0) coefficient => constant field
    sp = np.zeros((T + 1) * (T + 2) * nf)
    sp[H.spec_index(T, 0, 0, 0, nf, 0)] = 4.0
    sp[H.spec_index(T, 0, 0, 0, nf, 1)] = -1.5
    d_sp = torch.from_numpy(sp).cuda()
    d_gp = torch.full((nf * grid.size(),), float("nan"), dtype=torch.float64, device="cuda")
    trans.invtrans(nf, d_sp, d_gp)
    g = d_gp.view(nf, -1)
    assert float((g[0] - 4.0).abs().max()) < 1e-12 and float((g[1] + 1.5).abs().max()) < 1e-12
    # a sectoral harmonic at high wavenumber against its closed form on every row where it is resolved
    m = 1000
    sp[:] = 0.0
    sp[H.spec_index(T, m, m, 0, nf, 0)] = 1.0
    d_sp.copy_(torch.from_numpy(sp))
    trans.invtrans(nf, d_sp, d_gp)
    lon, latp = H.grid_lonlat(grid.nx(), grid.y())
    want = H.analytic_harmonic(m, m, 0, lon, latp)
    nlat0 = trans.nlat0()
    N = 2560
    rowoff = grid.rowoff()
    keep = np.zeros(grid.ny(), dtype=bool)
    keep[nlat0[m]:2 * N - nlat0[m]] = True
    want = np.where(np.repeat(keep, grid.nx()), want, 0.0)
    got = d_gp.view(nf, -1)[0].cpu().numpy()
    assert H.compute_rms(got, want) < 1e-12
    # round trip through the direct transform
    back = torch.empty_like(d_sp)
    trans.dirtrans(nf, d_gp, back)
    b = back.cpu().numpy()
    assert abs(b[H.spec_index(T, m, m, 0, nf, 0)] - 1.0) < 1e-10
    b[H.spec_index(T, m, m, 0, nf, 0)] = 0.0
    assert np.abs(b).max() < 1e-10


@pytest.mark.parametrize("gridname,T,nf", [("F24", 23, 2), ("O48", 47, 3), ("L9", 17, 1)])
def test_invtrans_adjoint_identity(gridname, T, nf, fft_path):
    """<invtrans x, y>_grid == <x, invtrans_adj y>_spec with the spectral inner product that counts m > 0 twice (the
    reference's adjoint test, test_transgeneral.cc:1591-1722, runs this identity through TransIFS; TransLocal has no
    adjoint)."""
    grid, trans, plan = make(gridname, T)
    rng = np.random.default_rng(5)
    x = H.synthetic_spectra(T, nf, seed=41)
    y = rng.standard_normal(nf * grid.size())
    ix = np.full(nf * grid.size(), np.nan)
    trans.invtrans(nf, x, ix)
    ay = np.full_like(x, np.nan)
    trans.invtrans_adj(nf, y, ay)
    lhs, rhs = float(ix @ y), H.spectral_dot(T, nf, x, ay)
    assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), abs(rhs), np.linalg.norm(ix) * np.linalg.norm(y) * 1e-3)
    # the adjoint annihilates what the inverse ignores: Im(m = 0) and the m == T column
    ay3 = ay.reshape(-1, 2, nf)
    assert np.all(ay3[: T + 1, 1, :] == 0.0) and np.all(ay3[-1] == 0.0)


@pytest.mark.parametrize("gridname,T,nf,nb_uv", [("O1280", 1279, 7, 2), ("O400", 399, 13, 0), ("O640", 639, 3, 0)])
def test_fourier_v2_kernels_match_v1_kernels(torch_cuda, monkeypatch, gridname, T, nf, nb_uv):
    """The register-tiled (v2) Fourier kernels against the shared-memory-pass (v1) kernels on the same random exchange
    buffer / grid fields, whole grid (every block-level radix M1 occurs at O1280), both directions.  Both are chirp-z
    evaluations of the same sums, so they agree to rounding: <= 1e-12 of the field maximum."""
    import atlas_b200

    torch = torch_cuda
    grid = atlas_b200.Grid(gridname)
    monkeypatch.setenv("SPTRANS_FFT_DIRECT", "0")  # every row on the chirp-z kernels, in both plans
    monkeypatch.setenv("SPTRANS_FFT_V2", "0")
    t1 = atlas_b200.Trans(grid, T, atlas_b200.option.type("b200"))
    monkeypatch.setenv("SPTRANS_FFT_V2", "1")
    monkeypatch.setenv("SPTRANS_FFT2_F", "3")
    t2 = atlas_b200.Trans(grid, T, atlas_b200.option.type("b200"))
    kw = dict(dtype=torch.float64, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(1234)
    nfb = t1.fourier_elems_per_field() * 2 * nf
    fb = torch.rand(nfb, generator=g, **kw) - 0.5
    gp1 = torch.full((nf * grid.size(),), 7.0, **kw)
    gp2 = torch.full((nf * grid.size(),), -7.0, **kw)
    torch.cuda.synchronize()
    t1.invtrans_fourier(nf, T - 1, fb, gp1, nb_uv)
    t2.invtrans_fourier(nf, T - 1, fb, gp2, nb_uv)
    torch.cuda.synchronize()
    scale = gp1.abs().max().item()
    err = (gp1 - gp2).abs().max().item()
    assert scale > 1.0 and err <= 1e-12 * scale, (err, scale)
    gp = torch.rand(nf * grid.size(), generator=g, **kw) - 0.5
    fb1 = torch.zeros(nfb, **kw)
    fb2 = torch.zeros(nfb, **kw)
    torch.cuda.synchronize()
    t1.dirtrans_fourier(nf, gp, fb1, nb_uv)
    t2.dirtrans_fourier(nf, gp, fb2, nb_uv)
    torch.cuda.synchronize()
    scale = fb1.abs().max().item()
    err = (fb1 - fb2).abs().max().item()
    assert scale > 0 and err <= 1e-12 * scale, (err, scale)


@pytest.mark.parametrize("gridname,T,nf,nb_uv", [("O1280", 1279, 7, 2), ("O400", 399, 137, 0), ("O48", 47, 21, 3),
                                                 ("F24", 23, 5, 0), ("odd", 7, 19, 1), ("F28", 27, 33, 0)])
def test_fourier_direct_kernels_match_chirpz_kernels(torch_cuda, monkeypatch, gridname, T, nf, nb_uv):
    """The direct mixed-radix kernels (rows whose length has no prime factor above 23, no chirp-z) against the
    chirp-z kernels on the same random exchange buffer / grid fields, whole grid, both directions and the adjoint
    weighting.  Both evaluate the same trigonometric sums: <= 1e-12 of the field maximum.  The plan with the direct kernels
    must actually hold direct rows (launch count differs), else the comparison would be vacuous."""
    import atlas_b200

    torch = torch_cuda
    if gridname == "odd":  # a reduced Gaussian grid with odd and 2 (mod 4) row lengths, every prime radix among them
        from atlas_b200.grid import StructuredGrid, gaussian_latitudes
        lat, w = gaussian_latitudes(16)
        half = [17, 18, 19, 21, 23, 26, 33, 34, 35, 38, 39, 46, 49, 55, 57, 69]
        grid = StructuredGrid("odd", np.array(half + half[::-1], dtype=np.int32), lat, w, regular=False)
    else:
        grid = atlas_b200.Grid(gridname)
    nsmooth = sum(1 for n in grid.nx() if all_small_factors(int(n)))
    assert nsmooth > 0
    monkeypatch.setenv("SPTRANS_FFT_DIRECT", "0")
    t1 = atlas_b200.Trans(grid, T, atlas_b200.option.type("b200"))
    monkeypatch.setenv("SPTRANS_FFT_DIRECT", "1")
    t2 = atlas_b200.Trans(grid, T, atlas_b200.option.type("b200"))
    kw = dict(dtype=torch.float64, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(4321)
    nfb = t1.fourier_elems_per_field() * 2 * nf
    fb = torch.rand(nfb, generator=g, **kw) - 0.5
    for mlimit in (T - 1, T // 3):
        gp1 = torch.full((nf * grid.size(),), 7.0, **kw)
        gp2 = torch.full((nf * grid.size(),), -7.0, **kw)
        torch.cuda.synchronize()
        t1.invtrans_fourier(nf, mlimit, fb, gp1, nb_uv)
        t2.invtrans_fourier(nf, mlimit, fb, gp2, nb_uv)
        torch.cuda.synchronize()
        scale = gp1.abs().max().item()
        err = (gp1 - gp2).abs().max().item()
        assert scale > 1.0 and err <= 1e-12 * scale, (mlimit, err, scale)
    gp = torch.rand(nf * grid.size(), generator=g, **kw) - 0.5
    fb1 = torch.zeros(nfb, **kw)
    fb2 = torch.zeros(nfb, **kw)
    torch.cuda.synchronize()
    t1.dirtrans_fourier(nf, gp, fb1, nb_uv)
    t2.dirtrans_fourier(nf, gp, fb2, nb_uv)
    torch.cuda.synchronize()
    scale = fb1.abs().max().item()
    err = (fb1 - fb2).abs().max().item()
    assert scale > 0 and err <= 1e-12 * scale, (err, scale)
    # sptrans_fourier_path_stats: every row is accounted for, the smooth ones on the direct kernels of the second plan only
    p1, p2 = t1.fourier_paths(), t2.fourier_paths()
    assert sum(v["grid_points"] for v in p1.values()) == grid.size() == sum(v["grid_points"] for v in p2.values())
    assert p1["direct_mixed_radix"]["grid_points"] == 0
    assert p2["direct_mixed_radix"]["grid_points"] == sum(int(n) for n in grid.nx() if all_small_factors(int(n)))
    assert abs(sum(v["byte_share"] for v in p2.values()) - 1.0) < 1e-12


def all_small_factors(n):
    for q in (2, 3, 5, 7, 11, 13, 17, 19, 23):
        while n % q == 0:
            n //= q
    return n == 1


@pytest.mark.parametrize("gridname,T,nf", [("O48", 47, 5), ("O160", 159, 23), ("O400", 399, 137)])
def test_async_cloned_plans_pipelined_host_buffers(torch_cuda, gridname, T, nf):
    """sptrans_set_async + sptrans_plan_clone + the field-chunked host pipelines: an inverse on one plan and a direct
    transform on its clone (which borrows the tables) in flight at the same time, both on pinned host buffers, give
    bit-identical results to the blocking calls on device buffers (which the other tests pin against the oracle)."""
    torch = torch_cuda
    grid, trans, plan = make(gridname, T)
    npts = grid.size()
    sp = H.synthetic_spectra(T, nf)
    d_sp = torch.from_numpy(sp).cuda()
    d_gp = torch.empty(nf * npts, dtype=torch.float64, device="cuda")
    d_back = torch.empty_like(d_sp)
    trans.invtrans(nf, d_sp, d_gp)
    trans.dirtrans(nf, d_gp, d_back)
    want_gp, want_back = d_gp.cpu().numpy(), d_back.cpu().numpy()
    # blocking calls on host buffers go through the chunked pipelines too
    gp_blk = np.full(nf * npts, np.nan)
    back_blk = np.full_like(sp, np.nan)
    trans.invtrans(nf, sp, gp_blk)
    trans.dirtrans(nf, gp_blk, back_blk)
    assert np.array_equal(gp_blk, want_gp) and np.array_equal(back_blk, want_back)
    other = trans.clone()
    h_sp = torch.from_numpy(sp).pin_memory()
    h_gp_in = torch.from_numpy(want_gp).pin_memory()
    h_gp = torch.full((nf * npts,), float("nan"), dtype=torch.float64).pin_memory()
    h_back = torch.full((sp.size,), float("nan"), dtype=torch.float64).pin_memory()
    trans.set_async(True)
    other.set_async(True)
    for _ in range(3):   # back-to-back asynchronous calls reuse the plans' staging buffers in stream order
        trans.invtrans(nf, h_sp.numpy(), h_gp.numpy())
        other.dirtrans(nf, h_gp_in.numpy(), h_back.numpy())
    trans.synchronize()
    other.synchronize()
    assert np.array_equal(h_gp.numpy(), want_gp)
    assert np.array_equal(h_back.numpy(), want_back)
    # a chain through host memory ordered by device-side marks only: inverse -> host grid fields -> direct transform
    h_gp.fill_(float("nan"))
    h_back.fill_(float("nan"))
    for _ in range(2):
        trans.invtrans(nf, h_sp.numpy(), h_gp.numpy())
        m = trans.mark()
        other.wait_mark(m)
        other.dirtrans(nf, h_gp.numpy(), h_back.numpy())
        m2 = other.mark()
        trans.wait_mark(m2)       # the next inverse overwrites the host grid fields the direct transform is reading
        trans.release_mark(m)
        trans.release_mark(m2)
    other.synchronize()
    trans.synchronize()
    assert np.array_equal(h_gp.numpy(), want_gp)
    assert np.array_equal(h_back.numpy(), want_back)
    trans.set_async(False)
    del other


@pytest.mark.parametrize("gridname,T,nf,R", [("O48", 47, 5, 2), ("O160", 159, 9, 4), ("F24", 23, 3, 3)])
def test_single_process_multi_gpu_plan_emulated(torch_cuda, gridname, T, nf, R):
    """sptrans_multi_*: one host thread drives R sharded plans whose exchange regions see each other, global arrays in the
    reference's layouts in, every device given only its share (SPTRANS_SHARD_LOCAL_IO), the device-side barrier the only
    synchronisation.  Here the R "devices" are all device 0 (the real thing: tests/test_gpu_dist.py on >= 2 GPUs), host and
    device arrays, against the CPU oracle."""
    import atlas_b200

    torch = torch_cuda
    grid, trans, plan = make(gridname, T)
    del trans
    mt = atlas_b200.MultiTrans(grid, T, [0] * R)
    assert mt.size() == R
    sp = H.synthetic_spectra(T, nf)
    want = plan.invtrans(nf, sp, mode=2)
    want_sp = plan.dirtrans(nf, want)
    for it in range(2):
        gp = np.full(nf * grid.size(), np.nan)
        mt.invtrans(nf, sp, gp)
        assert H.rel_max(gp, want) < TOL_MAX
        back = np.full_like(sp, np.nan)
        mt.dirtrans(nf, gp, back)
        assert H.rel_max(back, want_sp) < TOL_MAX
    d_sp = torch.from_numpy(sp).cuda()
    d_gp = torch.full((nf * grid.size(),), float("nan"), dtype=torch.float64, device="cuda")
    mt.invtrans(nf, d_sp, d_gp)
    assert H.rel_max(d_gp.cpu().numpy(), want) < TOL_MAX
    assert mt.kernel_launches() > 0
