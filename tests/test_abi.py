"""No-GPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/sptrans_b200.h declares, and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(REPO, "include", "sptrans_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sptrans_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from atlas_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 20
    raw = C.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    unbound = [n for n in names if n not in _lib.SIGNATURES]
    assert not unbound, f"declared in the header but not bound in atlas_b200/_lib.py: {unbound}"


def test_host_side_grid_helpers_need_no_gpu():
    import atlas_b200

    g = atlas_b200.Grid("O32")
    assert g.ny() == 64 and g.size() == 5248 and g.nxmax() == 144  # config 1 sizes
    assert abs(g.weights().sum() - 1.0) < 1e-14
    gold = np.load(os.path.join(REPO, "tests", "golden", "gaussian_latitudes_N32.npy"))
    assert np.abs(g.y()[:32] - gold).max() < 5e-12
    g = atlas_b200.Grid("O1280")
    assert g.size() == 6599680 and g.nxmax() == 5136
    gold = np.load(os.path.join(REPO, "tests", "golden", "gaussian_latitudes_N1280.npy"))
    assert np.abs(g.y()[:1280] - gold).max() < 5e-12
    f = atlas_b200.Grid("F64")
    assert f.ny() == 128 and f.nx(0) == 256 and f.regular


def test_product_latitudes_agree_with_oracle():
    import atlas_b200
    from oracle import pyoracle as po

    for N in (32, 400, 1280):
        lat, w = atlas_b200.grid.gaussian_latitudes(N)
        olat, ow = po.gaussian_quadrature(N)
        assert np.abs(lat - olat).max() < 1e-12
        assert np.abs(w - ow).max() < 1e-15 * 50


def test_fourier_truncation_matches_oracle():
    from atlas_b200 import _lib
    from oracle import pyoracle as po

    rng = np.random.default_rng(1)
    for _ in range(2000):
        ndgl = int(rng.integers(4, 3000))
        T = int(rng.integers(1, 2 * ndgl))
        nx = int(rng.integers(4, 6000))
        lat = float(rng.uniform(-1.57, 1.57))
        full = int(rng.integers(0, 2))
        assert _lib.lib.sptrans_fourier_truncation(T, nx, 6000, ndgl, lat, full) == \
            po.lib().orc_fourier_truncation(T, nx, 6000, ndgl, lat, full)


def test_no_cpu_fallback():
    """Without a CUDA device the product must refuse to run (it must never route through the oracle)."""
    import atlas_b200
    from atlas_b200 import _lib

    if _lib.lib.sptrans_device_count() > 0:
        pytest.skip("a CUDA device is visible; the refusal path is covered on the CPU-only box")
    with pytest.raises(_lib.SptransError) as e:
        atlas_b200.Trans("O32", 31)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    z = np.zeros(12)
    with pytest.raises(_lib.SptransError):
        atlas_b200.VorDivToUV(1).execute(6, 1, z, z, z.copy(), z.copy())


def test_point_set_plan_refuses_loudly_and_validates_arguments():
    """sptrans_plan_create_points: bad arguments are SPTRANS_ERR_INVALID; without a device it refuses like every plan."""
    import atlas_b200
    from atlas_b200 import _lib

    h = C.c_void_p()
    lon = np.array([0.0, 10.0])
    lat = np.array([45.0, 95.0])  # latitude outside [-90, 90]
    dp = _lib.c_double_p
    rc = _lib.lib.sptrans_plan_create_points(C.byref(h), 2, lon.ctypes.data_as(dp), lat.ctypes.data_as(dp), 15, 0)
    assert rc == 1 and b"latitude" in _lib.lib.sptrans_last_error()
    rc = _lib.lib.sptrans_plan_create_points(C.byref(h), 0, lon.ctypes.data_as(dp), lat.ctypes.data_as(dp), 15, 0)
    assert rc == 1
    if _lib.lib.sptrans_device_count() == 0:
        with pytest.raises(_lib.SptransError) as e:
            atlas_b200.Trans(atlas_b200.UnstructuredGrid([0.0, 90.0], [10.0, -10.0]), 15)
        assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    with pytest.raises(ValueError):
        atlas_b200.UnstructuredGrid([0.0, 1.0], [0.0])


def test_product_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(REPO, "atlas_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".hpp", ".cuh", ".h")) or f == "Makefile":
                txt = open(os.path.join(root, f), errors="replace").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "sht_oracle" not in txt, f
