"""TEST INFRASTRUCTURE ONLY: ctypes wrapper of the CPU oracle (oracle/liboracle.so) and of the
reference's own Legendre code compiled in place (oracle/_ref/libref_legendre.so).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  Nothing under atlas_b200/ imports this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_sp = C.POINTER(C.c_size_t)


def _p(a):
    return a.ctypes.data_as(_dp)


def load_oracle():
    path = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(path):
        raise ImportError(f"{path} missing: run `make -C oracle`")
    lib = C.CDLL(path)
    lib.orc_plan_create.restype = C.c_void_p
    lib.orc_plan_create.argtypes = [C.c_int, _ip, _dp, C.c_int, C.c_int, _dp, C.c_int]
    lib.orc_plan_destroy.argtypes = [C.c_void_p]
    lib.orc_plan_nlat0.argtypes = [C.c_void_p, _ip]
    lib.orc_plan_table_sizes.restype = C.c_size_t
    lib.orc_plan_table_sizes.argtypes = [C.c_void_p, _sp, _sp]
    lib.orc_plan_leg_sym.restype = _dp
    lib.orc_plan_leg_sym.argtypes = [C.c_void_p]
    lib.orc_plan_leg_asym.restype = _dp
    lib.orc_plan_leg_asym.argtypes = [C.c_void_p]
    lib.orc_plan_npts.restype = C.c_size_t
    lib.orc_plan_npts.argtypes = [C.c_void_p]
    lib.orc_invtrans.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, _dp, _dp, _dp, C.c_int]
    lib.orc_invtrans_legendre.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, C.c_int]
    lib.orc_vd2uv.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp]
    lib.orc_extend_truncation.argtypes = [C.c_int, C.c_int, _dp, _dp]
    lib.orc_dirtrans.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    lib.orc_dirtrans_wind.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
    lib.orc_invtrans_grad.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    lib.orc_gaussian_quadrature.argtypes = [C.c_int, _dp, _dp]
    lib.orc_compute_zfn.argtypes = [C.c_int, _dp]
    lib.orc_legendre_lat.argtypes = [C.c_int, C.c_double, _dp, _dp]
    lib.orc_fourier_truncation.restype = C.c_int
    lib.orc_fourier_truncation.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
    lib.orc_invtrans_unstructured.argtypes = [C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp, _dp, _dp, C.c_int]
    lib.orc_c2r.argtypes = [C.c_int, _dp, _dp, C.c_int]
    lib.orc_r2c.argtypes = [C.c_int, _dp, _dp]
    lib.orc_max_threads.restype = C.c_int
    return lib


def load_ref_legendre():
    """The unmodified reference LegendrePolynomials.cc (None if it was not built)."""
    path = os.path.join(_HERE, "_ref", "libref_legendre.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_compute_zfn.argtypes = [C.c_int, _dp]
    lib.ref_legendre_lat.argtypes = [C.c_int, C.c_double, _dp, _dp]
    lib.ref_legendre_tables.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _sp, _sp]
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = load_oracle()
    return _lib


def max_threads():
    return int(lib().orc_max_threads())


def gaussian_quadrature(N):
    lat = np.empty(2 * N)
    w = np.empty(2 * N)
    lib().orc_gaussian_quadrature(N, _p(lat), _p(w))
    return lat, w


def legendre_lat(trc, lat_rad, ref=False):
    n = (trc + 2) * (trc + 1) // 2
    zfn = np.zeros((trc + 1) * (trc + 1))
    out = np.zeros(n)
    if ref:
        r = load_ref_legendre()
        r.ref_compute_zfn(trc, _p(zfn))
        r.ref_legendre_lat(trc, lat_rad, _p(out), _p(zfn))
    else:
        lib().orc_compute_zfn(trc, _p(zfn))
        lib().orc_legendre_lat(trc, lat_rad, _p(out), _p(zfn))
    return out


def idx_mn(trc, m, n):
    return (2 * trc + 3 - m) * m // 2 + n - m


class OraclePlan:
    """State of the TransLocal constructor for a global structured grid (TransLocal.cc:322-770)."""

    def __init__(self, nx, lat_deg, truncation, regular=False, weights=None, nthreads=0):
        self.nx = np.ascontiguousarray(nx, dtype=np.int32)
        self.lat = np.ascontiguousarray(lat_deg, dtype=np.float64)
        self.T = int(truncation)
        self.w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        if nthreads <= 0:
            nthreads = max_threads()
        self.h = lib().orc_plan_create(
            self.nx.size, self.nx.ctypes.data_as(_ip), _p(self.lat), self.T, 1 if regular else 0,
            None if self.w is None else _p(self.w), nthreads)
        self.npts = int(lib().orc_plan_npts(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_plan_destroy(self.h)
            self.h = None

    def nlat0(self):
        out = np.empty(self.T + 1, dtype=np.int32)
        lib().orc_plan_nlat0(self.h, out.ctypes.data_as(_ip))
        return out

    def tables(self):
        sb = np.zeros(self.T + 3, dtype=np.uintp)
        ab = np.zeros(self.T + 3, dtype=np.uintp)
        lib().orc_plan_table_sizes(self.h, sb.ctypes.data_as(_sp), ab.ctypes.data_as(_sp))
        sym = np.ctypeslib.as_array(lib().orc_plan_leg_sym(self.h), shape=(int(sb[-1]),)).copy()
        asym = np.ctypeslib.as_array(lib().orc_plan_leg_asym(self.h), shape=(int(ab[-1]),)).copy()
        return sym, asym, sb, ab

    def invtrans(self, nb_scalar, scalar_spectra, nb_vordiv=0, vor=None, div=None, mode=2):
        nall = nb_scalar + 2 * nb_vordiv
        gp = np.zeros(nall * self.npts)
        lib().orc_invtrans(self.h, nb_scalar, None if scalar_spectra is None else _p(scalar_spectra), nb_vordiv,
                           None if vor is None else _p(vor), None if div is None else _p(div), _p(gp), mode)
        return gp

    def invtrans_legendre(self, truncation, nf, spectra, fast=True):
        out = np.zeros(nf * 2 * self.nx.size * (self.T + 1))
        lib().orc_invtrans_legendre(self.h, truncation, nf, _p(spectra), _p(out), 1 if fast else 0)
        return out

    def dirtrans(self, nf, gp):
        sp = np.zeros((self.T + 1) * (self.T + 2) * nf)
        lib().orc_dirtrans(self.h, nf, _p(np.ascontiguousarray(gp)), _p(sp))
        return sp


    def dirtrans_wind(self, nf, wind):
        n = (self.T + 1) * (self.T + 2) * nf
        vor, div = np.zeros(n), np.zeros(n)
        lib().orc_dirtrans_wind(self.h, nf, _p(np.ascontiguousarray(wind)), _p(vor), _p(div))
        return vor, div

    def invtrans_grad(self, nf, spectra):
        grad = np.zeros(2 * nf * self.npts)
        lib().orc_invtrans_grad(self.h, nf, _p(np.ascontiguousarray(spectra)), _p(grad))
        return grad


def invtrans_unstructured(truncation, nb_fields, nb_vordiv_fields, spectra, lon_deg, lat_deg, nthreads=0):
    """TransLocal::invtrans_unstructured (TransLocal.cc:1289-1392) at the given points; gp is [field][point]."""
    lon = np.ascontiguousarray(lon_deg, dtype=np.float64)
    lat = np.ascontiguousarray(lat_deg, dtype=np.float64)
    gp = np.zeros(nb_fields * lon.size)
    lib().orc_invtrans_unstructured(int(truncation), int(nb_fields), int(nb_vordiv_fields),
                                    _p(np.ascontiguousarray(spectra, dtype=np.float64)), lon.size, _p(lon), _p(lat), _p(gp),
                                    nthreads if nthreads > 0 else max_threads())
    return gp


def vd2uv(T, nf, vor, div):
    U = np.zeros_like(vor)
    V = np.zeros_like(vor)
    lib().orc_vd2uv(T, nf, _p(vor), _p(div), _p(U), _p(V))
    return U, V


def c2r(n, spec_complex, naive=False):
    a = np.ascontiguousarray(spec_complex, dtype=np.complex128)
    out = np.empty(n)
    lib().orc_c2r(n, a.view(np.float64).ctypes.data_as(_dp), _p(out), 1 if naive else 0)
    return out


def r2c(n, x):
    out = np.empty(n // 2 + 1, dtype=np.complex128)
    lib().orc_r2c(n, _p(np.ascontiguousarray(x, dtype=np.float64)), out.view(np.float64).ctypes.data_as(_dp))
    return out
