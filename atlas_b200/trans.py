"""Host-side mirror of `atlas::trans::Trans` (trans/Trans.h:42-189) over the sptrans C ABI.

Same method names, argument meaning and error behaviour as the reference's IFS-style raw-pointer
interface (trans/detail/TransImpl.h:116-181), so that the parity tests read like
src/tests/trans/test_transgeneral.cc.  Arrays may be NumPy arrays (host; staged by the library) or
torch CUDA tensors (used in place).  No arithmetic happens in this file.
"""
import ctypes as C

import numpy as np

from . import _lib
from .grid import CroppedGrid, Grid, StructuredGrid, UnstructuredGrid


class option:
    """`atlas::option::*` sugar (option/TransOptions.h): only `type` matters here."""

    @staticmethod
    def type(name):
        return {"type": name}


def _ptr(a):
    """Raw pointer of a NumPy array or torch tensor (fp64, contiguous)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
            raise ValueError("expected a C-contiguous float64 array")
        return C.c_void_p(a.ctypes.data)
    # torch tensor (duck-typed so that torch stays optional at import time)
    if hasattr(a, "data_ptr"):
        import torch

        if a.dtype != torch.float64 or not a.is_contiguous():
            raise ValueError("expected a contiguous float64 tensor")
        return C.c_void_p(a.data_ptr())
    raise TypeError(f"unsupported array type {type(a)}")


def _wait_for_producers(*arrays):
    """The plan runs on its own (non-blocking) CUDA stream and the calls are blocking like the reference's; device
    buffers handed in as torch tensors may still be being written by kernels queued on torch's current stream, so
    that stream is drained first (one host synchronisation, only when a CUDA tensor is passed)."""
    for a in arrays:
        if a is not None and hasattr(a, "data_ptr") and getattr(a, "is_cuda", False):
            import torch

            torch.cuda.current_stream(a.device).synchronize()
            return


class Trans:
    """`trans::Trans(grid, truncation, option::type("b200"))`."""

    def __init__(self, grid, truncation, config=None, device=0, rank=0, nranks=1, local_io=False):
        if isinstance(grid, str):
            grid = Grid(grid)
        if not isinstance(grid, (StructuredGrid, UnstructuredGrid, CroppedGrid)):
            raise TypeError("Trans needs a StructuredGrid, a CroppedGrid or an UnstructuredGrid")
        cfg = dict(config or {})
        backend = cfg.get("type", "b200")
        if backend != "b200":
            raise ValueError(f"backend {backend!r} not available here; this package provides type('b200') only")
        self._grid = grid
        self._T = int(truncation)
        self._h = C.c_void_p()
        if isinstance(grid, UnstructuredGrid):  # TransLocal's unstructured path (TransLocal.cc:740-770, :1289-1392)
            if nranks != 1:
                raise ValueError("point-set plans are not sharded")
            lon, lat = grid.lonlat()
            _lib.check(_lib.lib.sptrans_plan_create_points(C.byref(self._h), lon.size, lon.ctypes.data_as(_lib.c_double_p),
                                                           lat.ctypes.data_as(_lib.c_double_p), self._T, int(device)))
            return
        if isinstance(grid, CroppedGrid):   # Trans(global_grid, domain, truncation): TransLocal.cc:371-531
            if nranks != 1:
                raise ValueError("cropped-grid plans are not sharded")
            g = grid.global_grid
            gnx, glat = g.nx(), g.y()
            cnx, start = grid.nx(), grid.jlon_min()
            _lib.check(_lib.lib.sptrans_plan_create_cropped(
                C.byref(self._h), g.ny(), gnx.ctypes.data_as(_lib.c_int_p), glat.ctypes.data_as(_lib.c_double_p), self._T,
                1 if g.regular else 0, int(device), int(grid.jlat_min), grid.ny(), cnx.ctypes.data_as(_lib.c_int_p),
                start.ctypes.data_as(_lib.c_int_p)))
            return
        nx = grid.nx()
        lat = grid.y()
        w = grid.weights()
        flags = (1 if grid.regular else 0) | (4 if local_io else 0)  # SPTRANS_GRID_REGULAR | SPTRANS_SHARD_LOCAL_IO
        _lib.check(
            _lib.lib.sptrans_plan_create_sharded(
                C.byref(self._h), grid.ny(), nx.ctypes.data_as(_lib.c_int_p), lat.ctypes.data_as(_lib.c_double_p),
                None if w is None else w.ctypes.data_as(_lib.c_double_p), self._T, flags, int(device), int(rank), int(nranks),
            )
        )

    def clone(self):
        """A second plan on the same device that borrows this plan's tables (sptrans_plan_clone): one transform can be in
        flight on each.  The clone keeps this object alive."""
        other = object.__new__(Trans)
        other._grid, other._T, other._parent = self._grid, self._T, self
        other._h = C.c_void_p()
        _lib.check(_lib.lib.sptrans_plan_clone(self._h, C.byref(other._h)))
        return other

    def set_async(self, on=True):
        """Whole-transform calls return after enqueueing; call synchronize() before touching the buffers."""
        _lib.check(_lib.lib.sptrans_set_async(self._h, 1 if on else 0))

    def synchronize(self):
        _lib.check(_lib.lib.sptrans_synchronize(self._h))

    def mark(self):
        """End of everything enqueued on this plan so far (sptrans_mark); hand it to another plan's wait_mark."""
        m = C.c_void_p()
        _lib.check(_lib.lib.sptrans_mark(self._h, C.byref(m)))
        return m

    def wait_mark(self, mark):
        _lib.check(_lib.lib.sptrans_wait_mark(self._h, mark))

    @staticmethod
    def release_mark(mark):
        _lib.lib.sptrans_release_mark(mark)

    def local_sizes(self):
        """(spectral doubles per field, grid points per field) of the arrays the sharded entry points address."""
        a, b = C.c_size_t(), C.c_size_t()
        _lib.check(_lib.lib.sptrans_local_sizes(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and _lib is not None and getattr(_lib, "lib", None) is not None:
            _lib.lib.sptrans_plan_destroy(h)  # (module globals may already be gone at interpreter shutdown)
            self._h = C.c_void_p()

    # --- inspectors (TransImpl.h:42-52) ---
    def type(self):
        return "b200"

    def truncation(self):
        return self._T

    def grid(self):
        return self._grid

    def nb_spectral_coefficients(self):
        return int(_lib.lib.sptrans_nb_spectral_coefficients(self._h))

    def nb_gridpoints(self):
        return int(_lib.lib.sptrans_nb_gridpoints(self._h))

    def nlat0(self):
        out = np.empty(self._T + 1, dtype=np.int32)
        _lib.check(_lib.lib.sptrans_get_nlat0(self._h, out.ctypes.data_as(_lib.c_int_p)))
        return out

    def device_bytes(self):
        return int(_lib.lib.sptrans_device_bytes(self._h))

    def kernel_launches(self):
        return int(_lib.lib.sptrans_kernel_launches(self._h))

    def last_timings(self):
        out = (C.c_float * 8)()
        _lib.check(_lib.lib.sptrans_last_timings(self._h, out))
        t = list(out)
        return {"pack": t[0], "legendre": t[1], "fourier": t[2], "h2d": t[3], "d2h": t[4], "exchange_wait": t[5], "repack": t[6]}

    def set_precision(self, name):
        """'fp64' (default, DMMA) or 'tc' (tcgen05 split-TF32 Legendre stage, fp32-level accuracy)."""
        code = {"fp64": 0, "tc": 1}[name]
        _lib.check(_lib.lib.sptrans_set_precision(self._h, code))

    def set_stream(self, cuda_stream_ptr):
        self._external_stream = True  # stream-ordered use (dist.py): the caller orders producers on that stream
        _lib.check(_lib.lib.sptrans_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def export_legendre_cache(self):
        n = int(_lib.lib.sptrans_legendre_cache_size(self._h))
        out = np.empty(n // 8, dtype=np.float64)
        _lib.check(_lib.lib.sptrans_export_legendre_cache(self._h, C.c_void_p(out.ctypes.data)))
        return out

    def import_legendre_cache(self, blob):
        """Replace the Legendre tables by a reference-layout cache blob (`Cache::legendre()`, trans/Cache.h:98-136)."""
        blob = np.ascontiguousarray(blob, dtype=np.float64)
        _lib.check(_lib.lib.sptrans_import_legendre_cache(self._h, C.c_void_p(blob.ctypes.data), blob.nbytes))

    def _sync(self, *arrays):
        if not getattr(self, "_external_stream", False):
            _wait_for_producers(*arrays)

    # --- transforms, IFS-style buffers (TransImpl.h:116-181) ---
    def invtrans(self, *args):
        """invtrans(nb_scalar, scalar_spectra, gp)
        invtrans(nb_vordiv, vor, div, gp)
        invtrans(nb_scalar, scalar_spectra, nb_vordiv, vor, div, gp)"""
        self._sync(*args)
        if len(args) == 3:
            n, sp, gp = args
            _lib.check(_lib.lib.sptrans_invtrans_scalar(self._h, int(n), _ptr(sp), _ptr(gp)))
        elif len(args) == 4:
            n, vor, div, gp = args
            _lib.check(_lib.lib.sptrans_invtrans_vordiv2wind(self._h, int(n), _ptr(vor), _ptr(div), _ptr(gp)))
        elif len(args) == 6:
            ns, sp, nv, vor, div, gp = args
            _lib.check(_lib.lib.sptrans_invtrans(self._h, int(ns), _ptr(sp), int(nv), _ptr(vor), _ptr(div), _ptr(gp)))
        else:
            raise TypeError("invtrans: wrong number of arguments")

    def dirtrans(self, *args):
        """dirtrans(nb_fields, scalar_fields, scalar_spectra)
        dirtrans(nb_fields, wind_fields, vorticity_spectra, divergence_spectra)"""
        self._sync(*args)
        if len(args) == 3:
            n, gp, sp = args
            _lib.check(_lib.lib.sptrans_dirtrans_scalar(self._h, int(n), _ptr(gp), _ptr(sp)))
        elif len(args) == 4:
            n, wind, vor, div = args
            _lib.check(_lib.lib.sptrans_dirtrans_wind2vordiv(self._h, int(n), _ptr(wind), _ptr(vor), _ptr(div)))
        else:
            raise TypeError("dirtrans: wrong number of arguments")

    def invtrans_adj(self, *args):
        """Adjoints of the three invtrans overloads, <invtrans x, y> == <x, invtrans_adj y>  (TransImpl.h:147-166):
        invtrans_adj(nb_scalar, gp, scalar_spectra)
        invtrans_adj(nb_vordiv, wind, vor, div)
        invtrans_adj(nb_scalar, gp, nb_vordiv, vor, div, scalar_spectra)"""
        self._sync(*args)
        if len(args) == 3:
            n, gp, sp = args
            _lib.check(_lib.lib.sptrans_invtrans_adj_scalar(self._h, int(n), _ptr(gp), _ptr(sp)))
        elif len(args) == 4:
            n, wind, vor, div = args
            _lib.check(_lib.lib.sptrans_invtrans_vordiv2wind_adj(self._h, int(n), _ptr(wind), _ptr(vor), _ptr(div)))
        elif len(args) == 6:
            ns, gp, nv, vor, div, sp = args
            _lib.check(_lib.lib.sptrans_invtrans_adj(self._h, int(ns), _ptr(gp), int(nv), _ptr(vor), _ptr(div), _ptr(sp)))
        else:
            raise TypeError("invtrans_adj: wrong number of arguments")

    def invtrans_grad_adj(self, nb_fields, grad_fields, scalar_spectra):
        """adjoint of invtrans_grad (TransImpl.h:93-97)"""
        self._sync(grad_fields, scalar_spectra)
        _lib.check(_lib.lib.sptrans_invtrans_grad_adj(self._h, int(nb_fields), _ptr(grad_fields), _ptr(scalar_spectra)))

    def dirtrans_adj(self, nb_fields, scalar_spectra, gp_fields):
        """adjoint of dirtrans(nb_fields, gp, spectra) (TransImpl.h:63-67): spectra in, grid fields out"""
        self._sync(scalar_spectra, gp_fields)
        _lib.check(_lib.lib.sptrans_dirtrans_adj_scalar(self._h, int(nb_fields), _ptr(scalar_spectra), _ptr(gp_fields)))

    def dirtrans_wind2vordiv_adj(self, nb_fields, vorticity_spectra, divergence_spectra, wind_fields):
        """adjoint of dirtrans(nb_fields, wind, vor, div) (TransImpl.h:69-70): vor/div spectra in, wind fields out"""
        self._sync(vorticity_spectra, divergence_spectra, wind_fields)
        _lib.check(_lib.lib.sptrans_dirtrans_wind2vordiv_adj(self._h, int(nb_fields), _ptr(vorticity_spectra),
                                                             _ptr(divergence_spectra), _ptr(wind_fields)))

    def dirtrans_wind2vordiv_adj_field(self, spvor, spdiv, gpwind):
        self._sync(spvor, spdiv, gpwind)
        nlev = self._nlev(spvor, gpwind, 2)
        self._nlev(spdiv, gpwind, 2)
        _lib.check(_lib.lib.sptrans_dirtrans_wind2vordiv_adj_field(self._h, nlev, _ptr(spvor), _ptr(spdiv), _ptr(gpwind)))

    # --- atlas Field layouts (TransImpl.h:54-100): spectral (nspec2, nlev); grid (npts, nlev) or (npts, nlev, 2) ---
    def _nlev(self, sp, gp, ncomp):
        """Number of levels from the array shapes, checked like the reference's Field overloads check ranks/sizes."""
        nspec2, npts = self.nb_spectral_coefficients(), self.nb_gridpoints()
        sshape, gshape = tuple(sp.shape), tuple(gp.shape)
        if len(sshape) != 2 or sshape[0] != nspec2:
            raise ValueError(f"spectral field must have shape ({nspec2}, nlev), got {sshape}")
        nlev = int(sshape[1])
        want = (npts, nlev) if ncomp == 1 else (npts, nlev, ncomp)
        if gshape != want:
            raise ValueError(f"grid-point field must have shape {want}, got {gshape}")
        return nlev

    def invtrans_field(self, spfield, gpfield):
        self._sync(spfield, gpfield)
        _lib.check(_lib.lib.sptrans_invtrans_field(self._h, self._nlev(spfield, gpfield, 1), _ptr(spfield), _ptr(gpfield)))

    def dirtrans_field(self, gpfield, spfield):
        self._sync(spfield, gpfield)
        _lib.check(_lib.lib.sptrans_dirtrans_field(self._h, self._nlev(spfield, gpfield, 1), _ptr(gpfield), _ptr(spfield)))

    def invtrans_adj_field(self, gpfield, spfield):
        self._sync(spfield, gpfield)
        _lib.check(_lib.lib.sptrans_invtrans_adj_field(self._h, self._nlev(spfield, gpfield, 1), _ptr(gpfield), _ptr(spfield)))

    def invtrans_vordiv2wind_field(self, spvor, spdiv, gpwind):
        self._sync(spvor, spdiv, gpwind)
        nlev = self._nlev(spvor, gpwind, 2)
        self._nlev(spdiv, gpwind, 2)
        _lib.check(_lib.lib.sptrans_invtrans_vordiv2wind_field(self._h, nlev, _ptr(spvor), _ptr(spdiv), _ptr(gpwind)))

    def dirtrans_wind2vordiv_field(self, gpwind, spvor, spdiv):
        self._sync(spvor, spdiv, gpwind)
        nlev = self._nlev(spvor, gpwind, 2)
        self._nlev(spdiv, gpwind, 2)
        _lib.check(_lib.lib.sptrans_dirtrans_wind2vordiv_field(self._h, nlev, _ptr(gpwind), _ptr(spvor), _ptr(spdiv)))

    def dirtrans_adj_field(self, spfield, gpfield):
        self._sync(spfield, gpfield)
        _lib.check(_lib.lib.sptrans_dirtrans_adj_field(self._h, self._nlev(spfield, gpfield, 1), _ptr(spfield), _ptr(gpfield)))

    def invtrans_vordiv2wind_adj_field(self, gpwind, spvor, spdiv):
        self._sync(spvor, spdiv, gpwind)
        nlev = self._nlev(spvor, gpwind, 2)
        self._nlev(spdiv, gpwind, 2)
        _lib.check(_lib.lib.sptrans_invtrans_vordiv2wind_adj_field(self._h, nlev, _ptr(gpwind), _ptr(spvor), _ptr(spdiv)))

    def invtrans_grad_adj_field(self, gradfield, spfield):
        self._sync(spfield, gradfield)
        _lib.check(_lib.lib.sptrans_invtrans_grad_adj_field(self._h, self._nlev(spfield, gradfield, 2), _ptr(gradfield), _ptr(spfield)))

    def invtrans_grad_field(self, spfield, gradfield):
        self._sync(spfield, gradfield)
        _lib.check(_lib.lib.sptrans_invtrans_grad_field(self._h, self._nlev(spfield, gradfield, 2), _ptr(spfield), _ptr(gradfield)))

    def invtrans_grad(self, nb_fields, scalar_spectra, grad_fields):
        """grad_fields = [E-W_1..E-W_k | N-S_1..N-S_k][npts]  (TransIFS::__invtrans_grad, ifs/TransIFS.cc:2075-2142)"""
        self._sync(scalar_spectra, grad_fields)
        _lib.check(_lib.lib.sptrans_invtrans_grad(self._h, int(nb_fields), _ptr(scalar_spectra), _ptr(grad_fields)))

    # --- stage level (device pointers only) ---
    def fourier_elems_per_field(self):
        return int(_lib.lib.sptrans_fourier_elems_per_field(self._h))

    def fourier_paths(self):
        """Grid points and exchange-buffer rows (per field) served by each family of Fourier kernels, and the share of the
        stage's algorithmic bytes (8 B per grid point + 16 B per exchange row) on each."""
        out = (C.c_longlong * 8)()
        _lib.check(_lib.lib.sptrans_fourier_path_stats(self._h, C.cast(out, C.c_void_p)))
        names = ["direct_mixed_radix", "chirpz_register_tiled", "chirpz_smem_passes", "chirpz_row_mode"]
        byts = [8.0 * out[i] + 16.0 * out[4 + i] for i in range(4)]
        tot = max(sum(byts), 1.0)
        return {n: {"grid_points": int(out[i]), "exchange_rows": int(out[4 + i]), "byte_share": byts[i] / tot}
                for i, n in enumerate(names)}

    def invtrans_legendre(self, nf, trunc, d_spec, d_fourier):
        self._sync(d_spec, d_fourier)
        _lib.check(_lib.lib.sptrans_invtrans_legendre(self._h, int(nf), int(trunc), _ptr(d_spec), _ptr(d_fourier)))

    def invtrans_fourier(self, nf, mlimit, d_fourier, d_gp, nb_uv=0):
        self._sync(d_fourier, d_gp)
        _lib.check(_lib.lib.sptrans_invtrans_fourier(self._h, int(nf), int(mlimit), _ptr(d_fourier), _ptr(d_gp), int(nb_uv)))

    def dirtrans_fourier(self, nf, d_gp, d_fourier, nb_uv=0):
        self._sync(d_gp, d_fourier)
        _lib.check(_lib.lib.sptrans_dirtrans_fourier(self._h, int(nf), _ptr(d_gp), _ptr(d_fourier), int(nb_uv)))

    def dirtrans_legendre(self, nf, d_fourier, d_spec):
        self._sync(d_fourier, d_spec)
        _lib.check(_lib.lib.sptrans_dirtrans_legendre(self._h, int(nf), _ptr(d_fourier), _ptr(d_spec)))


class MultiTrans:
    """Single-process multi-GPU transform (sptrans_multi_*): what `Trans(grid, T, option::type("b200") | ("gpus", N))`
    selects in the C++ adaptor.  Global arrays in the reference's layouts, NumPy (host) or torch CUDA tensors; scalar
    inverse and direct transforms."""

    def __init__(self, grid, truncation, devices):
        if isinstance(grid, str):
            grid = Grid(grid)
        self._grid, self._T = grid, int(truncation)
        devs = np.ascontiguousarray(devices, dtype=np.int32)
        nx, lat, w = grid.nx(), grid.y(), grid.weights()
        self._h = C.c_void_p()
        _lib.check(_lib.lib.sptrans_multi_create(C.byref(self._h), grid.ny(), nx.ctypes.data_as(_lib.c_int_p),
                                                 lat.ctypes.data_as(_lib.c_double_p),
                                                 None if w is None else w.ctypes.data_as(_lib.c_double_p), self._T,
                                                 1 if grid.regular else 0, devs.size, devs.ctypes.data_as(_lib.c_int_p)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and _lib is not None and getattr(_lib, "lib", None) is not None:
            _lib.lib.sptrans_multi_destroy(h)
            self._h = C.c_void_p()

    def size(self):
        return int(_lib.lib.sptrans_multi_size(self._h))

    def kernel_launches(self):
        return sum(int(_lib.lib.sptrans_kernel_launches(C.c_void_p(_lib.lib.sptrans_multi_plan(self._h, r)))) for r in range(self.size()))

    def invtrans(self, nb_fields, scalar_spectra, gp_fields):
        _wait_for_producers(scalar_spectra, gp_fields)
        _lib.check(_lib.lib.sptrans_multi_invtrans_scalar(self._h, int(nb_fields), _ptr(scalar_spectra), _ptr(gp_fields)))

    def dirtrans(self, nb_fields, gp_fields, scalar_spectra):
        _wait_for_producers(scalar_spectra, gp_fields)
        _lib.check(_lib.lib.sptrans_multi_dirtrans_scalar(self._h, int(nb_fields), _ptr(gp_fields), _ptr(scalar_spectra)))


class VorDivToUV:
    """`trans::VorDivToUV(truncation, option::type("b200"))` (trans/VorDivToUV.h:109-128)."""

    def __init__(self, truncation, config=None, device=0):
        self._T = int(truncation)
        self._device = int(device)

    def truncation(self):
        return self._T

    def execute(self, nb_coeff, nb_fields, vorticity, divergence, U, V):
        _lib.check(_lib.lib.sptrans_vordiv_to_uv(self._T, int(nb_fields), _ptr(vorticity), _ptr(divergence), _ptr(U), _ptr(V), self._device))
