"""ctypes binding of the C ABI in include/sptrans_b200.h.

The shared library is the product; this module only loads it.  There is no Python/NumPy/torch
fallback for any transform: if the library (or a CUDA device) is missing the calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPTRANS_LIB", os.path.join(_HERE, "libsptrans_b200.so"))  # override: tuning variants only

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class SptransError(RuntimeError):
    """Mirrors the C++ exceptions the reference throws (runtime/Exception.h:23-76)."""

    def __init__(self, code, msg):
        super().__init__(f"sptrans error {code}: {msg}")
        self.code = code


class NotImplementedInBackend(SptransError):
    """ATLAS_NOTIMPLEMENTED equivalent."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  atlas_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp = C.c_void_p
    sig = {
        "sptrans_last_error": (C.c_char_p, []),
        "sptrans_device_count": (C.c_int, []),
        "sptrans_gaussian_latitudes": (C.c_int, [C.c_int, c_double_p, c_double_p]),
        "sptrans_octahedral_nx": (C.c_int, [C.c_int, c_int_p]),
        "sptrans_fourier_truncation": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]),
        "sptrans_plan_create": (C.c_int, [C.POINTER(vp), C.c_int, c_int_p, c_double_p, c_double_p, C.c_int, C.c_uint, C.c_int]),
        "sptrans_plan_create_sharded": (C.c_int, [C.POINTER(vp), C.c_int, c_int_p, c_double_p, c_double_p, C.c_int, C.c_uint, C.c_int, C.c_int, C.c_int]),
        "sptrans_plan_create_points": (C.c_int, [C.POINTER(vp), C.c_size_t, c_double_p, c_double_p, C.c_int, C.c_int]),
        "sptrans_plan_create_cropped": (C.c_int, [C.POINTER(vp), C.c_int, c_int_p, c_double_p, C.c_int, C.c_uint, C.c_int, C.c_int, C.c_int,
                                                 c_int_p, c_int_p]),
        "sptrans_plan_destroy": (C.c_int, [vp]),
        "sptrans_truncation": (C.c_int, [vp]),
        "sptrans_nb_gridpoints": (C.c_size_t, [vp]),
        "sptrans_nb_spectral_coefficients": (C.c_size_t, [vp]),
        "sptrans_get_nlat0": (C.c_int, [vp, c_int_p]),
        "sptrans_device_bytes": (C.c_size_t, [vp]),
        "sptrans_legendre_cache_size": (C.c_size_t, [vp]),
        "sptrans_export_legendre_cache": (C.c_int, [vp, vp]),
        "sptrans_import_legendre_cache": (C.c_int, [vp, vp, C.c_size_t]),
        "sptrans_legendre_cache_uid": (C.c_int, [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                                C.c_int, c_double_p, C.c_int]),
        "sptrans_legendre_cache_estimate": (C.c_size_t, [C.c_int]),
        "sptrans_set_stream": (C.c_int, [vp, vp]),
        "sptrans_set_precision": (C.c_int, [vp, C.c_int]),
        "sptrans_invtrans_scalar": (C.c_int, [vp, C.c_int, vp, vp]),
        "sptrans_invtrans": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, vp, vp]),
        "sptrans_invtrans_vordiv2wind": (C.c_int, [vp, C.c_int, vp, vp, vp]),
        "sptrans_dirtrans_scalar": (C.c_int, [vp, C.c_int, vp, vp]),
        "sptrans_invtrans_adj_scalar": (C.c_int, [vp, C.c_int, vp, vp]),
        "sptrans_dirtrans_wind2vordiv": (C.c_int, [vp, C.c_int, vp, vp, vp]),
        "sptrans_invtrans_grad": (C.c_int, [vp, C.c_int, vp, vp]),
        "sptrans_invtrans_adj": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, vp, vp]),
        "sptrans_invtrans_vordiv2wind_adj": (C.c_int, [vp, C.c_int, vp, vp, vp]),
        "sptrans_invtrans_grad_adj": (C.c_int, [vp, C.c_int, vp, vp]),
        "sptrans_dirtrans_adj_scalar": (C.c_int, [vp, C.c_int, vp, vp]),
        "sptrans_dirtrans_wind2vordiv_adj": (C.c_int, [vp, C.c_int, vp, vp, vp]),
        "sptrans_dirtrans_wind2vordiv_adj_field": (C.c_int, [vp, C.c_int, vp, vp, vp]),
        "sptrans_set_async": (C.c_int, [vp, C.c_int]),
        "sptrans_synchronize": (C.c_int, [vp]),
        "sptrans_mark": (C.c_int, [vp, C.POINTER(vp)]),
        "sptrans_wait_mark": (C.c_int, [vp, vp]),
        "sptrans_release_mark": (C.c_int, [vp]),
        "sptrans_plan_clone": (C.c_int, [vp, C.POINTER(vp)]),
        "sptrans_local_sizes": (C.c_int, [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
        "sptrans_invtrans_field": (C.c_int, [vp, C.c_int, vp, vp]),
        "sptrans_dirtrans_field": (C.c_int, [vp, C.c_int, vp, vp]),
        "sptrans_invtrans_adj_field": (C.c_int, [vp, C.c_int, vp, vp]),
        "sptrans_invtrans_vordiv2wind_field": (C.c_int, [vp, C.c_int, vp, vp, vp]),
        "sptrans_dirtrans_wind2vordiv_field": (C.c_int, [vp, C.c_int, vp, vp, vp]),
        "sptrans_invtrans_grad_field": (C.c_int, [vp, C.c_int, vp, vp]),
        "sptrans_dirtrans_adj_field": (C.c_int, [vp, C.c_int, vp, vp]),
        "sptrans_invtrans_vordiv2wind_adj_field": (C.c_int, [vp, C.c_int, vp, vp, vp]),
        "sptrans_invtrans_grad_adj_field": (C.c_int, [vp, C.c_int, vp, vp]),
        "sptrans_vordiv_to_uv": (C.c_int, [C.c_int, C.c_int, vp, vp, vp, vp, C.c_int]),
        "sptrans_fourier_elems_per_field": (C.c_size_t, [vp]),
        "sptrans_fourier_path_stats": (C.c_int, [vp, vp]),
        "sptrans_invtrans_legendre": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
        "sptrans_invtrans_fourier": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, C.c_int]),
        "sptrans_dirtrans_fourier": (C.c_int, [vp, C.c_int, vp, vp, C.c_int]),
        "sptrans_dirtrans_legendre": (C.c_int, [vp, C.c_int, vp, vp]),
        "sptrans_shard_layout": (C.c_int, [C.c_int, c_int_p, c_double_p, C.c_int, C.c_uint, C.c_int, C.c_int, c_int_p, c_int_p,
                                           C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
        "sptrans_shard_segments": (C.c_longlong, [C.c_int, c_int_p, c_double_p, C.c_int, C.c_uint, C.c_int, C.c_int, C.c_int,
                                                  C.POINTER(C.c_longlong)]),
        "sptrans_exchange_rows": (C.c_int, [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
        "sptrans_exchange_pack": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
        "sptrans_exchange_unpack": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
        "sptrans_peer_alloc": (C.c_int, [vp, C.c_int, C.c_char_p]),
        "sptrans_peer_attach_ipc": (C.c_int, [vp, C.c_int, C.c_char_p]),
        "sptrans_peer_attach_ptrs": (C.c_int, [vp, C.c_int, C.POINTER(vp)]),
        "sptrans_peer_region": (C.c_int, [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "sptrans_peer_buffer": (C.c_int, [vp, C.POINTER(vp)]),
        "sptrans_peer_free": (C.c_int, [vp]),
        "sptrans_invtrans_sharded": (C.c_int, [vp, C.c_int, vp, vp]),
        "sptrans_dirtrans_sharded": (C.c_int, [vp, C.c_int, vp, vp]),
        "sptrans_invtrans_legendre_peers": (C.c_int, [vp, C.c_int, vp]),
        "sptrans_dirtrans_fourier_peers": (C.c_int, [vp, C.c_int, vp]),
        "sptrans_dirtrans_fourier_local": (C.c_int, [vp, C.c_int, vp]),
        "sptrans_dirtrans_legendre_pull": (C.c_int, [vp, C.c_int, vp]),
        "sptrans_peer_barrier": (C.c_int, [vp]),
        "sptrans_peer_advance": (C.c_int, [vp]),
        "sptrans_multi_create": (C.c_int, [C.POINTER(vp), C.c_int, c_int_p, c_double_p, c_double_p, C.c_int, C.c_uint, C.c_int, c_int_p]),
        "sptrans_multi_destroy": (C.c_int, [vp]),
        "sptrans_multi_size": (C.c_int, [vp]),
        "sptrans_multi_plan": (vp, [vp, C.c_int]),
        "sptrans_multi_invtrans_scalar": (C.c_int, [vp, C.c_int, vp, vp]),
        "sptrans_multi_dirtrans_scalar": (C.c_int, [vp, C.c_int, vp, vp]),
        "sptrans_last_timings": (C.c_int, [vp, C.POINTER(C.c_float)]),
        "sptrans_kernel_launches": (C.c_uint64, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib, sig


lib, SIGNATURES = _load()


def check(rc):
    if rc != 0:
        msg = lib.sptrans_last_error().decode("utf-8", "replace")
        if rc == 3:
            raise NotImplementedInBackend(rc, msg)
        raise SptransError(rc, msg)
