"""LegendreCacheCreator unique identifiers: sptrans_legendre_cache_uid against the 66 strings the reference's
own test expects (src/tests/trans/test_trans_localcache.cc:264-360, extracted by tests/golden/make_uid_golden.py).
Host-only: no GPU needed.  eckit::MD5 (third party, absent) is restated from RFC 1321 in host_setup.cc; these fixtures pin
it together with the way the reference feeds it (strings without terminator, bool as one byte, lround(lat * 1e8) as long)."""
import ctypes as C
import json
import os

import numpy as np

from atlas_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
GAUSSIAN, LONLAT, SHIFTED, REGIONAL, OTHER = range(5)


def uid(prefix, T, kind, n=0, south=0.0, north=0.0, lat=None, flt=False):
    buf = C.create_string_buffer(256)
    la = None if lat is None else np.ascontiguousarray(lat, dtype=np.float64)
    r = _lib.lib.sptrans_legendre_cache_uid(buf, 256, prefix.encode(), T, kind, n, south, north, 0 if la is None else la.size,
                                            None if la is None else la.ctypes.data_as(_lib.c_double_p), 1 if flt else 0)
    assert r == len(buf.value)
    return buf.value.decode()


def gaussian_lat(N):
    lat, w = np.empty(2 * N), np.empty(2 * N)
    _lib.check(_lib.lib.sptrans_gaussian_latitudes(N, lat.ctypes.data_as(_lib.c_double_p), w.ctypes.data_as(_lib.c_double_p)))
    return lat


def test_uids_match_reference_fixtures():
    fx = json.load(open(os.path.join(HERE, "golden", "legendre_cache_uids.json")))
    assert len(fx["cases"]) == 66
    lat_cache = {}
    for c in fx["cases"]:
        kind, N = c["grid"][0], int(c["grid"][1:])
        T = c["truncation"]
        if c["domain"] == "global":
            if kind in "FNO":   # any global Gaussian grid shares one cache (LegendreCacheCreatorLocal.cc:82-85)
                got = uid("local", T, GAUSSIAN, N)
            else:               # L<N>: regular lon-lat with pole rows, ny = 2N + 1 (:86-100)
                got = uid("local", T, LONLAT, 2 * N + 1)
        else:
            # the cropped grids are neither GaussianGrid nor RegularLonLatGrid for the reference: rows with |lat| <= 20,
            # identified by the hash of their latitudes (:70-73, :46-58)
            if kind in "FNO":
                if N not in lat_cache:
                    lat_cache[N] = gaussian_lat(N)
                lat = lat_cache[N][np.abs(lat_cache[N]) <= 20.0]
            else:
                full = 90.0 - 180.0 * np.arange(2 * N + 1) / (2 * N)
                lat = full[np.abs(full) <= 20.0 + 1e-9]
            got = uid("local", T, OTHER, lat=lat)
        assert got == c["uid"], (c, got)


def test_uid_other_kinds_and_errors():
    assert uid("local", 20, SHIFTED, 180) == "local-T20-S-ny180-OPT4189816c2e"
    assert uid("local", 20, REGIONAL, 41, -20.0, 20.0) == "local-T20-Regional-south-20-north20-ny41-OPT4189816c2e"
    assert uid("local", 20, GAUSSIAN, 320, flt=True).endswith("-OPT4446948bb5")   # eckit::MD5("flt" + '\x01')
    assert uid("b200", 159, GAUSSIAN, 160).startswith("b200-T159-GaussianN160-")
    assert _lib.lib.sptrans_legendre_cache_uid(None, 0, b"local", 1, 0, 1, 0.0, 0.0, 0, None, 0) == -1
    small = C.create_string_buffer(8)
    assert _lib.lib.sptrans_legendre_cache_uid(small, 8, b"local", 1, 0, 1, 0.0, 0.0, 0, None, 0) == -1
    assert _lib.lib.sptrans_legendre_cache_estimate(1279) == 1279 ** 3 // 2 * 8
