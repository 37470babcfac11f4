"""Host-side mirror of the few atlas Grid getters the transform consumes.

Reference: `StructuredGrid::{ny, nx(j), nxmax, y(j), x(i,j), size}` (grid/detail/grid/Structured.h:298-312),
the named-grid builders O<N> / F<N> (grid/detail/grid/Gaussian.cc:101-172) and the Gaussian latitudes
(grid/detail/spacing/gaussian/Latitudes.cc).  Latitudes/weights are computed by the product library
(sptrans_gaussian_latitudes), not by the test oracle.
"""
import ctypes as C
import re

import numpy as np

from . import _lib


class StructuredGrid:
    def __init__(self, name, nx, lat_deg, weights=None, regular=False):
        self.name = name
        self._nx = np.ascontiguousarray(nx, dtype=np.int32)
        self._lat = np.ascontiguousarray(lat_deg, dtype=np.float64)
        self._w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self.regular = bool(regular)
        self._rowoff = np.concatenate([[0], np.cumsum(self._nx, dtype=np.int64)])

    # --- atlas::StructuredGrid getters ---
    def ny(self):
        return int(self._nx.size)

    def nx(self, j=None):
        return self._nx if j is None else int(self._nx[j])

    def nxmax(self):
        return int(self._nx.max())

    def y(self, j=None):
        return self._lat if j is None else float(self._lat[j])

    def x(self, i, j):
        return 360.0 * i / int(self._nx[j])  # xmin + i*dx with xmin = 0 (Structured.h:312)

    def size(self):
        return int(self._rowoff[-1])

    def weights(self):
        return self._w

    def rowoff(self):
        return self._rowoff

    def lonlat(self):
        """(lon_deg, lat_deg) of every point in grid order."""
        lon = np.concatenate([360.0 * np.arange(n) / n for n in self._nx])
        lat = np.repeat(self._lat, self._nx)
        return lon, lat


class CroppedGrid:
    """`atlas::Grid(global_grid, domain)` for a RectangularDomain: the regional grid TransLocal accepts, a cropping of a
    global structured grid (TransLocal.cc:371-531).  Rows whose latitude lies in [south, north] are kept; in each kept row
    the points whose longitude, normalised into [west, west + 360), lies in [west, east], ordered west -> east
    (grid/detail/grid/Structured.cc crops the same way).  `jlat_min`, `nx()` and `jlon_min()` are what the reference derives
    in its constructor (jlatMin_ :441-447, jlonMin_ :501-531)."""

    regular = False

    def __init__(self, global_grid, west, east, south, north, name=None):
        if not isinstance(global_grid, StructuredGrid):
            raise TypeError("CroppedGrid needs a global StructuredGrid")
        g = self.global_grid = global_grid
        self.regular = g.regular
        self.name = name or f"{g.name}[{west},{east}]x[{south},{north}]"
        eps = 1e-10
        rows = [j for j in range(g.ny()) if south - eps <= g.y(j) <= north + eps]
        if not rows or rows != list(range(rows[0], rows[-1] + 1)):
            raise ValueError("the domain keeps no (contiguous) latitude rows of the global grid")
        self.jlat_min = rows[0]
        nxc, start = [], []
        for j in rows:
            n = g.nx(j)
            lon = 360.0 * np.arange(n) / n
            rel = np.mod(lon - west, 360.0)
            rel[rel > 360.0 - eps] = 0.0
            keep = np.nonzero(rel <= (east - west) + eps)[0]
            if keep.size == 0:
                raise ValueError(f"the domain keeps no point of global row {j}")
            first = int(keep[np.argmin(rel[keep])])
            nxc.append(int(keep.size))
            start.append(first)
        self._nx = np.ascontiguousarray(nxc, dtype=np.int32)
        self._start = np.ascontiguousarray(start, dtype=np.int32)
        self._rows = rows

    def ny(self):
        return len(self._rows)

    def nx(self, j=None):
        return self._nx if j is None else int(self._nx[j])

    def jlon_min(self):
        return self._start

    def y(self, j=None):
        lat = self.global_grid.y()[self._rows]
        return lat if j is None else float(lat[j])

    def size(self):
        return int(self._nx.sum())

    def weights(self):
        return None

    def global_indices(self):
        """index into the global grid's point order of every point of the crop (the copy-out of TransLocal.cc:1180-1187)"""
        ro = self.global_grid.rowoff()
        out = []
        for r, j in enumerate(self._rows):
            n = self.global_grid.nx(j)
            out.append(ro[j] + (self._start[r] + np.arange(self._nx[r])) % n)
        return np.concatenate(out)

    def lonlat(self):
        lon, lat = self.global_grid.lonlat()
        idx = self.global_indices()
        return lon[idx], lat[idx]


class UnstructuredGrid:
    """`atlas::UnstructuredGrid(points)` (grid/detail/grid/Unstructured.h): a list of (lon, lat) points in degrees."""

    regular = False

    def __init__(self, lon_deg, lat_deg, name="unstructured"):
        self.name = name
        self._lon = np.ascontiguousarray(lon_deg, dtype=np.float64).reshape(-1)
        self._lat = np.ascontiguousarray(lat_deg, dtype=np.float64).reshape(-1)
        if self._lon.size != self._lat.size:
            raise ValueError("lon and lat must have the same length")

    def size(self):
        return int(self._lon.size)

    def lonlat(self):
        return self._lon, self._lat

    def weights(self):
        return None


def gaussian_latitudes(N):
    lat = np.empty(2 * N)
    w = np.empty(2 * N)
    _lib.check(_lib.lib.sptrans_gaussian_latitudes(N, lat.ctypes.data_as(_lib.c_double_p), w.ctypes.data_as(_lib.c_double_p)))
    return lat, w


def Grid(name):
    """`atlas::Grid("O1280")`-style factory for the global Gaussian grids the benchmarks use:
    O<N> octahedral reduced, F<N> regular Gaussian (nx = 4N), L<N> regular lon-lat incl. poles
    (2N+1 latitudes, nx = 4N; grid/detail/grid/LonLat.cc)."""
    m = re.fullmatch(r"([OoFfLl])(\d+)", name)
    if not m:
        raise ValueError(f"unsupported grid name {name!r} (supported: O<N>, F<N>, L<N>)")
    kind, N = m.group(1).upper(), int(m.group(2))
    if kind == "O":
        lat, w = gaussian_latitudes(N)
        nx = np.empty(2 * N, dtype=np.int32)
        _lib.check(_lib.lib.sptrans_octahedral_nx(N, nx.ctypes.data_as(_lib.c_int_p)))
        return StructuredGrid(name, nx, lat, w, regular=False)
    if kind == "F":
        lat, w = gaussian_latitudes(N)
        return StructuredGrid(name, np.full(2 * N, 4 * N, dtype=np.int32), lat, w, regular=True)
    # L<N>: equally spaced latitudes 90 .. -90 (no quadrature weights -> inverse transform only)
    lat = 90.0 - 180.0 * np.arange(2 * N + 1) / (2 * N)
    lat[N] = 0.0
    return StructuredGrid(name, np.full(2 * N + 1, 4 * N, dtype=np.int32), lat, None, regular=True)
