"""atlas_b200 -- B200-native spectral-transform engine behind the atlas::trans::Trans API.

The package holds only what the hot path needs: the CUDA library (csrc/ -> libsptrans_b200.so,
C ABI in include/sptrans_b200.h) and a thin host-side mirror of the reference's
`atlas::trans::Trans` / `atlas::Grid` interface (trans.py, grid.py) for tests and benchmarks.
"""
from .grid import CroppedGrid, Grid, StructuredGrid, UnstructuredGrid  # noqa: F401
from .trans import MultiTrans, Trans, VorDivToUV, option  # noqa: F401
