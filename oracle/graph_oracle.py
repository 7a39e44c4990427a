"""CPU restatement of the reference's graph build (TEST INFRASTRUCTURE ONLY).

Follows ``src/step2_graph_builder.py``:

* ``__convert_wgs2utm`` (:8-13): column 0 (latitude) -> ``R * ln(tan(pi/4 + lat*pi/360))``,
  column 1 (longitude) -> ``R * (lon*pi/180)``, ``R = 6378137`` — scalar libm ``log``/``tan``.
* ``build_graph`` (:16-40): ``A_hat[i][i] = 1``; ``A_hat[i][j] = 1/sqrt((X*X + Y*Y)/1e8)``
  with ``X = c_i[0]-c_j[0]``, ``Y = c_i[1]-c_j[1]`` (:24-31); ``D = diag(sum(A_hat, axis=0))``
  (:34); ``A* = D^-1/2 . A_hat . D^-1/2`` via ``scipy.linalg.fractional_matrix_power`` (:37-38).

Bit pattern of the normalisation (re-verified by
``tests/test_oracle.py`` against the reference's own ``build_graph`` output): SciPy
evaluates the -0.5 power of a diagonal matrix as ``inv(D) . diag(sqrt(d))``, so

    dh_i  = fl( fl(1/d_i) * fl(sqrt(d_i)) )
    A*_ij = fl( fl(dh_i * a_ij) * dh_j )

with ``d_j`` the column sums accumulated row by row (``np.sum(axis=0)`` on a C-ordered
matrix adds row 0, then row 1, ...).  All fp64; the caller casts to fp32
(``src/main.py:26``).

The kNN generator has no counterpart in the reference (it is the "new synthetic
large-graph generator" BASELINE.json asks for); its definition lives here and in
DESIGN.md, and the CUDA implementation must reproduce it bit for bit.
"""

from __future__ import annotations

import math

import numpy as np

EARTH_RADIUS = 6378137  # step2:9

# bounding box of the shipped stations (data/ACISStationCoordinates.csv:2-36)
LAT_RANGE = (50.10, 51.59)
LON_RANGE = (-113.36, -110.09)

_MASK = (1 << 64) - 1
_GOLDEN = 0x9E3779B97F4A7C15


def mercator(latlon):
    """step2:8-13.  ``latlon [S,2]`` degrees -> ``[S,2]`` metres (northing, easting)."""
    latlon = np.asarray(latlon, dtype=np.float64)
    out = np.empty_like(latlon)
    for i in range(latlon.shape[0]):
        out[i, 0] = EARTH_RADIUS * math.log(math.tan(math.pi / 4 + latlon[i, 0] * math.pi / 360))
        out[i, 1] = EARTH_RADIUS * (latlon[i, 1] * math.pi / 180)
    return out


def _a_hat(xy):
    """step2:24-31, vectorised with the same operation order (no FMA in NumPy)."""
    X = xy[:, 0][:, None] - xy[:, 0][None, :]
    Y = xy[:, 1][:, None] - xy[:, 1][None, :]
    Hm = (X * X + Y * Y) / 100000000
    with np.errstate(divide="ignore"):
        A = 1 / np.sqrt(Hm)
    np.fill_diagonal(A, 1.0)
    return A, Hm


def _normalise(A):
    """step2:34-38 with SciPy's evaluation order (see module docstring)."""
    S = A.shape[0]
    d = np.zeros(S, dtype=np.float64)
    for i in range(S):  # sequential row-order column sums
        d = d + A[i]
    dh = (1.0 / d) * np.sqrt(d)
    return (dh[:, None] * A) * dh[None, :]


def dense_graph_f64(latlon):
    """Reference ``build_graph`` on ``[S,2]`` (lat, lon) degrees -> fp64 ``[S,S]``."""
    A, _ = _a_hat(mercator(latlon))
    return _normalise(A)


def knn_pattern(Hm, k):
    """Boolean ``[S,S]`` edge mask: ``j in kNN(i) or i in kNN(j)``, plus the diagonal.

    Neighbours ranked by ``Hm[i, j]`` ascending (the squared planar distance / 1e8 that
    the edge weight is computed from), ties broken by the smaller index ``j``; a node
    is never its own neighbour.
    """
    S = Hm.shape[0]
    k = min(k, S - 1)
    key = Hm.copy()
    np.fill_diagonal(key, np.inf)
    order = np.argsort(key, axis=1, kind="stable")[:, :k]
    mask = np.zeros((S, S), dtype=bool)
    mask[np.arange(S)[:, None], order] = True
    mask |= mask.T
    np.fill_diagonal(mask, True)
    return mask


def knn_graph_f64(latlon, k=8):
    """Symmetrised-kNN variant of the reference build: same weights, same normalisation,
    zero where there is no edge.  Returns the dense fp64 matrix (use ``dense_to_csr``)."""
    A, Hm = _a_hat(mercator(latlon))
    mask = knn_pattern(Hm, k)
    A = np.where(mask, A, 0.0)
    return _normalise(A)


def dense_to_csr(A, dtype=np.float32):
    """CSR (indptr int32[S+1], indices int32[nnz] ascending per row, values[nnz])."""
    S = A.shape[0]
    nz = A != 0
    indptr = np.zeros(S + 1, dtype=np.int32)
    indptr[1:] = np.cumsum(nz.sum(axis=1))
    rows, cols = np.nonzero(nz)
    return indptr, cols.astype(np.int32), A[rows, cols].astype(dtype)


def _splitmix64_nth(seed, n):
    """n-th output (n >= 1) of a SplitMix64 stream seeded with ``seed``; vectorised over n."""
    n = np.asarray(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed & _MASK) + n * np.uint64(_GOLDEN)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def synthetic_coordinates(S, seed=0):
    """``[S,2]`` (lat, lon) degrees, uniform in the shipped stations' bounding box.

    Station i takes SplitMix64 outputs ``2i+1`` (lat) and ``2i+2`` (lon);
    ``u = (z >> 11) * 2**-53``; ``coord = lo + u * (hi - lo)`` (two roundings).
    """
    i = np.arange(S, dtype=np.uint64)
    u_lat = (_splitmix64_nth(seed, 2 * i + 1) >> np.uint64(11)).astype(np.float64) * 2.0**-53
    u_lon = (_splitmix64_nth(seed, 2 * i + 2) >> np.uint64(11)).astype(np.float64) * 2.0**-53
    lat = LAT_RANGE[0] + u_lat * (LAT_RANGE[1] - LAT_RANGE[0])
    lon = LON_RANGE[0] + u_lon * (LON_RANGE[1] - LON_RANGE[0])
    return np.stack([lat, lon], axis=1)
