"""Graph build on the GPU, bit-identical to the reference's ``build_graph``.

Reference: src/step2_graph_builder.py:8-40 (and the fp32 cast at src/main.py:26).
The O(S) Mercator projection uses libm ``log``/``tan`` exactly as the reference does
(step2:8-13) and stays on the host; the O(S^2) weight matrix, degree sums and symmetric
normalisation run in fp64 on the device (``csrc/graph.cu``, compiled without FMA contraction).
"""

from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib

EARTH_RADIUS = 6378137  # step2:9


def mercator(latlon) -> np.ndarray:
    """step2:8-13: (lat, lon) degrees -> (northing, easting) metres, row by row."""
    latlon = np.asarray(latlon, dtype=np.float64)
    if latlon.ndim != 2 or latlon.shape[1] != 2:
        raise ValueError("latlon must be [S, 2]")
    out = np.empty_like(latlon)
    for i in range(latlon.shape[0]):
        out[i, 0] = EARTH_RADIUS * math.log(math.tan(math.pi / 4 + latlon[i, 0] * math.pi / 360))
        out[i, 1] = EARTH_RADIUS * (latlon[i, 1] * math.pi / 180)
    return out


def _device(device):
    dev = torch.device("cuda" if device is None else device)
    if dev.type != "cuda":
        raise RuntimeError("windgnn_b200 graph build is CUDA-only (no CPU fallback)")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _build(xy: torch.Tensor, k: int, want64: bool, want32: bool):
    lib = _lib.load()
    if not xy.is_cuda or xy.dtype != torch.float64 or xy.dim() != 2 or xy.shape[1] != 2:
        raise RuntimeError("xy must be a CUDA float64 tensor [S, 2]")
    xy = xy.contiguous()
    S = xy.shape[0]
    dev = xy.device
    a64 = torch.empty((S, S), dtype=torch.float64, device=dev) if want64 else None
    a32 = torch.empty((S, S), dtype=torch.float32, device=dev) if want32 else None
    nbytes = lib.wg_build_graph_workspace_bytes(S, k)
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
    _lib.check(
        lib.wg_build_graph_f64(
            xy.data_ptr(), a64.data_ptr() if want64 else None, a32.data_ptr() if want32 else None,
            S, k, ws.data_ptr(), ws.numel(), dev.index or 0, torch.cuda.current_stream(dev).cuda_stream,
        )
    )
    torch.cuda.current_stream(dev).synchronize()  # ws is freed on return
    return a64, a32


def build_graph(xy, device=None, dtype=torch.float32, k: int = 0) -> torch.Tensor:
    """Normalised adjacency from projected coordinates ``xy [S, 2]`` (metres).

    ``dtype=torch.float64`` returns the reference's ``build_graph`` output bit for bit;
    ``torch.float32`` its ``torch.tensor(...).float()`` cast (main.py:26).  ``k > 0`` keeps only
    the symmetrised k-nearest-neighbour edges (synthetic large-graph generator).
    """
    dev = _device(device)
    if not isinstance(xy, torch.Tensor):
        xy = torch.from_numpy(np.ascontiguousarray(np.asarray(xy, dtype=np.float64)))
    xy = xy.to(device=dev, dtype=torch.float64)
    if dtype == torch.float64:
        return _build(xy, k, True, False)[0]
    if dtype == torch.float32:
        return _build(xy, k, False, True)[1]
    raise ValueError("dtype must be torch.float32 or torch.float64")


def build_graph_from_latlon(latlon, device=None, dtype=torch.float32) -> torch.Tensor:
    """The reference's full ``build_graph``: Mercator (host libm) + dense normalised graph."""
    return build_graph(mercator(latlon), device=device, dtype=dtype, k=0)


def knn_graph_from_latlon(latlon, k: int = 8, device=None, dtype=torch.float32) -> torch.Tensor:
    return build_graph(mercator(latlon), device=device, dtype=dtype, k=k)


def synthetic_coordinates(S: int, seed: int = 0, device=None) -> torch.Tensor:
    """``[S, 2]`` (lat, lon) degrees on the device, uniform in the shipped stations' bounding
    box; SplitMix64 stream (station i takes outputs 2i+1, 2i+2 of the stream seeded ``seed``)."""
    lib = _lib.load()
    dev = _device(device)
    out = torch.empty((S, 2), dtype=torch.float64, device=dev)
    _lib.check(
        lib.wg_synthetic_coordinates_f64(
            out.data_ptr(), S, seed & ((1 << 64) - 1), dev.index or 0, torch.cuda.current_stream(dev).cuda_stream
        )
    )
    return out


class CsrGraph:
    """A normalised adjacency in CSR form on the GPU: ``rowptr`` int32 ``[S+1]``, ``colidx`` int32 ``[nnz]``
    (ascending within a row), ``vals`` fp32 ``[nnz]``.  ``GCN_GRU.forward`` accepts it in place of the dense
    ``adj_matrix`` (the scaled-shape path never needs the S x S matrix)."""

    def __init__(self, rowptr: torch.Tensor, colidx: torch.Tensor, vals: torch.Tensor, num_stations: int):
        self.rowptr, self.colidx, self.vals, self.num_stations = rowptr, colidx, vals, int(num_stations)

    @property
    def shape(self):
        return (self.num_stations, self.num_stations)

    @property
    def nnz(self) -> int:
        return int(self.colidx.numel())

    @property
    def device(self):
        return self.vals.device

    def to_dense(self) -> torch.Tensor:
        return torch.sparse_csr_tensor(self.rowptr.long(), self.colidx.long(), self.vals, size=self.shape).to_dense()


def knn_graph_csr(xy, k: int = 8, device=None, dtype=torch.float32) -> CsrGraph:
    """The synthetic large-graph generator straight into CSR (``wg_build_graph_csr_f64``): same pattern, weights
    and normalisation as ``build_graph(xy, k=k)`` — the stored values are bit-identical to that matrix's
    non-zeros — without ever forming the S x S matrix."""
    lib = _lib.load()
    dev = _device(device)
    if k <= 0:
        raise ValueError("k must be positive (the dense reference graph has no CSR form here)")
    if not isinstance(xy, torch.Tensor):
        xy = torch.from_numpy(np.ascontiguousarray(np.asarray(xy, dtype=np.float64)))
    xy = xy.to(device=dev, dtype=torch.float64).contiguous()
    if xy.dim() != 2 or xy.shape[1] != 2:
        raise RuntimeError("xy must be [S, 2]")
    S = xy.shape[0]
    cap = S * (2 * min(k, S - 1) + 1)
    rowptr = torch.empty(S + 1, dtype=torch.int32, device=dev)
    colidx = torch.empty(cap, dtype=torch.int32, device=dev)
    if dtype not in (torch.float32, torch.float64):
        raise ValueError("dtype must be torch.float32 or torch.float64")
    vals = torch.empty(cap, dtype=dtype, device=dev)
    nbytes = lib.wg_build_graph_csr_workspace_bytes(S, k)
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
    v64 = vals.data_ptr() if dtype == torch.float64 else None
    v32 = vals.data_ptr() if dtype == torch.float32 else None
    _lib.check(lib.wg_build_graph_csr_f64(xy.data_ptr(), S, k, rowptr.data_ptr(), colidx.data_ptr(), v64, v32, cap,
                                          ws.data_ptr(), ws.numel(), dev.index or 0,
                                          torch.cuda.current_stream(dev).cuda_stream))
    nnz = int(rowptr[-1].item())   # synchronises: ws may be freed on return
    return CsrGraph(rowptr, colidx[:nnz].contiguous(), vals[:nnz].contiguous(), S)


def knn_graph_csr_from_latlon(latlon, k: int = 8, device=None, dtype=torch.float32) -> CsrGraph:
    return knn_graph_csr(mercator(latlon), k=k, device=device, dtype=dtype)
