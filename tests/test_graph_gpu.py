"""GPU graph build: bit-exact against the reference's ``build_graph`` output (committed golden
fixtures) and against the oracle for the synthetic kNN generator."""

import numpy as np
import pytest
import torch

import windgnn_b200
from conftest import golden, station_latlon
from oracle import dense_graph_f64, knn_graph_f64
from oracle import synthetic_coordinates as oracle_coords

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("S", [7, 34])
def test_dense_graph_bit_exact_vs_reference(S):
    ref = golden(f"adj_ref_{S}.npy")
    ll = station_latlon(S)
    a64 = windgnn_b200.build_graph_from_latlon(ll, device=DEV, dtype=torch.float64).cpu().numpy()
    a32 = windgnn_b200.build_graph_from_latlon(ll, device=DEV, dtype=torch.float32).cpu().numpy()
    assert np.array_equal(a64, ref)                                   # fp64 bit pattern (step2:38)
    assert np.array_equal(a32, ref.astype(np.float32))                # main.py:26 cast


def test_dense_graph_random_coordinates_vs_reference():
    g = golden("adj_ref_rand50.npz")
    ll = np.stack([g["lat"], g["lon"]], axis=1)
    a = windgnn_b200.build_graph_from_latlon(ll, device=DEV, dtype=torch.float64).cpu().numpy()
    assert np.array_equal(a, g["adj"])


@pytest.mark.parametrize("S", [1, 2, 96, 700])
def test_dense_graph_vs_oracle(S):
    ll = oracle_coords(S, seed=11)
    a = windgnn_b200.build_graph_from_latlon(ll, device=DEV, dtype=torch.float64).cpu().numpy()
    assert np.array_equal(a, dense_graph_f64(ll))


def test_synthetic_coordinates_bit_exact():
    for S, seed in ((1, 0), (35, 7), (4096, 0)):
        got = windgnn_b200.synthetic_coordinates(S, seed=seed, device=DEV).cpu().numpy()
        assert np.array_equal(got, oracle_coords(S, seed=seed))


@pytest.mark.parametrize("S,k", [(9, 8), (96, 8), (512, 8), (300, 3), (64, 1)])
def test_knn_graph_bit_exact_vs_oracle(S, k):
    ll = oracle_coords(S, seed=3)
    a = windgnn_b200.knn_graph_from_latlon(ll, k=k, device=DEV, dtype=torch.float64).cpu().numpy()
    ref = knn_graph_f64(ll, k)
    assert np.array_equal(a != 0, ref != 0)                           # same edge set
    assert np.array_equal(a, ref)


def test_knn_graph_4096_properties():
    """BASELINE config 4's graph: S = 4096, k = 8 — checked through properties (the oracle's dense
    O(S^2) fp64 path is used too; it takes a couple of seconds)."""
    S, k = 4096, 8
    ll = windgnn_b200.synthetic_coordinates(S, seed=0, device=DEV).cpu().numpy()
    a = windgnn_b200.knn_graph_from_latlon(ll, k=k, device=DEV, dtype=torch.float64)
    nz = a != 0
    assert torch.equal(nz, nz.T)                                      # symmetric pattern
    deg = nz.sum(1)
    assert int(deg.min()) >= k + 1 and int(deg.max()) < 64
    assert torch.all(torch.diagonal(a) > 0)
    assert np.array_equal(a.cpu().numpy(), knn_graph_f64(ll, k))


@pytest.mark.parametrize("S,k", [(300, 8), (4096, 8), (37, 40)])
def test_knn_graph_straight_into_csr_is_bit_identical_to_the_dense_build(S, k):
    """wg_build_graph_csr_f64 never forms the S x S matrix; its values must be the non-zeros of the dense
    build bit for bit (fp64 and the fp32 cast), columns ascending, pattern symmetric with self loops."""
    ll = windgnn_b200.synthetic_coordinates(S, seed=S, device=DEV).cpu().numpy()
    xy = windgnn_b200.mercator(ll)
    for dtype in (torch.float64, torch.float32):
        dense = windgnn_b200.build_graph(xy, device=DEV, dtype=dtype, k=k)
        csr = windgnn_b200.knn_graph_csr(xy, k=k, device=DEV, dtype=dtype)
        ref = dense.to_sparse_csr()
        assert torch.equal(csr.rowptr.long(), ref.crow_indices())
        assert torch.equal(csr.colidx.long(), ref.col_indices())
        assert torch.equal(csr.vals, ref.values())
        assert csr.nnz <= S * (2 * min(k, S - 1) + 1) and csr.shape == (S, S)
    assert torch.equal(csr.to_dense(), dense)


def test_forward_accepts_the_csr_graph_object():
    from oracle import gcn_gru_forward, normalised_max_error

    S, Fh, H, B, T = 300, 64, 48, 3, 5
    ll = windgnn_b200.synthetic_coordinates(S, seed=9, device=DEV).cpu().numpy()
    csr = windgnn_b200.knn_graph_csr_from_latlon(ll, k=8, device=DEV)
    dense = windgnn_b200.knn_graph_from_latlon(ll, k=8, device=DEV)
    torch.manual_seed(5)
    m = windgnn_b200.GCN_GRU(13, Fh, 13, 13 * S, H)
    with torch.no_grad():
        m.conv1.weight.mul_(0.05)
        m.conv2.weight.mul_(0.05)
    sd = {k_: v.detach().clone() for k_, v in m.state_dict().items()}
    m = m.to(DEV).eval()
    x = torch.rand((B, T, S, 13), device=DEV)
    with torch.no_grad():
        y_csr = m(csr, x)
        y_dense = m(dense, x)
        m.precision = "tensor"
        y_tc = m(csr, x)
    assert torch.equal(y_csr, y_dense)
    ref = gcn_gru_forward(dense.cpu().numpy(), x.cpu().numpy(), sd, dtype=np.float32)
    assert normalised_max_error(y_csr.cpu().numpy(), ref) <= 1e-5
    assert normalised_max_error(y_tc.cpu().numpy(), ref) <= 1e-5       # tensor path on the CSR graph path
