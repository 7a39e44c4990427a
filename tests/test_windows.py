"""step4 window extraction and the de-normalised last-step slice: oracle pinned to the reference's
own ``__create_sequences`` (CPU), CUDA path bit-exact against the oracle (GPU)."""

import numpy as np
import pytest
import torch

import windgnn_b200
from conftest import golden
from oracle import create_sequences, denorm_last_step, pivot_long_table
from windgnn_b200 import _lib

DEV = "cuda:0"


def test_oracle_matches_reference_create_sequences():
    g = golden("windows.npz")
    x, y = create_sequences(g["table"], int(g["seq_length"]), g["indices"])
    assert x.shape == g["x"].shape == (3, 24, 4, 13) and y.shape == g["y"].shape == (3, 24, 12)
    assert np.array_equal(x, g["x"]) and np.array_equal(y, g["y"])
    # label layout: y[n, l, k*S + s] = wind speed (column 13) k+1 rows later
    n = int(g["indices"][0])
    assert y[0, 5, 2 * 4 + 1] == g["table"][n * 24 + 5 + 3, 1, 13]


def test_num_windows_host_logic():
    assert windgnn_b200.num_windows(168 * 5 + 3) == 5
    assert windgnn_b200.num_windows(168 * 5 + 2) == 4     # labels of the 5th window would be ragged
    assert windgnn_b200.num_windows(100) == 0
    assert windgnn_b200.num_windows(77, 24, 3) == 3
    lib = _lib.load()
    assert lib.wg_make_windows_f32(None, None, None, None, 10, 2, 13, 24, 11, 3, 1, 0, None) == _lib.WG_ERR_BAD_ARG
    assert lib.wg_denorm_last_step_f32(None, None, 0, 168, 102, 0.0, 1.0, 0, None) == _lib.WG_OK


@pytest.mark.gpu
def test_windows_bit_exact_vs_reference_golden():
    g = golden("windows.npz")
    L = int(g["seq_length"])
    table = torch.from_numpy(g["table"][:, :, 2:15].astype(np.float32)).to(DEV)
    perm = torch.from_numpy(g["indices"].astype(np.int64)).to(DEV)
    x, y = windgnn_b200.create_sequences(table, L, perm=perm)
    assert torch.equal(x.cpu(), torch.from_numpy(g["x"].astype(np.float32)))
    assert torch.equal(y.cpu(), torch.from_numpy(g["y"].astype(np.float32)))
    # chronological order: x is a zero-copy view of the table
    x0, y0 = windgnn_b200.create_sequences(table, L)
    assert x0.data_ptr() == table.data_ptr() and x0.shape == (3, L, 4, 13)
    xo, yo = create_sequences(g["table"], L)
    assert torch.equal(x0.cpu(), torch.from_numpy(xo.astype(np.float32)))
    assert torch.equal(y0.cpu(), torch.from_numpy(yo.astype(np.float32)))


@pytest.mark.gpu
def test_windows_full_size_properties():
    """The real dataset's shape: 3 years of hourly rows, 34 stations (26,304 x 34 x 13)."""
    Ttot, S = 26304, 34
    gen = torch.Generator(device=DEV).manual_seed(5)
    table = torch.rand((Ttot, S, 13), generator=gen, device=DEV)
    N = windgnn_b200.num_windows(Ttot)
    assert N == Ttot // 168                                   # 156 full windows, 96 spare rows
    perm = torch.randperm(N, device=DEV, generator=gen)
    x, y = windgnn_b200.create_sequences(table, perm=perm)
    assert x.shape == (N, 168, S, 13) and y.shape == (N, 168, 3 * S)
    flat = table.view(-1, S, 13)
    n = 17
    w = int(perm[n])
    assert torch.equal(x[n], flat[w * 168:(w + 1) * 168])
    for k in range(3):
        assert torch.equal(y[n, :, k * S:(k + 1) * S], flat[w * 168 + 1 + k:(w + 1) * 168 + 1 + k, :, 11])
    # a permutation only reorders windows
    x2, y2 = windgnn_b200.create_sequences(table)
    assert torch.equal(x2[perm], x) and torch.equal(y2[perm], y)
    with pytest.raises(RuntimeError, match="complete windows"):
        windgnn_b200.create_sequences(table[: 168 * 3 + 2], perm=torch.arange(3, device=DEV))
    # a SHORT perm naming a window that does not exist (or a negative one) must not read out of bounds
    with pytest.raises(RuntimeError, match="complete windows"):
        windgnn_b200.create_sequences(table, perm=torch.tensor([N], device=DEV))
    with pytest.raises(RuntimeError, match="complete windows"):
        windgnn_b200.create_sequences(table, perm=torch.tensor([0, -1], device=DEV))


@pytest.mark.gpu
def test_denormalise_last_step_bit_exact():
    rng = np.random.default_rng(3)
    out = (rng.random((37, 9, 21), dtype=np.float32) * 2 - 1).astype(np.float32)
    wmin, wmax = 0.0, 83.6                                     # km/h range of the kind step1:56-58 returns
    pred = windgnn_b200.denormalise_last_step(torch.from_numpy(out).to(DEV), wmin, wmax).cpu().numpy()
    assert np.array_equal(pred, denorm_last_step(out, wmin, wmax))
    wmin, wmax = 1.25, 97.30000000000001
    pred = windgnn_b200.denormalise_last_step(torch.from_numpy(out).to(DEV), wmin, wmax).cpu().numpy()
    assert np.array_equal(pred, denorm_last_step(out, wmin, wmax))
    one = windgnn_b200.denormalise_last_step(torch.from_numpy(out[0]).to(DEV), wmin, wmax)   # [T, H] like the reference
    assert one.shape == (1, 21)


def _pivot_golden_long_table():
    g = golden("pivot.npz")
    n = len(g["station"])
    lt = np.concatenate([g["station"].astype(object)[:, None], np.zeros((n, 1), dtype=object),
                         g["values"].astype(object)], axis=1)
    return g, lt


def test_oracle_pivot_and_windows_match_reference_generate_sequences():
    """pivot.npz was produced by the reference's own generate_sequences (step4:30-74) on an interleaved long
    table: oracle pivot (step4:36-47) + 70/30 split (:50) + oracle windows (:7-27) reproduce its loaders."""
    g, lt = _pivot_golden_long_table()
    tab, stations = pivot_long_table(lt)
    assert tab.shape == (600, 3, 15) and list(stations) == sorted(set(g["station"]))
    n_test = int(np.ceil(0.3 * tab.shape[0]))
    n_train = tab.shape[0] - n_test
    np.random.seed(int(g["seed"]))                    # the reference shuffles train windows first, then test
    for part, data in (("train", tab[:n_train]), ("test", tab[n_train:])):
        idx = np.arange(len(data) // 168)
        np.random.shuffle(idx)
        x, y = create_sequences(data, 168, idx)
        assert np.array_equal(x.astype(np.float32), g[f"{part}_x"])
        assert np.array_equal(y.astype(np.float32), g[f"{part}_y"])


@pytest.mark.gpu
def test_device_pivot_bit_exact_vs_reference_golden():
    g, lt = _pivot_golden_long_table()
    values = torch.from_numpy(g["values"].astype(np.float32)).to(DEV)
    table, stations = windgnn_b200.pivot_long_table(g["station"], values)
    ref, ref_st = pivot_long_table(lt)
    assert list(stations) == list(ref_st)
    assert torch.equal(table.cpu(), torch.from_numpy(ref[:, :, 2:].astype(np.float32)))
    # and the whole step4 chain on the device against the reference's loaders: split, windows, labels
    n_test = int(np.ceil(0.3 * table.shape[0]))
    n_train = table.shape[0] - n_test
    np.random.seed(int(g["seed"]))
    for part, data in (("train", table[:n_train]), ("test", table[n_train:])):
        idx = np.arange(data.shape[0] // 168)
        np.random.shuffle(idx)
        x, y = windgnn_b200.create_sequences(data.contiguous(), 168, perm=torch.from_numpy(idx).to(DEV))
        assert torch.equal(x.cpu(), torch.from_numpy(g[f"{part}_x"]))
        assert torch.equal(y.cpu(), torch.from_numpy(g[f"{part}_y"]))
    with pytest.raises(RuntimeError, match="split evenly|unequal"):
        windgnn_b200.pivot_long_table(g["station"][:-1], values[:-1])
