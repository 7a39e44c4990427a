"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs ``/root/reference``; the GPU box has no copy):

    python tests/golden/make_golden.py

What it writes (all small, all committed):

``coords.json``          station names + (lat, lon) from ``data/ACISStationCoordinates.csv``
``adj_ref_{7,34}.npy``   fp64 output of the reference ``build_graph`` (step2:16-40) for the
                         first 7 stations / all stations minus "Enchant 2 AGCM" (step1:32)
``adj_ref_rand50.npz``   ``build_graph`` on 50 seeded random coordinates
``wind_gnn_{7,34}.pth``  the shipped checkpoints, byte-for-byte copies (sha256 in manifest)
``fwd_{7,34}.npz``       seeded inputs ``x ~ U[0,1)`` and the reference module's outputs,
                         one ``model(adj, x[b:b+1])`` call per window (the only shape the
                         reference accepts, step6:20), in fp32 and from ``model.double()``
``fwd_34_short.npz``     same, T = 1 and T = 5 windows
``fwd_rand.npz``         a randomly initialised ``GCN_GRU(6, 10, 13, 65, 11)`` on S = 5
                         (exercises F_in != F_hid != 13 and H != 3S), with its parameters
``fwd_wide.npz``         the reference CLASS at the scaled shape's widths, ``GCN_GRU(13, 128, 13, 13*300, 128)`` on a
                         300-station kNN(8) graph (dense matrix, as the reference takes it), B = 2, T = 6, with its
                         parameters: pins the CSR path, which re-associates layer 2 as ``A.(G1.W2)``
                         (``python tests/golden/make_golden.py --only-wide`` writes just this file)
``pivot.npz``            the reference's own ``generate_sequences`` (pivot + split + windows) on a small interleaved
                         long table (``--only-pivot`` writes just this file)
``manifest.json``        sha256 of every file + library versions
"""

import hashlib
import json
import os
import shutil
import sys

import numpy as np
import pandas as pd
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REF, "src"))

import step2_graph_builder as step2  # noqa: E402
from step6_gcn_gru_combined_model import GCN_GRU  # noqa: E402


def ref_build_graph(names, lat, lon):
    df = pd.DataFrame({"Station Name": names, "Latitude": lat, "Longitude": lon})
    # pandas-3 copy-on-write hands step2 a read-only ``.values``; the reference mutates it
    # in place (step2:11-12).  Give it a writable copy without touching its arithmetic.
    orig = step2.__dict__["__convert_wgs2utm"]
    step2.__dict__["__convert_wgs2utm"] = lambda c: orig(np.array(c, dtype=np.float64))
    try:
        return step2.build_graph(df)
    finally:
        step2.__dict__["__convert_wgs2utm"] = orig


def ref_forward(model, adj, x):
    """One reference call per window; returns [B, T, H]."""
    with torch.no_grad():
        return torch.stack([model(adj, x[b : b + 1]) for b in range(x.shape[0])])


def make_windows_golden():
    """``windows.npz``: the reference's own ``__create_sequences`` (step4:7-27) on a small seeded
    table, plus the shuffle indices it drew (recovered by re-seeding NumPy's global RNG)."""
    import step4_sequence_preparer as step4

    create = step4.__dict__["__create_sequences"]
    rng = np.random.default_rng(99)
    Ttot, S, L = 3 * 24 + 5, 4, 24
    table = rng.random((Ttot, S, 15)).astype(np.float32).astype(np.float64)
    np.random.seed(4321)
    x, y = create(table, L)
    np.random.seed(4321)
    idx = np.arange(x.shape[0])
    np.random.shuffle(idx)
    np.savez_compressed(os.path.join(HERE, "windows.npz"), table=table, x=x, y=y, indices=idx, seq_length=L)


def make_pivot_golden():
    """``pivot.npz``: the reference's own ``generate_sequences`` (step4:30-74: pivot, 70/30 split, windows) on a
    small long table whose rows are interleaved by station and whose station names are NOT in alphabetical
    order of first appearance.  Saved: the long table, and the tensors the reference's loaders serve."""
    import step4_sequence_preparer as step4

    rng = np.random.default_rng(2024)
    names = ["Verger AGDM", "Bassano AGCM", "Rosemary IMCIN"]       # first-appearance order != np.unique order
    n_t = 600                                                       # 420 train rows (2 windows), 180 test rows (1)
    rows = []
    for t in range(n_t):
        for st in (names if t % 2 == 0 else names[::-1]):           # file order of the stations varies per hour
            rows.append([st, f"2021-01-01 {t:05d}"] + list(rng.random(13).astype(np.float32).astype(np.float64)))
    long_table = np.array(rows, dtype=object)
    np.random.seed(777)
    train_loader, test_loader, num_attr, num_stations = step4.generate_sequences(long_table, 168, "cpu")
    tr_x, tr_y = train_loader.dataset.tensors
    te_x, te_y = test_loader.dataset.tensors
    assert num_attr == 13 and num_stations == 3
    np.savez_compressed(os.path.join(HERE, "pivot.npz"), station=long_table[:, 0].astype(str),
                        values=long_table[:, 2:].astype(np.float64), train_x=tr_x.numpy(), train_y=tr_y.numpy(),
                        test_x=te_x.numpy(), test_y=te_y.numpy(), seed=777)
    print("pivot.npz: train", tuple(tr_x.shape), tuple(tr_y.shape), "test", tuple(te_x.shape), tuple(te_y.shape))


def make_wide_golden():
    """``fwd_wide.npz``: the unmodified reference class at wide dims (hidden 128, GRU hidden 128) on a
    300-station kNN graph.  The graph comes from the oracle's kNN generator (the reference has none);
    the forward is the reference's own ``model(adj, x[b:b+1])``, in fp32 and in fp64."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import knn_graph_f64, synthetic_coordinates

    S, Fh, H, T, B = 300, 128, 128, 6, 2
    adj64 = knn_graph_f64(synthetic_coordinates(S, seed=3), k=8)
    adj = torch.tensor(adj64).float()
    torch.manual_seed(11)
    model = GCN_GRU(13, Fh, 13, 13 * S, H).eval()
    with torch.no_grad():
        model.conv1.weight.mul_(0.05)       # unscaled randn saturates the GRU at this width
        model.conv2.weight.mul_(0.05)
        model.conv1.bias.uniform_(-0.1, 0.1)
        model.conv2.bias.uniform_(-0.1, 0.1)
    x = torch.rand(B, T, S, 13, generator=torch.Generator().manual_seed(12))
    y32 = ref_forward(model, adj, x)
    m64 = GCN_GRU(13, Fh, 13, 13 * S, H).double()
    m64.load_state_dict(model.state_dict())
    y64 = ref_forward(m64, adj.double(), x.double())
    np.savez_compressed(
        os.path.join(HERE, "fwd_wide.npz"), adj=adj.numpy(), x=x.numpy(), y_ref_f32=y32.numpy(), y_ref_f64=y64.numpy(),
        **{k.replace(".", "__"): v.numpy() for k, v in model.state_dict().items()})
    print("fwd_wide.npz: out", tuple(y32.shape), "fp32 vs fp64 normalised max",
          float(np.abs(y32.numpy() - y64.numpy()).max() / np.abs(y64.numpy()).max()))


def update_manifest_entry(fn):
    path = os.path.join(HERE, "manifest.json")
    with open(path) as f:
        manifest = json.load(f)
    with open(os.path.join(HERE, fn), "rb") as f:
        manifest["sha256"][fn] = hashlib.sha256(f.read()).hexdigest()
    with open(path, "w") as f:
        json.dump(manifest, f, indent=1)


def main():
    if "--only-pivot" in sys.argv:
        make_pivot_golden()
        update_manifest_entry("pivot.npz")
        return
    if "--only-wide" in sys.argv:
        torch.set_num_threads(1)
        make_wide_golden()
        update_manifest_entry("fwd_wide.npz")
        return
    make_windows_golden()
    make_pivot_golden()
    make_wide_golden()
    torch.manual_seed(0)
    np.random.seed(0)
    torch.set_num_threads(1)  # deterministic summation order in MKL

    coords = pd.read_csv(os.path.join(REF, "data", "ACISStationCoordinates.csv"))
    names = coords["Station Name"].tolist()
    lat = coords["Latitude"].tolist()
    lon = coords["Longitude"].tolist()
    with open(os.path.join(HERE, "coords.json"), "w") as f:
        json.dump({"names": names, "lat": lat, "lon": lon}, f, indent=1)

    keep34 = [i for i, n in enumerate(names) if n != "Enchant 2 AGCM"]
    sets = {7: list(range(7)), 34: keep34}
    adj = {}
    for S, idx in sets.items():
        a = ref_build_graph([names[i] for i in idx], [lat[i] for i in idx], [lon[i] for i in idx])
        assert a.shape == (S, S) and a.dtype == np.float64
        np.save(os.path.join(HERE, f"adj_ref_{S}.npy"), a)
        adj[S] = torch.tensor(a).float()  # main.py:26

    rng = np.random.default_rng(1234)
    rlat = rng.uniform(49.0, 53.0, 50)
    rlon = rng.uniform(-115.0, -109.0, 50)
    a = ref_build_graph([f"s{i}" for i in range(50)], rlat.tolist(), rlon.tolist())
    np.savez(os.path.join(HERE, "adj_ref_rand50.npz"), lat=rlat, lon=rlon, adj=a)

    for S in (7, 34):
        shutil.copyfile(os.path.join(REF, f"wind_gnn_{S}.pth"), os.path.join(HERE, f"wind_gnn_{S}.pth"))
        sd = torch.load(os.path.join(HERE, f"wind_gnn_{S}.pth"), map_location="cpu", weights_only=True)
        model = GCN_GRU(13, 13, 13, 13 * S, 3 * S)  # main.py:39-42
        model.load_state_dict(sd, strict=True)  # main.py:99
        model.eval()
        B = 3 if S == 7 else 2
        g = torch.Generator().manual_seed(100 + S)
        x = torch.rand(B, 168, S, 13, generator=g)
        y32 = ref_forward(model, adj[S], x)
        m64 = GCN_GRU(13, 13, 13, 13 * S, 3 * S).double()
        m64.load_state_dict(sd, strict=True)
        y64 = ref_forward(m64, adj[S].double(), x.double())
        np.savez(
            os.path.join(HERE, f"fwd_{S}.npz"),
            x=x.numpy(), y_ref_f32=y32.numpy(), y_ref_f64=y64.numpy(),
        )
        if S == 34:
            out = {}
            for T in (1, 5):
                xs = torch.rand(1, T, S, 13, generator=g)
                out[f"x_T{T}"] = xs.numpy()
                out[f"y_T{T}"] = ref_forward(model, adj[S], xs).numpy()
            np.savez(os.path.join(HERE, "fwd_34_short.npz"), **out)

    # random small model with non-default dims (the ctor allows them; only output_dim
    # must be 13 because of the hard-coded 13 at step6:16)
    torch.manual_seed(7)
    S, Fin, Fh, H, T, B = 5, 6, 10, 11, 9, 4
    model = GCN_GRU(Fin, Fh, 13, 13 * S, H).eval()
    with torch.no_grad():
        for p in model.conv1.parameters():
            p.mul_(0.4)
        for p in model.conv2.parameters():
            p.mul_(0.4)
        model.conv1.bias.uniform_(-0.3, 0.3)
        model.conv2.bias.uniform_(-0.3, 0.3)
    a = torch.rand(S, S) * 0.4
    x = torch.rand(B, T, S, Fin)
    y = ref_forward(model, a, x)
    np.savez(
        os.path.join(HERE, "fwd_rand.npz"),
        adj=a.numpy(), x=x.numpy(), y_ref_f32=y.numpy(),
        **{k.replace(".", "__"): v.numpy() for k, v in model.state_dict().items()},
    )

    manifest = {
        "versions": {
            "torch": torch.__version__, "numpy": np.__version__,
            "scipy": __import__("scipy").__version__, "pandas": pd.__version__,
        },
        "reference": "NagsTheProgrammer/WindGNN mounted at /root/reference (unmodified)",
        "sha256": {},
    }
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith((".npy", ".npz", ".pth")) or fn == "coords.json":
            with open(os.path.join(HERE, fn), "rb") as f:
                manifest["sha256"][fn] = hashlib.sha256(f.read()).hexdigest()
    for S in (7, 34):
        with open(os.path.join(REF, f"wind_gnn_{S}.pth"), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == manifest["sha256"][f"wind_gnn_{S}.pth"]
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    main()
