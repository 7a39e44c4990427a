"""A/B of the first- and second-generation FP32 kernels on a B200 (run through gpurun).

For each shape: full forward with WG_FORCE_LEGACY=1 (gcn_kernel / gru_recur_kernel) and with the
default dispatch (gcn_rows_kernel / gru_recur_unit_kernel); the outputs must be bit-identical.
Then the stage entry points of the C-ABI are timed alone for both generations.

    python scripts/kernel_ab.py [--quick]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import windgnn_b200  # noqa: E402
from windgnn_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
GOLD = os.path.join(ROOT, "tests", "golden")


def model_for(S):
    sd = torch.load(os.path.join(GOLD, f"wind_gnn_{S}.pth"), map_location="cpu", weights_only=True)
    m = windgnn_b200.GCN_GRU(13, 13, 13, 13 * S, 3 * S)
    m.load_state_dict(sd, strict=True)
    adj = torch.from_numpy(np.load(os.path.join(GOLD, f"adj_ref_{S}.npy")).astype(np.float32)).to(dev)
    return m.to(dev).eval(), adj


def fwd(m, adj, x, legacy):
    os.environ["WG_FORCE_LEGACY"] = "1" if legacy else "0"
    with torch.no_grad():
        y = m(adj, x)
    torch.cuda.synchronize()
    return y


def time_stage(fn, reps=10):
    for _ in range(3):
        _lib.check(fn())
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        _lib.check(fn())
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


results = {}
ok = True
shapes = [(34, 4096, 168), (34, 4736, 168), (34, 2000, 168), (34, 1301, 7), (7, 4096, 168), (7, 1500, 24)]
if "--quick" in sys.argv:
    shapes = shapes[:1]
for S, B, T in shapes:
    m, adj = model_for(S)
    x = torch.rand((B, T, S, 13), generator=torch.Generator(device=dev).manual_seed(S * 1000 + B), device=dev)
    y_old = fwd(m, adj, x, True)
    y_new = fwd(m, adj, x, False)
    same = bool(torch.equal(y_old, y_new))
    diff = float((y_old - y_new).abs().max())
    finite = bool(torch.isfinite(y_new).all())
    ok = ok and same and finite
    rec = {"bit_identical": same, "max_abs_diff": diff, "finite": finite}
    # stage timings
    H = 3 * S
    dims = (T, S, 13, 13, 13, H)
    Bc = min(B, 148 * 32)
    nbytes = lib.wg_gcn_gru_workspace_bytes(Bc, *dims, Bc, 0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty((Bc, T, H), device=dev)
    p = [t.detach().contiguous() for t in (
        m.conv1.weight, m.conv1.bias, m.conv2.weight, m.conv2.bias,
        m.gru.weight_ih_l0, m.gru.weight_hh_l0, m.gru.bias_ih_l0, m.gru.bias_hh_l0)]
    st = torch.cuda.current_stream(dev).cuda_stream
    xs = x[:Bc]
    _lib.check(lib.wg_stage_pack_f32(*(t.data_ptr() for t in p[4:]), *dims, Bc, 0, ws.data_ptr(), nbytes, 0, st))
    for legacy in (True, False):
        os.environ["WG_FORCE_LEGACY"] = "1" if legacy else "0"
        tag = "legacy" if legacy else "new"
        rec[f"gcn_ms_{tag}"] = time_stage(lambda: lib.wg_stage_gcn_f32(
            adj.data_ptr(), xs.data_ptr(), *(t.data_ptr() for t in p[:4]), Bc, *dims, Bc, 0, ws.data_ptr(), nbytes, 0, st))
        _lib.check(lib.wg_stage_inproj_f32(Bc, *dims, Bc, 0, ws.data_ptr(), nbytes, 0, st))
        rec[f"recur_ms_{tag}"] = time_stage(lambda: lib.wg_stage_recur_f32(
            out.data_ptr(), Bc, *dims, Bc, 0, ws.data_ptr(), nbytes, 0, st))
    results[f"S{S}_B{B}_T{T}"] = rec
    print(f"S={S} B={B} T={T}: {json.dumps(rec)}", flush=True)
    del x, y_old, y_new, ws, out
os.environ["WG_FORCE_LEGACY"] = "0"
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(results, open(os.path.join(ROOT, "gpurun_out", "kernel_ab.json"), "w"), indent=1)
print("A/B", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
