"""A/B of gru_recur_unit_kernel with and without the FMA-pipe hand-over between the two groups of a
scheduler (WG_RU_HANDOVER=0 / 1): outputs must be bit-identical; the recurrence stage is timed alone.

    python scripts/handover_ab.py
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


results, ok = {}, True
for S, B, T in [(34, 4096, 168), (34, 4736, 168), (34, 2000, 168), (34, 1301, 7), (7, 4096, 168), (7, 1500, 24)]:
    m, adj = model_for(S)
    x = torch.rand((B, T, S, 13), generator=torch.Generator(device=dev).manual_seed(S * 1000 + B), device=dev)
    ys = {}
    for h in ("0", "1"):
        os.environ["WG_RU_HANDOVER"] = h
        with torch.no_grad():
            ys[h] = m(adj, x)
        torch.cuda.synchronize()
    same = bool(torch.equal(ys["0"], ys["1"]))
    ok = ok and same
    rec = {"bit_identical": same}
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
    _lib.check(lib.wg_stage_gcn_f32(adj.data_ptr(), xs.data_ptr(), *(t.data_ptr() for t in p[:4]), Bc, *dims, Bc, 0,
                                    ws.data_ptr(), nbytes, 0, st))
    _lib.check(lib.wg_stage_inproj_f32(Bc, *dims, Bc, 0, ws.data_ptr(), nbytes, 0, st))
    for h in ("0", "1", "0", "1"):
        os.environ["WG_RU_HANDOVER"] = h
        rec.setdefault(f"recur_ms_handover{h}", []).append(round(time_stage(lambda: lib.wg_stage_recur_f32(
            out.data_ptr(), Bc, *dims, Bc, 0, ws.data_ptr(), nbytes, 0, st)), 4))
    results[f"S{S}_B{B}_T{T}"] = rec
    print(f"S={S} B={B} T={T}: {json.dumps(rec)}", flush=True)
    del x, ys, ws, out
os.environ.pop("WG_RU_HANDOVER", None)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(results, open(os.path.join(ROOT, "gpurun_out", "handover_ab.json"), "w"), indent=1)
print("A/B", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
