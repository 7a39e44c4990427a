"""Tensor path on a B200: error against the FP32 path / the goldens, stage times (run through gpurun)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import windgnn_b200
from windgnn_b200 import _lib
from oracle import normalised_max_error

dev = torch.device("cuda:0")
lib = _lib.load()
GOLD = os.path.join(ROOT, "tests", "golden")

def model_for(S):
    sd = torch.load(os.path.join(GOLD, f"wind_gnn_{S}.pth"), map_location="cpu", weights_only=True)
    m = windgnn_b200.GCN_GRU(13, 13, 13, 13 * S, 3 * S)
    m.load_state_dict(sd, strict=True)
    adj = torch.from_numpy(np.load(os.path.join(GOLD, f"adj_ref_{S}.npy")).astype(np.float32)).to(dev)
    return m.to(dev).eval(), adj

res = {}
for S in (7, 34):
    m, adj = model_for(S)
    g = np.load(os.path.join(GOLD, f"fwd_{S}.npz"))
    x = torch.from_numpy(g["x"]).to(dev)
    m.precision = "tf32x3"
    with torch.no_grad():
        y = m(adj, x)
    torch.cuda.synchronize()
    y = y.cpu().numpy()
    res[f"golden_S{S}"] = {"finite": bool(np.isfinite(y).all()), "err_f32": normalised_max_error(y, g["y_ref_f32"]),
                           "err_f64": normalised_max_error(y, g["y_ref_f64"])}
    print(f"S={S}", res[f"golden_S{S}"], flush=True)

S, B, T = 34, 4096, 168
m, adj = model_for(S)
x = torch.rand((B, T, S, 13), generator=torch.Generator(device=dev).manual_seed(7), device=dev)
with torch.no_grad():
    y32 = m(adj, x)
    m.precision = "tf32x3"
    yt = m(adj, x)
    yt_small = m(adj, x[1000:1300])
torch.cuda.synchronize()
res["full"] = {"err_vs_fp32_path": float((yt - y32).abs().max() / y32.abs().max()),
               "shard_invariant": bool(torch.equal(yt_small, yt[1000:1300])), "finite": bool(torch.isfinite(yt).all())}
print("full", res["full"], flush=True)

H = 3 * S
dims = (T, S, 13, 13, 13, H)
for flags in (0, 1):
    nbytes = lib.wg_gcn_gru_workspace_bytes(B, *dims, B, flags)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty((B, T, H), device=dev)
    p = [t.detach().contiguous() for t in (m.conv1.weight, m.conv1.bias, m.conv2.weight, m.conv2.bias,
         m.gru.weight_ih_l0, m.gru.weight_hh_l0, m.gru.bias_ih_l0, m.gru.bias_hh_l0)]
    st = torch.cuda.current_stream(dev).cuda_stream
    calls = {
        "pack": lambda: lib.wg_stage_pack_f32(*(t.data_ptr() for t in p[4:]), *dims, B, flags, ws.data_ptr(), nbytes, 0, st),
        "gcn": lambda: lib.wg_stage_gcn_f32(adj.data_ptr(), x.data_ptr(), *(t.data_ptr() for t in p[:4]), B, *dims, B, flags, ws.data_ptr(), nbytes, 0, st),
        "inproj": lambda: lib.wg_stage_inproj_f32(B, *dims, B, flags, ws.data_ptr(), nbytes, 0, st),
        "recur": lambda: lib.wg_stage_recur_f32(out.data_ptr(), B, *dims, B, flags, ws.data_ptr(), nbytes, 0, st),
    }
    t_ms = {}
    for name, fn in calls.items():
        for _ in range(3):
            _lib.check(fn())
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            _lib.check(fn())
        b.record()
        torch.cuda.synchronize()
        t_ms[name] = a.elapsed_time(b) / 10
    res[f"stage_ms_flags{flags}"] = t_ms
    print("flags", flags, {k: round(v, 4) for k, v in t_ms.items()}, "sum", round(sum(t_ms.values()), 3), flush=True)
    del ws, out
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tc_check.json"), "w"), indent=1)
