"""Debug: per-phase cycle counts of gru_recur_unit_kernel (needs a -DWG_RC_TRACE build:
WG_LIB_SUFFIX=trace WG_NVCC_FLAGS=-DWG_RC_TRACE python -m windgnn_b200.build; run with WINDGNN_B200_LIB set)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, windgnn_b200
from windgnn_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sd, latlon = bench.load_workload()
dev = torch.device("cuda:0")
model = windgnn_b200.GCN_GRU(13, 13, 13, 442, 102); model.load_state_dict(sd); model = model.to(dev).eval()
adj = windgnn_b200.build_graph_from_latlon(latlon, device=dev)
x = torch.rand((B, 168, 34, 13), device=dev)
with torch.no_grad():
    for _ in range(3): y = model(adj, x)
torch.cuda.synchronize()
raw = ctypes.CDLL(_lib.LIB_PATH)
n = 8 * 256 * 5
buf = (ctypes.c_longlong * n)()
rc = raw.wg_debug_read_unit_trace(buf, n)
a = np.array(buf[:], dtype=np.int64).reshape(8, 256, 5)
names = ["gemm", "cpwait+gi", "gates", "barrier"]
for w in range(8):
    d = np.diff(a[w, 20:160, :], axis=1)
    step = a[w, 21:161, 0] - a[w, 20:160, 0]
    wait = a[w, 21:161, 0] - a[w, 20:160, 4]   # hand-over wait (+ loop overhead) before the product
    print(f"B={B} warp {w}: step {step.mean():.0f} cyc | wait {wait.mean():.0f} | " + " ".join(f"{n_}={v:.0f}" for n_, v in zip(names, d.mean(axis=0))))
print("gemm start offsets vs warp 0:", [(a[w, 20:160, 0] - a[0, 20:160, 0]).mean().round() for w in range(8)])
