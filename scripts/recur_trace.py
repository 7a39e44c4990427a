"""Debug: per-phase cycle counts of the recurrence kernel (needs a -DWG_RC_TRACE build)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, windgnn_b200
from windgnn_b200 import _lib
sd, latlon = bench.load_workload()
dev = torch.device("cuda:0")
model = windgnn_b200.GCN_GRU(13, 13, 13, 442, 102); model.load_state_dict(sd); model = model.to(dev).eval()
adj = windgnn_b200.build_graph_from_latlon(latlon, device=dev)
x = torch.rand((4096, 168, 34, 13), device=dev)
with torch.no_grad():
    for _ in range(3): y = model(adj, x)
torch.cuda.synchronize()
raw = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_longlong * (2 * 8 * 256))()
rc = raw.wg_debug_read_trace(buf, 2 * 8 * 256)
a = np.array(buf[:], dtype=np.int64).reshape(2, 256, 8)
names = ["token_wait", "gemm", "cpwait+merge", "bar1", "gate", "bar2", "prefetch", "->next"]
for g in range(2):
    d = np.diff(a[g, 20:160, :], axis=1)          # phases within a step
    nxt = a[g, 21:161, 0] - a[g, 20:160, 7]
    step = a[g, 21:161, 0] - a[g, 20:160, 0]
    print(f"group {g}: step {step.mean():.0f} cyc | " + " ".join(f"{n}={v:.0f}" for n, v in zip(names[:7], d.mean(axis=0))) + f" loop_back={nxt.mean():.0f}")
print("offset between groups' GEMM starts:", (a[1, 20:160, 1] - a[0, 20:160, 1]).mean())
