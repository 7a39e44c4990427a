"""Small shapes through every new kernel (for compute-sanitizer memcheck / racecheck)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import windgnn_b200
from windgnn_b200 import train
dev = torch.device("cuda:0")
torch.manual_seed(0)
for S, B, T in ((7, 5, 3), (34, 3, 4), (34, 37, 2)):
    sd = torch.load(os.path.join(ROOT, f"tests/golden/wind_gnn_{S}.pth"), map_location="cpu", weights_only=True)
    adj = torch.from_numpy(np.load(os.path.join(ROOT, f"tests/golden/adj_ref_{S}.npy")).astype(np.float32)).to(dev)
    m = windgnn_b200.GCN_GRU(13, 13, 13, 13 * S, 3 * S); m.load_state_dict(sd); m = m.to(dev)
    x = torch.rand((B, T, S, 13), device=dev); y = torch.rand((B, T, 3 * S), device=dev)
    with torch.no_grad():
        out = m(adj, x)                                   # small-batch recurrence
        xh = x.cpu().pin_memory(); m.chunk = 2; oh = m.forward_host(adj, xh); m.chunk = 0
    tr = train.Trainer(m, adj)
    l1 = tr.step(x, y); l2 = tr.step(x, y)
    torch.cuda.synchronize()
    print(S, B, T, float(out.abs().max()), float(l1), float(l2), bool(torch.equal(oh.to(dev), out)))
# throughput recurrence + 16-sequence BPTT kernel (batch past both thresholds), short T
S = 7
sd = torch.load(os.path.join(ROOT, "tests/golden/wind_gnn_7.pth"), map_location="cpu", weights_only=True)
adj = torch.from_numpy(np.load(os.path.join(ROOT, "tests/golden/adj_ref_7.npy")).astype(np.float32)).to(dev)
m = windgnn_b200.GCN_GRU(13, 13, 13, 91, 21); m.load_state_dict(sd); m = m.to(dev)
x = torch.rand((1201, 2, S, 13), device=dev); y = torch.rand((1201, 2, 21), device=dev)
tr = train.Trainer(m, adj); print("big", float(tr.step(x, y)))
# sparse row-resident path
from oracle import knn_graph_f64, synthetic_coordinates
S = 96
adjs = torch.from_numpy(knn_graph_f64(synthetic_coordinates(S, seed=1), 3).astype(np.float32)).to(dev)
ms = windgnn_b200.GCN_GRU(13, 40, 13, 13 * S, 30).to(dev); ms.DENSE_MAX_STATIONS = 64
with torch.no_grad():
    print("sparse", float(ms(adjs, torch.rand((3, 4, S, 13), device=dev)).abs().max()))
torch.cuda.synchronize()
print("done")
