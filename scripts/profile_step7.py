"""A few forwards of the 7-station checkpoint (configs[0] shapes, 4096 windows) — the command profiled under ncu."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import windgnn_b200
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda:0")
G = os.path.join(ROOT, "tests", "golden")
sd = torch.load(os.path.join(G, "wind_gnn_7.pth"), map_location="cpu", weights_only=True)
model = windgnn_b200.GCN_GRU(13, 13, 13, 91, 21)
model.load_state_dict(sd)
model = model.to(dev).eval()
adj = torch.from_numpy(np.load(os.path.join(G, "adj_ref_7.npy")).astype(np.float32)).to(dev)
x = torch.rand((B, 168, 7, 13), device=dev, generator=torch.Generator(device=dev).manual_seed(0))
with torch.no_grad():
    for _ in range(steps):
        y = model(adj, x)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
