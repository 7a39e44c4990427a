import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import windgnn_b200
dev = torch.device("cuda:0")
S = 34
sd = torch.load("/root/repo/tests/golden/wind_gnn_34.pth", map_location="cpu", weights_only=True)
adj = torch.from_numpy(np.load("/root/repo/tests/golden/adj_ref_34.npy").astype(np.float32)).to(dev)
m = windgnn_b200.GCN_GRU(13, 13, 13, 13 * S, 3 * S); m.load_state_dict(sd); m = m.to(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
x = torch.rand((B, 168, S, 13), device=dev)
with torch.no_grad():
    for _ in range(3): y = m(adj, x)
torch.cuda.synchronize()
print(y.shape)
