"""A few forwards of the bench workload (configs[1]) — the command profiled under ncu."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, windgnn_b200

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
precision = sys.argv[3] if len(sys.argv) > 3 else "fp32"
sd, latlon = bench.load_workload()
dev = torch.device("cuda:0")
model = windgnn_b200.GCN_GRU(13, 13, 13, 442, 102)
model.load_state_dict(sd)
model = model.to(dev).eval()
model.precision = precision
adj = windgnn_b200.build_graph_from_latlon(latlon, device=dev)
x = torch.rand((B, 168, 34, 13), device=dev, generator=torch.Generator(device=dev).manual_seed(0))
with torch.no_grad():
    for _ in range(steps):
        y = model(adj, x)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
