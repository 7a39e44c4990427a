"""A few training steps (B sequences of the 34-station model) for ncu launch lists.  Usage:
    python scripts/train_step_once.py [steps] [B]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import windgnn_b200  # noqa: E402
from windgnn_b200 import train  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
S, T = 34, 168
sd = torch.load(os.path.join(ROOT, "tests/golden/wind_gnn_34.pth"), map_location="cpu", weights_only=True)
adj = torch.from_numpy(np.load(os.path.join(ROOT, "tests/golden/adj_ref_34.npy")).astype(np.float32)).to(dev)
model = windgnn_b200.GCN_GRU(13, 13, 13, 13 * S, 3 * S)
model.load_state_dict(sd)
model = model.to(dev)
tr = train.Trainer(model, adj)
x = torch.rand((B, T, S, 13), device=dev)
y = torch.rand((B, T, 3 * S), device=dev)
for _ in range(steps):
    loss = tr.step(x, y)
torch.cuda.synchronize()
print("loss", loss.item())
