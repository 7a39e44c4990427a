"""Times the training step (forward-with-save, MSE, backward, Adam) on one GPU.  Usage:
    python scripts/train_time.py [B ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import windgnn_b200  # noqa: E402
from windgnn_b200 import train  # noqa: E402

dev = torch.device("cuda:0")
S, T = 34, 168
sd = torch.load(os.path.join(ROOT, "tests/golden/wind_gnn_34.pth"), map_location="cpu", weights_only=True)
adj = torch.from_numpy(np.load(os.path.join(ROOT, "tests/golden/adj_ref_34.npy")).astype(np.float32)).to(dev)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for B in [int(v) for v in sys.argv[1:]] or [512, 4096]:
    model = windgnn_b200.GCN_GRU(13, 13, 13, 13 * S, 3 * S)
    model.load_state_dict(sd)
    model = model.to(dev)
    tr = train.Trainer(model, adj)
    x = torch.rand((B, T, S, 13), device=dev)
    y = torch.rand((B, T, 3 * S), device=dev)
    ps = [p.data for p in tr.params]
    with torch.no_grad():
        out, ws = train.forward_train(adj, x, ps)
        loss, d_out = train.mse_loss_grad(out, y)
        t_inf = timeit(lambda: model(adj, x))
        t_fwd = timeit(lambda: train.forward_train(adj, x, ps))
        t_mse = timeit(lambda: train.mse_loss_grad(out, y))
        t_bwd = timeit(lambda: train.backward(adj, x, ps, out, d_out, ws, tr.grads))
        t_step = timeit(lambda: tr.step(x, y))
    print(f"B={B}: inference fwd {t_inf:.3f} ms | train fwd {t_fwd:.3f} mse {t_mse:.3f} bwd {t_bwd:.3f} "
          f"| full step {t_step:.3f} ms = {B / t_step * 1e3:.0f} seq/s", flush=True)
    del tr, model, x, y, out, ws, d_out
    torch.cuda.empty_cache()
