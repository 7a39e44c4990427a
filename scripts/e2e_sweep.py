"""End-to-end (host buffers) throughput vs pipeline chunk size, next to the raw pinned H2D / D2H rates."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import windgnn_b200
dev = torch.device("cuda:0")
S, T, B = 34, 168, 4096
sd = torch.load(os.path.join(ROOT, "tests/golden/wind_gnn_34.pth"), map_location="cpu", weights_only=True)
adj = torch.from_numpy(np.load(os.path.join(ROOT, "tests/golden/adj_ref_34.npy")).astype(np.float32)).to(dev)
model = windgnn_b200.GCN_GRU(13, 13, 13, 13 * S, 3 * S); model.load_state_dict(sd); model = model.to(dev).eval()
xh = torch.rand((B, T, S, 13)).pin_memory(); oh = torch.empty((B, T, 3 * S)).pin_memory()
xd = torch.empty_like(xh, device=dev); od = torch.empty_like(oh, device=dev)
def tm(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
t = tm(lambda: xd.copy_(xh, non_blocking=True)); print(f"H2D {xh.numel()*4/t/1e9:.1f} GB/s ({t*1e3:.2f} ms)")
t = tm(lambda: oh.copy_(od, non_blocking=True)); print(f"D2H {oh.numel()*4/t/1e9:.1f} GB/s ({t*1e3:.2f} ms)")
for chunk in (128, 256, 512, 1024, 2048):
    model.chunk = chunk
    t = tm(lambda: model.forward_host(adj, xh, oh))
    print(f"chunk {chunk}: {t*1e3:.2f} ms/step  {B/t:.0f} seq/s", flush=True)
