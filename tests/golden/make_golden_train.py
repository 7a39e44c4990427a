"""Generate the training-step golden fixtures from the UNMODIFIED reference.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden_train.py

Writes ``train_{7,34}.npz``: seeded windows ``x``, targets ``y``, and what the reference's own
training-loop body (``src/main.py:64-77``) produces with the shipped checkpoint as the starting
point — ``nn.MSELoss()`` (:49), ``loss.backward()`` (:76), ``torch.optim.Adam(lr=0.001)`` (:52,:77):

``loss``, ``grad__<key>``           loss and autograd gradients of ONE step over all B windows
                                    (the reference module is called once per window, the only shape
                                    it accepts, step6:20; the loss is the mean over all windows'
                                    elements), computed by ``model.double()`` and stored as fp32
``loss_f32``, ``grad32__<key>``     the same from the fp32 module (shows the fp32 noise floor)
``losses3``, ``param3__<key>``      three consecutive Adam steps on the same batch (fp64 module):
                                    the three losses and the parameters afterwards
"""

import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REF, "src"))

from step6_gcn_gru_combined_model import GCN_GRU  # noqa: E402


def batch_loss(model, adj, x, y):
    outs = torch.stack([model(adj, x[b : b + 1]) for b in range(x.shape[0])])  # main.py:66, per window
    return torch.nn.MSELoss()(outs, y)  # main.py:49,72


def main():
    torch.set_num_threads(1)
    for S, B in ((7, 3), (34, 2)):
        sd = torch.load(os.path.join(HERE, f"wind_gnn_{S}.pth"), map_location="cpu", weights_only=True)
        adj = torch.tensor(np.load(os.path.join(HERE, f"adj_ref_{S}.npy"))).float()  # main.py:26
        g = torch.Generator().manual_seed(500 + S)
        x = torch.rand(B, 168, S, 13, generator=g)
        y = torch.rand(B, 168, 3 * S, generator=g)
        out = {"x": x.numpy(), "y": y.numpy()}
        for tag, dt in (("", torch.float64), ("32", torch.float32)):
            model = GCN_GRU(13, 13, 13, 13 * S, 3 * S).to(dt)
            model.load_state_dict(sd, strict=True)
            model.train()
            loss = batch_loss(model, adj.to(dt), x.to(dt), y.to(dt))
            model.zero_grad()
            loss.backward()  # main.py:76
            out["loss" + ("_f32" if tag else "")] = np.float64(loss.item())
            for k, p in model.named_parameters():
                out[f"grad{tag}__" + k.replace(".", "__")] = p.grad.detach().to(torch.float32).numpy()
        model = GCN_GRU(13, 13, 13, 13 * S, 3 * S).double()
        model.load_state_dict(sd, strict=True)
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=0.001)  # main.py:45,52
        losses = []
        for _ in range(3):
            loss = batch_loss(model, adj.double(), x.double(), y.double())
            opt.zero_grad()  # main.py:69
            loss.backward()
            opt.step()  # main.py:77
            losses.append(loss.item())
        out["losses3"] = np.array(losses, dtype=np.float64)
        for k, p in model.named_parameters():
            out["param3__" + k.replace(".", "__")] = p.detach().to(torch.float32).numpy()
        path = os.path.join(HERE, f"train_{S}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), "bytes; loss", out["loss"], "losses3", losses)


if __name__ == "__main__":
    main()
