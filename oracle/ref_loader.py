"""Loader of the unmodified reference classes from ``oracle/_ref/`` (TEST / MEASUREMENT ONLY).

``reference_classes()`` returns ``(GCN_GRU, GraphConvLayer)`` of the reference
(src/step6_gcn_gru_combined_model.py:6, src/step5_gcn_layer_model.py:5) or ``None`` when
``oracle/_ref/`` has not been built (``oracle/build_ref.py``).

``reference_forward_batched`` runs a batch through the reference model's OWN sub-modules: the
reference's ``forward`` accepts a leading batch of exactly 1 (``view(1, T, S*13)``, step6:20), so a
batch goes through ``model.conv1`` / ``model.conv2`` / ``model.gru`` with the flatten generalised to
``reshape(B, T, -1)`` — same modules, same library kernels, same parameters (equal
to the per-window loop to 1.3e-6 normalised).  ``reference_forward_as_written`` is the loop of
src/main.py:101-102: one ``model(adj_matrix, batch_x)`` call per window.
"""

from __future__ import annotations

import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def reference_classes():
    if not os.path.exists(os.path.join(REF_DIR, "step6_gcn_gru_combined_model.py")):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)   # step6 does `from step5_gcn_layer_model import *`
    step6 = importlib.import_module("step6_gcn_gru_combined_model")
    step5 = importlib.import_module("step5_gcn_layer_model")
    return step6.GCN_GRU, step5.GraphConvLayer


def reference_model(state_dict, dims, device="cpu"):
    """The reference ``GCN_GRU(*dims)`` with ``state_dict`` loaded (main.py:96-99), or None."""
    classes = reference_classes()
    if classes is None:
        return None
    model = classes[0](*dims)
    model.load_state_dict(state_dict, strict=True)
    return model.to(device).eval()


def reference_forward_batched(model, adj, x):
    import torch

    with torch.no_grad():
        B, T = x.shape[0], x.shape[1]
        hidden2 = model.conv2(adj, model.conv1(adj, x))          # step6:17,20
        return model.gru(hidden2.reshape(B, T, -1))[0]           # step6:20 generalised, :23


def reference_forward_as_written(model, adj, x):
    import torch

    with torch.no_grad():
        return torch.stack([model(adj, x[b:b + 1]) for b in range(x.shape[0])])   # main.py:101-102
