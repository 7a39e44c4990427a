"""CPU restatement of the reference's GCN-GRU forward (TEST INFRASTRUCTURE ONLY).

Follows, line by line:

* ``src/step5_gcn_layer_model.py:13-23``  — ``GraphConvLayer.forward``:
  aggregate ``adj @ attr`` first (:15), then ``@ weight + bias`` (:18), then ReLU (:21).
* ``src/step6_gcn_gru_combined_model.py:13-27`` — ``GCN_GRU.forward``: conv1 (:17),
  conv2 + flatten to ``[T, S*13]`` with flat index ``s*13 + f`` (:20), single-layer
  ``nn.GRU(batch_first=True)`` with ``h0 = 0`` (:11, :23), return every hidden state (:26-27).
* The GRU cell arithmetic is PyTorch's (third party, version not pinned by the
  reference; torch 2.11.0 here): gates ordered r, z, n along the rows of
  ``weight_ih_l0 [3H, I]`` / ``weight_hh_l0 [3H, H]``;
  ``r = σ(gi_r + gh_r)``, ``z = σ(gi_z + gh_z)``, ``n = tanh(gi_n + r ⊙ gh_n)``,
  ``h' = (h - n) ⊙ z + n``  where ``gi = W_ih u + b_ih`` and ``gh = W_hh h + b_hh``.

The reference only accepts a leading batch of exactly 1 (``view(1, T, S*13)`` at
step6:20).  This restatement is batch-generalised: ``x [B, T, S, F] -> [B, T, H]``;
for ``B == 1`` the reference's ``squeeze(0)`` result is ``out[0]``.

Two implementations are provided:

``gcn_gru_forward``        NumPy, explicit time loop, dtype-selectable (float32 or
                           float64).  Independent of ``torch.nn.GRU``.
``gcn_gru_forward_torch``  the same equations through torch CPU ops (``torch.matmul``,
                           ``torch.relu``, ``torch.nn.functional``-level GRU via
                           ``torch._VF.gru``) — the same library kernels the reference
                           calls, batched.  This is the timed CPU baseline ("port").

Parity pinning: there are no golden vectors in the reference; both functions are
pinned against the reference module's own outputs (``tests/golden``).
"""

from __future__ import annotations

import numpy as np

PARAM_KEYS = (
    "conv1.weight",
    "conv1.bias",
    "conv2.weight",
    "conv2.bias",
    "gru.weight_ih_l0",
    "gru.weight_hh_l0",
    "gru.bias_ih_l0",
    "gru.bias_hh_l0",
)


def _as_np(v, dtype):
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(v), dtype=dtype)


def gcn_layer(adj, attr, weight, bias):
    """``relu((adj @ attr) @ weight + bias)`` — step5:15,18,21.

    ``adj [S,S]``, ``attr [..., S, Fin]``, ``weight [Fin, Fout]`` (in×out, not
    transposed), ``bias [Fout]``.
    """
    adj_attr = np.matmul(adj, attr)  # step5:15
    out = np.matmul(adj_attr, weight) + bias  # step5:18
    return np.maximum(out, 0)  # step5:21


def _sigmoid(v):
    return 1.0 / (1.0 + np.exp(-v))


def gcn_gru_forward(adj, x, params, dtype=np.float32, h0=None):
    """Batch-generalised restatement of ``GCN_GRU.forward`` (step6:13-27).

    adj    [S, S]
    x      [B, T, S, F_in]
    params mapping with the reference's ``state_dict`` keys (``PARAM_KEYS``)
    returns ``[B, T, H]`` in ``dtype``.
    """
    dt = np.dtype(dtype)
    adj = _as_np(adj, dt)
    x = _as_np(x, dt)
    w1, b1, w2, b2, w_ih, w_hh, b_ih, b_hh = (_as_np(params[k], dt) for k in PARAM_KEYS)
    if x.ndim != 4:
        raise ValueError("x must be [B, T, S, F_in]")
    B, T, S, _ = x.shape
    H = w_hh.shape[1]
    hidden1 = gcn_layer(adj, x, w1, b1)  # step6:17
    hidden2 = gcn_layer(adj, hidden1, w2, b2)  # step6:20
    u = hidden2.reshape(B, T, S * hidden2.shape[-1])  # flat index s*F_out + f
    if u.shape[-1] != w_ih.shape[1]:
        raise ValueError("gru_input does not match S * output_dim")
    # input projection for every step at once (no dependence on h)
    gi = np.matmul(u, w_ih.T) + b_ih  # [B, T, 3H]
    h = np.zeros((B, H), dtype=dt) if h0 is None else _as_np(h0, dt).copy()
    out = np.empty((B, T, H), dtype=dt)
    for t in range(T):
        gh = np.matmul(h, w_hh.T) + b_hh
        r = _sigmoid(gi[:, t, 0:H] + gh[:, 0:H])
        z = _sigmoid(gi[:, t, H : 2 * H] + gh[:, H : 2 * H])
        n = np.tanh(gi[:, t, 2 * H : 3 * H] + r * gh[:, 2 * H : 3 * H])
        h = ((h - n) * z + n).astype(dt, copy=False)
        out[:, t, :] = h
    return out


def gcn_gru_forward_torch(adj, x, params, dtype=None):
    """Same equations through torch CPU library ops (the kernels the reference calls).

    Accepts torch tensors (or array-likes); returns a torch tensor ``[B, T, H]``.
    """
    import torch

    dt = dtype or torch.float32
    tt = lambda v: torch.as_tensor(v).detach().to(device="cpu", dtype=dt)  # noqa: E731
    adj = tt(adj)
    x = tt(x)
    w1, b1, w2, b2, w_ih, w_hh, b_ih, b_hh = (tt(params[k]) for k in PARAM_KEYS)
    B, T, S, _ = x.shape
    with torch.no_grad():
        g1 = torch.relu(torch.matmul(torch.matmul(adj, x), w1) + b1)
        g2 = torch.relu(torch.matmul(torch.matmul(adj, g1), w2) + b2)
        u = g2.reshape(B, T, -1)
        h0 = torch.zeros(1, B, w_hh.shape[1], dtype=dt)
        out, _ = torch._VF.gru(
            u, h0, [w_ih, w_hh, b_ih, b_hh], True, 1, 0.0, False, False, True
        )
    return out


def normalised_max_error(y, ref):
    """``max|y - ref| / max|ref|`` — the parity metric of SURVEY.md §8(c).

    Element-wise relative error is meaningless here: GRU states cross zero.
    """
    y = np.asarray(y, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(y - ref)) / np.max(np.abs(ref)))
