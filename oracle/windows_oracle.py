"""CPU restatement of the reference's window extraction (TEST INFRASTRUCTURE ONLY).

Follows ``__create_sequences`` (src/step4_sequence_preparer.py:7-27): ``r = floor(len / L)``
windows (:10); window ``i``: ``x = data[i*L:(i+1)*L, :, 2:15]`` (:13) and
``y = concat(data[i*L+k:(i+1)*L+k, :, 13] for k in 1, 2, 3)`` along the station axis (:14-18);
then the windows are permuted by a shuffled index vector (:23-26) — passed in explicitly here,
because the reference's shuffle is unseeded.  ``denorm_last_step`` restates
``outputs.cpu().numpy() * (wind_max - wind_min) + wind_min`` (src/main.py:103) for the last
timestep (main.py:116).  Pinned by ``tests/golden/windows.npz`` (made from the reference itself).
"""

from __future__ import annotations

import numpy as np


def create_sequences(data, seq_length, indices=None):
    """``data [Ttot, S, 15]`` (reference column layout) -> ``(x [r, L, S, 13], y [r, L, 3S])``."""
    data = np.asarray(data)
    r = int((len(data) - len(data) % seq_length) / seq_length)  # step4:10
    xs, ys = [], []
    for i in range(r):
        x = data[(i * seq_length):((i + 1) * seq_length), :, 2:15]
        y1 = data[(i * seq_length + 1):((i + 1) * seq_length + 1), :, 13]
        y2 = data[(i * seq_length + 2):((i + 1) * seq_length + 2), :, 13]
        y3 = data[(i * seq_length + 3):((i + 1) * seq_length + 3), :, 13]
        xs.append(x)
        ys.append(np.concatenate((np.concatenate((y1, y2), axis=1), y3), axis=1))
    xs1, ys1 = np.array(xs), np.array(ys)
    if indices is not None:
        xs1, ys1 = xs1[indices], ys1[indices]
    return xs1, ys1


def denorm_last_step(outputs, wind_min, wind_max):
    """``outputs [B, T, H]`` float32 -> ``[B, H]``; NumPy float32-array x python-float arithmetic."""
    outputs = np.asarray(outputs, dtype=np.float32)
    full = outputs * (wind_max - wind_min) + wind_min  # main.py:103
    return full[:, -1, :]


def pivot_long_table(long_table):
    """step4:36-47 — ``long_table [n_rows, C]`` (column 0 = station) -> ``[time, station, C]``: stations in
    ``np.unique`` order (:38), each station's rows in file order (:41), stacked along a new axis 1 (:42-47)."""
    import numpy as np

    stations = np.unique(long_table[:, 0])
    cols = [long_table[np.where(long_table[:, 0] == st)] for st in stations]
    return np.stack(cols, axis=1), stations
