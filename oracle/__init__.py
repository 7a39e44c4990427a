"""CPU oracle for the WindGNN GCN-GRU forward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``windgnn_b200/`` imports this package.
It may be imported by ``tests/``, by ``__graft_entry__.smoke()`` (as the checker)
and by ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs (as the timed
CPU arm).  The product path (``windgnn_b200``) is CUDA-only and raises when its
extension is missing.

Parity pinning: the reference ships no tests or golden vectors, so
parity is *unpinned by the reference's own tests*.  The oracle is instead pinned
against outputs of the reference module itself, run in the build container with
the two shipped checkpoints (``tests/golden/make_golden.py`` generates, and
``tests/test_oracle.py`` checks, those fixtures).
"""

from .gcn_gru_oracle import (  # noqa: F401
    gcn_layer,
    gcn_gru_forward,
    gcn_gru_forward_torch,
    normalised_max_error,
)
from .graph_oracle import (  # noqa: F401
    mercator,
    dense_graph_f64,
    knn_graph_f64,
    synthetic_coordinates,
)
from .windows_oracle import create_sequences, denorm_last_step, pivot_long_table  # noqa: F401,E402
