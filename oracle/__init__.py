"""CPU oracle for the Point-GNN per-frame message-passing hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, on the CPU, the algorithm of
the reference files /root/reference/models/graph_gen.py, models/gnn.py and
models/models.py::predict.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` leg may import it - and there
only as the checker / the timed CPU baseline, never as part of the product path
(``point-gnn_b200`` never imports ``oracle`` and fails loudly when the CUDA
library is missing).

Parity pinning: the reference ships NO golden vectors for
this path.  The graph half of the oracle is pinned against the reference's own
``models/graph_gen.py`` executed in the build container (scikit-learn present;
fixtures + generating script in tests/golden/).  The GNN half restates TF-1.15
graph ops that cannot be executed here (no TensorFlow wheel): it is pinned only
by the reference's trained checkpoints (weight shapes / scope names / concat
order) -> "parity unpinned" for the floating-point half, see DESIGN.md.
"""
