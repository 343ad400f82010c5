"""ggad_b200 -- B200-native GGAD message passing + outlier synthesis (hot path only).

Layout
  csrc/ + libggad_b200.so   hand-written sm_100a kernels behind the C ABI of include/ggad_b200.h
  _lib.py                   ctypes binding (raises if the library is missing: no CPU fallback)
  graph.py                  device CSR container, merge-path plan, host index work
  ops.py                    torch.autograd.Function ops (custom backward on the transposed CSR)
  model.py                  drop-in for the reference's model.py      (program A, full batch)
  losses.py                 the loss block of run.py:164-210 on CSR
  graphsage.py              drop-in for the reference's src/graphsage.py (program B, mini batch)
  dist.py                   node-range sharding + the per-layer collective
  synth.py                  synthetic graphs of the BASELINE.json shapes
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
from .graph import CSRGraph, full_batch_graphs  # noqa: F401
