"""colord_b200 — B200-native hot path of a CoLoRd-compatible long-read compressor.

The product is the C-ABI library ``libcolord_b200.so`` (include/colord_b200.h; CUDA, sm_100a only).
``colord_b200.lib`` is a thin ctypes binding used by the tests and bench.py; there is no CPU path.
"""
