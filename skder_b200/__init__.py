"""skder_b200: B200-native all-vs-all genome ANI/AF engine behind skDER's `skani` call sites.

The package holds the CUDA kernels + C-ABI (csrc/, include/skani_b200.h), the ctypes view of that
ABI (_lib.py), the host-side mirror of the skani sub-commands skDER invokes (engine.py, cli.py),
and the synthetic-genome generator used by the benches (synth.py).
"""
__all__ = ["engine", "cli", "synth", "build"]
