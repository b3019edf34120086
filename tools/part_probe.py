"""Scaling probe on ONE GPU: time skb_triangle(part=0, n_parts=N) on the bench workload for N = 1, 2, 4, 8.
The ideal is T(1)/N; what is above it is per-call fixed cost plus lost L2 reuse, i.e. the scaling loss the
N-GPU bench will show, without needing N GPUs.  Usage: python tools/part_probe.py [workload]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from skder_b200 import build, engine, synth  # noqa: E402

build.build()
wl = sys.argv[1] if len(sys.argv) > 1 else "config3"
eng = engine.Engine(0)
packed = synth.config_packed(wl, lambda g: engine.pack_contigs(g, eng.params.min_contig_len))
eng.add(packed)
eng.index()
base = None
for n_parts in (1, 2, 4, 8):
    for part in sorted({0, n_parts - 1}):
        ms = []
        for _ in range(4):
            edges, st = eng.triangle(89.5, 50.0, part=part, n_parts=n_parts)
            ms.append((st.ms_screen, st.ms_ani))
        scr, ani = np.median([m[0] for m in ms[1:]]), np.median([m[1] for m in ms[1:]])
        if base is None:
            base = scr + ani
        print("n_parts %d part %d: screen %.2f ms  ani %.2f ms  pairs %d  -> %.1f%% of ideal" % (
            n_parts, part, scr, ani, st.n_pairs_screened, 100.0 * base / n_parts / (scr + ani)), flush=True)
