#!/usr/bin/env python
"""Developer aid: WindowBatch::run breakdown (pack / device / scatter / total) for several chunk sizes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hypo_b200 import native
from hypo_b200.hostlib import HostWindows, synth_batch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
native.init((5, -4, -8, 3, -5, -4), 0)
b = synth_batch(2026, n, 120, 30, "internal", 0.01)
hw = HostWindows(b)
for chunk in [0] + [int(x) for x in sys.argv[2:]]:
    hw.run(chunk)
    best = None
    for _ in range(3):
        t = hw.run(chunk)
        if best is None or t["total_s"] < best["total_s"]:
            best = t
    print(json.dumps({"chunk": chunk, "mbp_s": b.polished_bp / 1e6 / best["total_s"], **best}))
