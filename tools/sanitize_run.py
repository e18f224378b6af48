"""Small mixed batch through the C ABI (run under compute-sanitizer by tools/sanitize.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hypo_b200 import native
from hypo_b200.batch import build_batch, WINDOW_LONG
from hypo_b200.synth import edge_case_windows, random_window
from tests.oracle_util import oracle_consensus, DEFAULT_SCORES

rng = np.random.default_rng(3)
specs = edge_case_windows()
for kind in ("internal", "backbone", "prefix", "suffix", "mixed"):
    specs += [random_window(rng, length=int(rng.integers(5, 125)), n_arms=int(rng.integers(3, 34)), kind=kind,
                            err=float(rng.choice([0.01, 0.05]))) for _ in range(6)]
specs += [random_window(rng, length=120, n_arms=30, kind="internal") for _ in range(8)]
specs += [random_window(rng, length=120, n_arms=60, kind="internal", err=0.1) for _ in range(2)]      # overflow tiers
specs += [random_window(rng, length=200, n_arms=10, kind="mixed") for _ in range(3)]                  # two tiles
specs += [random_window(rng, length=int(rng.integers(100, 300)), n_arms=8, kind="internal", wtype=WINDOW_LONG) for _ in range(3)]
b = build_batch(specs)
native.init(DEFAULT_SCORES, 0)
got = native.consensus(b)
want, _ = oracle_consensus(b)
bad = sum(a != c for a, c in zip(got, want))
print(f"sanitize_run: {b.n_win} windows, {bad} mismatches, tiers {native.last_timing()[2]}")
sys.exit(1 if bad else 0)
