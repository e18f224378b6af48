"""Small run of the group tiers (several windows per warp) against the oracle; used under compute-sanitizer."""
import sys
import numpy as np
sys.path.insert(0, ".")
from hypo_b200 import native
from hypo_b200.batch import build_batch
from hypo_b200.synth import edge_case_windows, random_window
from tests.oracle_util import DEFAULT_SCORES, oracle_consensus

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(5)
kinds = ("internal", "backbone", "prefix", "suffix", "mixed")
specs = list(edge_case_windows())
for i in range(n):
    specs.append(random_window(rng, length=int(rng.integers(1, 56)), n_arms=int(rng.integers(2, 40)), kind=kinds[i % 5],
                               err=float(rng.choice([0.0, 0.02, 0.08]))))
b = build_batch(specs)
native.init(DEFAULT_SCORES, 0)
native.set_option("group_tiers", 2)   # always (by default only batches of >= 131072 windows use the group tiers)
got = native.consensus(b)
_, _, tiers = native.last_timing()
want, _ = oracle_consensus(b)
bad = [i for i, (x, y) in enumerate(zip(got, want)) if x != y]
print("group_sanity:", b.n_win, "windows, tiers", tiers, "reasons", native.last_fail_hist()[:12], "mismatches", len(bad))
for i in bad[:5]:
    print("  window", i, b.spec(i), "got", got[i], "want", want[i])
sys.exit(1 if bad else 0)
