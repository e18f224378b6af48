"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle, the
committed golden vectors (produced by the compiled reference) and size-independent properties.
Bar: bit-exact consensus strings (integer/byte path)."""
import numpy as np
import pytest

from hypo_b200 import native
from hypo_b200.batch import WINDOW_LONG, WindowSpec, build_batch
from hypo_b200.synth import edge_case_windows, random_batch, random_window
from tests.golden_util import group_batch
from tests.oracle_util import DEFAULT_SCORES, oracle_consensus

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _init():
    native.init(DEFAULT_SCORES, 0)
    yield
    native.shutdown()


def _assert_same(got, want, batch, label=""):
    bad = [i for i, (a, b) in enumerate(zip(got, want)) if a != b]
    if bad:
        i = bad[0]
        raise AssertionError(f"{label}: {len(bad)}/{len(want)} windows differ; first={i}\n"
                             f"spec={batch.spec(i)}\n gpu={got[i]!r}\nwant={want[i]!r}")


def test_golden_vectors_from_compiled_reference(golden_windows):
    for group in golden_windows["groups"]:
        batch, expected = group_batch(group)
        native.init(tuple(group["scores"]), 0)
        _assert_same(native.consensus(batch), expected, batch, group["name"])
    native.init(DEFAULT_SCORES, 0)


def test_golden_large_windows_from_compiled_reference(golden_windows_large):
    """The multi-tile and global-memory tiers against consensus strings produced by the reference itself."""
    for group in golden_windows_large["groups"]:
        batch, expected = group_batch(group)
        native.init(tuple(group["scores"]), 0)
        _assert_same(native.consensus(batch), expected, batch, group["name"])
    native.init(DEFAULT_SCORES, 0)


def test_edge_cases():
    b = build_batch(edge_case_windows())
    want, _ = oracle_consensus(b)
    _assert_same(native.consensus(b), want, b, "edge")


@pytest.mark.parametrize("kind", ["internal", "backbone", "prefix", "suffix", "mixed"])
def test_short_kinds_vs_oracle(kind):
    for seed, kw in ((1, dict(length=120, n_arms=30)), (2, dict(length=40, n_arms=12, err=0.06)),
                     (3, dict(length=9, n_arms=30, err=0.03)), (4, dict(length=100, n_arms=30, err=0.05))):
        b = random_batch(seed * 17 + 3, 96, kind=kind, **kw)
        want, _ = oracle_consensus(b)
        _assert_same(native.consensus(b), want, b, f"{kind}/{kw}")


def test_long_windows_vs_oracle():
    for seed, kw in ((5, dict(length=200, n_arms=12, kind="mixed")), (6, dict(length=480, n_arms=20, kind="internal")),
                     (7, dict(length=150, n_arms=8, kind="internal", err=0.08))):
        b = random_batch(seed, 24, wtype=WINDOW_LONG, **kw)
        want, _ = oracle_consensus(b)
        _assert_same(native.consensus(b), want, b, f"long/{kw}")


def test_overflow_tiers_vs_oracle():
    """Windows that overflow the shared-memory tier (many nodes / long arms) are re-run in the
    larger tiers on the device, never on the CPU."""
    b = random_batch(8, 12, kind="internal", length=120, n_arms=60, err=0.12)   # > 320 nodes
    want, _ = oracle_consensus(b)
    _assert_same(native.consensus(b), want, b, "many-nodes")
    b = random_batch(9, 8, kind="mixed", length=300, n_arms=10, err=0.03)       # SHORT but > 127 columns
    want, _ = oracle_consensus(b)
    _assert_same(native.consensus(b), want, b, "wide-short")


def test_alternative_scores():
    sc = (2, -3, -2, 1, -1, -1)
    native.init(sc, 0)
    b = random_batch(10, 64, kind="mixed", length=50, n_arms=12, err=0.1)
    want, _ = oracle_consensus(b, sc)
    _assert_same(native.consensus(b), want, b, "alt-scores")
    native.init(DEFAULT_SCORES, 0)


def test_mixed_batch_and_order_independence():
    rng = np.random.default_rng(11)
    specs = edge_case_windows()
    for _ in range(60):
        specs.append(random_window(rng, length=int(rng.integers(4, 130)), n_arms=int(rng.integers(2, 34)),
                                   kind=str(rng.choice(["internal", "mixed", "prefix", "suffix", "backbone"])),
                                   err=float(rng.choice([0.01, 0.05]))))
    b = build_batch(specs)
    want, _ = oracle_consensus(b)
    got = native.consensus(b)
    _assert_same(got, want, b, "mixed-batch")
    perm = rng.permutation(len(specs))
    b2 = build_batch([specs[i] for i in perm])
    got2 = native.consensus(b2)
    assert [got[i] for i in perm] == got2


def test_identical_arms_property():
    """Size-independent property: if every arm equals the truth, the consensus is the truth."""
    rng = np.random.default_rng(12)
    specs, truths = [], []
    for _ in range(64):
        L = int(rng.integers(1, 126))
        truth = "".join("ACGT"[i] for i in rng.integers(0, 4, size=L))
        truths.append(truth)
        specs.append(WindowSpec(truth[::-1] or "A", [truth] * int(rng.integers(2, 20)), [], [], 0, 0))
    assert native.consensus(build_batch(specs)) == truths


def test_errors_are_loud():
    with pytest.raises(native.HypoGpuError):
        native.init((5, -4, 8, 3, -5, -4), 0)
    native.init(DEFAULT_SCORES, 0)
    b = random_batch(13, 4)
    b.arms["off"][0] = 1 << 40
    with pytest.raises(native.HypoGpuError):
        native.consensus(b)


def test_large_random_parity_stresses_concurrent_sort():
    """Thousands of windows per shape, including high-error ones whose graphs are full of cliques and
    insertion chains: this is what exercises the warp-parallel topological sort (claims, conflicts,
    serial fallback) under real concurrency.  Every window is compared with the oracle."""
    from hypo_b200.hostlib import synth_batch
    for seed, kw in ((31, dict(n_win=20000, length=120, n_arms=30, kind="internal", err=0.01)),
                     (32, dict(n_win=3000, length=110, n_arms=30, kind="mixed", err=0.05)),
                     (33, dict(n_win=3000, length=60, n_arms=25, kind="internal", err=0.12)),
                     (34, dict(n_win=6000, length=12, n_arms=35, kind="mixed", err=0.03))):
        b = synth_batch(seed, **kw)
        want, _ = oracle_consensus(b)
        _assert_same(native.consensus(b), want, b, f"large/{kw}")


def test_mid_size_windows_and_growth_projection():
    """250-500 bp windows (CCS-sized SHORT windows, LONG windows): the four-tile tier; at 5 % read error
    the DAG outgrows it, which the growth projection notices after a few reads (reason 10) - the window
    moves on with its projection (reason 11 where a later tier cannot hold it either) and the result
    stays bit-exact."""
    from hypo_b200.hostlib import synth_batch
    projected = 0
    for seed, kw in ((51, dict(n_win=96, length=250, n_arms=30, kind="internal", err=0.01)),
                     (52, dict(n_win=64, length=250, n_arms=30, kind="internal", err=0.05)),
                     (53, dict(n_win=48, length=500, n_arms=12, kind="mixed", err=0.03)),
                     (54, dict(n_win=32, length=400, n_arms=20, kind="internal", err=0.04, wtype=WINDOW_LONG))):
        b = synth_batch(seed, **kw)
        want, _ = oracle_consensus(b)
        _assert_same(native.consensus(b), want, b, f"mid/{kw}")
        _, _, tiers = native.last_timing()
        reasons = native.last_fail_hist()
        assert sum(reasons) == sum(tiers) - b.n_win
        projected += reasons[10]
    assert projected > 0


def test_many_read_windows_take_the_wide_tier():
    """100-200 reads per window: routed past the compact tiers by the node estimate (or abandoned
    there and re-run), bit-exact either way; the diagnostics explain every abandonment."""
    from hypo_b200.hostlib import synth_batch
    for seed, kw in ((41, dict(n_win=160, length=100, n_arms=100, kind="internal", err=0.01)),
                     (42, dict(n_win=120, length=50, n_arms=200, kind="mixed", err=0.02)),
                     (43, dict(n_win=300, length=120, n_arms=30, kind="internal", err=0.05))):
        b = synth_batch(seed, **kw)
        want, _ = oracle_consensus(b)
        _assert_same(native.consensus(b), want, b, f"wide/{kw}")
        _, _, tiers = native.last_timing()
        reasons = native.last_fail_hist()
        assert sum(tiers) >= b.n_win                      # every window ran, some in more than one tier
        assert sum(reasons) == sum(tiers) - b.n_win       # one recorded reason per abandonment


def test_clique_beyond_acgt_leaves_the_compact_tier():
    """An N in a backbone draft aligned with all four bases makes a five-letter clique: more peers than
    the compact tier's aligned-list blocks hold, so the window is re-run in a tier with wider blocks."""
    arms = ["AAAAAAAAA", "AAAACAAAA", "AAAAGAAAA", "AAAATAAAA"]
    specs = [WindowSpec("AAAANAAAA", [], arms * 2, arms * 2, 0, 0) for _ in range(8)]
    b = build_batch(specs)
    want, _ = oracle_consensus(b)
    _assert_same(native.consensus(b), want, b, "five-letter clique")


def test_regression_long_window_with_borderline_support():
    """Found by tools/fuzz_parity.py: a LONG window whose first consensus base sits exactly on the
    curation threshold; any deviation from spoa's exact topological order (MSA column ids) drops it.
    Many copies, so that every warp slot of a CTA sees it."""
    from tests.regression_specs import LONG_BORDERLINE_SUPPORT
    b = build_batch([LONG_BORDERLINE_SUPPORT] * 48)
    want, _ = oracle_consensus(b)
    assert want[0] == "AGGGGGAACGGTGTCTCACCCTTACGGGCCTTTAACTTTCGGAAAAAT"
    _assert_same(native.consensus(b), want, b, "long/borderline-support")
