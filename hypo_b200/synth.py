"""Seeded synthetic window generator for tests (small batches, pure numpy/python).

Distribution follows BASELINE.md §3 / SURVEY.md §8d: truth = iid-uniform ACGT; draft = truth
with 3 % sub / 3 % ins / 3 % del; every arm = (a slice of) truth with `err` sub / ins / del each.
The bench uses the C++ generator in hypo_b200/host (same distribution, much faster).
"""
from __future__ import annotations

import numpy as np

from .batch import WINDOW_LONG, WINDOW_SHORT, WindowBatch, WindowSpec, build_batch

_B = "ACGT"


def mutate(rng: np.random.Generator, s: str, sub: float, ins: float, dele: float) -> str:
    out = []
    for ch in s:
        r = rng.random()
        if r < dele:
            continue
        if r < dele + sub:
            out.append(_B[(_B.index(ch) + int(rng.integers(1, 4))) % 4] if ch in _B else ch)
        else:
            out.append(ch)
        if rng.random() < ins:
            out.append(_B[int(rng.integers(0, 4))])
    return "".join(out)


def random_window(
    rng: np.random.Generator,
    length: int = 120,
    n_arms: int = 30,
    kind: str = "internal",
    err: float = 0.01,
    draft_err: float = 0.03,
    wtype: int = WINDOW_SHORT,
    draft_n: float = 0.0,
) -> WindowSpec:
    """kind: internal | backbone (0 internal, pre+suf only) | prefix (3 int + rest pre) |
    suffix (3 int + rest suf) | mixed (arm r -> r%5: 0,1,2 internal, 3 pre, 4 suf)."""
    truth = "".join(_B[i] for i in rng.integers(0, 4, size=length))
    draft = mutate(rng, truth, draft_err, draft_err, draft_err) or "A"
    if draft_n > 0:
        draft = "".join("N" if rng.random() < draft_n else c for c in draft)
    internal, pre, suf = [], [], []
    for r in range(n_arms):
        if kind == "internal":
            k = 0
        elif kind == "backbone":
            k = 1 if r < n_arms // 2 else 2
        elif kind == "prefix":
            k = 0 if r < 3 else 1
        elif kind == "suffix":
            k = 0 if r < 3 else 2
        elif kind == "mixed":
            k = (0, 0, 0, 1, 2)[r % 5]
        else:
            raise ValueError(kind)
        if k == 0:
            internal.append(mutate(rng, truth, err, err, err))
        elif k == 1:
            cut = int(rng.integers(max(1, length // 2), length)) if length > 1 else 1
            pre.append(mutate(rng, truth[:cut], err, err, err))
        else:
            cut = int(rng.integers(max(1, length // 2), length)) if length > 1 else 1
            suf.append(mutate(rng, truth[length - cut :], err, err, err))
    return WindowSpec(draft, internal, pre, suf, 0, wtype)


def random_batch(seed: int, n: int, **kw) -> WindowBatch:
    rng = np.random.default_rng(seed)
    return build_batch(random_window(rng, **kw) for _ in range(n))


def edge_case_windows() -> list[WindowSpec]:
    """Hand-written windows covering every branch of Window::generate_consensus
    (reference src/Window.cpp:44-61,87-154,156-254)."""
    S, L = WINDOW_SHORT, WINDOW_LONG
    w = [
        WindowSpec("ACGTACGT", [], [], [], 0, S),  # no arms -> draft
        WindowSpec("ACGNNCGT", ["ACGT"], [], [], 0, S),  # 1 arm -> draft (N kept)
        WindowSpec("ACGTACGT", ["ACGT"], [], [], 2, S),  # empties dominate -> ""
        WindowSpec("ACGTACGT", ["ACGTACGT", "ACGTACGT"], [], [], 2, S),  # n_empty == n -> POA
        WindowSpec("ACGTACGT", ["ACGTTCGT", "ACGTTCGT", "ACGTACGT"], [], [], 0, S),
        WindowSpec("A", ["A", "A", "C"], [], [], 0, S),  # single-base window
        WindowSpec("A", ["", "", ""], [], [], 0, S),  # only zero-length arms -> draft
        WindowSpec("ACGTAC", ["", "ACGTAC", "ACGAC"], [], [], 0, S),  # zero-length arm skipped
        WindowSpec("ACGTACGTAC", [], ["ACGTA", "ACGTAC", "ACG"], ["TACGTAC", "GTAC", "CGTAC"], 0, S),  # backbone
        WindowSpec("ACGTACGTAC", ["ACGTACGTAC"], ["ACGTA", "ACGAAC"], ["TACGTAC"], 0, S),
        WindowSpec("ACGTACGTAC", [], ["TTTTT", "GGGGG"], [], 0, S),  # prefixes unlike the draft
        WindowSpec("ACGTACGTAC", [], [], ["TTTTT", "GGGGG"], 0, S),
        WindowSpec("ACGTNNGTAC", [], ["ACGTA", "ACGAAC"], ["TACGTAC"], 0, S),  # N in backbone draft
        WindowSpec("AAAAAAAAAA", ["AAAAAAAAA", "AAAAAAAAAAA", "AAAAAAAAAA"], [], [], 0, S),  # homopolymer ties
        WindowSpec("ACACACACAC", ["ACACACAC", "ACACACACACAC", "CACACACA"], ["ACAC"], ["CACAC"], 0, S),
        WindowSpec("ACGTACGTACGTACGT", ["ACGTACGTACGTACGT", "ACGTACTTACGTACGT"], [], [], 0, L),
        WindowSpec("ACGTACGTACGTACGT", ["ACGTACGTACGTACGT", "ACGTACTTACGTACGT", "ACGTACGTACGACGT"], ["ACGTACG"], ["ACGTACGT"], 0, L),
        WindowSpec("ACGTACGTACGTACGT", ["TTTTTTTT", "GGGGGGGG"], [], [], 0, L),  # curate may drop everything
        WindowSpec("ACGTACGTACGTACGT", ["", ""], [], [], 0, L),  # LONG, no arm added -> draft
        WindowSpec("ACGNACGTACGTNCGT", ["ACGTACGTACGTACGT", "ACGTACGTACGTACGT"], [], [], 0, L),
        WindowSpec("ACGT", [], ["AC", "ACG"], [], 0, L),
    ]
    return w
