#!/usr/bin/env python
"""Generates the golden fixtures in this directory.  Runs ONLY in the authoring container
(needs /root/reference and oracle/_ref/libhypo_ref.so); the fixtures are committed because
neither travels to the GPU box.

  spoa_global_consensus.json.gz
      Inputs + expected output of the reference's own known-answer test
      SpoaAlignmentTest.GlobalConsensus (reference external/spoa/test/spoa_test.cpp:220-239):
      the 55 reads of external/spoa/test/data/sample.fastq, kNW, linear 5/-4/-8, and the
      469-character consensus the test asserts.  The expected string is parsed out of the
      test source at generation time and cross-checked against the compiled reference.

  windows.json.gz
      Windows of every kind (internal / draft-backbone / prefix-heavy / suffix-heavy / mixed /
      N-in-draft / LONG two-round + curate / alternative scores / hand-written edge cases) with
      the consensus produced by the compiled, UNMODIFIED reference (default flags => SISD).
      LONG windows only contain the arms the reference's insert-time Filter accepted.

  inspect_ref.txt.gz
      A window stream in the reference's dump format (Contig::generate_inspect_file, reference
      src/Contig.cpp:368-453): every window printed by the reference's own Window::operator<<
      (src/Window.cpp:63-84) after the reference computed its consensus (oracle/ref_driver.cpp:
      hypo_ref_inspect_dump).  Input and expected output of hypo_b200/host/WindowStream and
      tools/replay_inspect.py.

Usage:  make -C oracle ref && python tests/golden/make_golden.py [inspect]
"""
import gzip
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from hypo_b200.batch import WINDOW_LONG, WindowSpec, build_batch  # noqa: E402
from hypo_b200.synth import edge_case_windows, random_window  # noqa: E402
from tests.oracle_util import DEFAULT_SCORES, drop_rejected_arms, ref_consensus, ref_spoa  # noqa: E402

REF = "/root/reference/external/spoa/test"
GENERATOR_VERSION = 1


def spoa_fixture():
    lines = open(os.path.join(REF, "data", "sample.fastq")).read().split("\n")
    seqs = [lines[i + 1] for i in range(0, len(lines) - 3, 4) if lines[i].startswith("@")]
    src = open(os.path.join(REF, "spoa_test.cpp")).read()
    body = src[src.index("TEST_F(SpoaAlignmentTest, GlobalConsensus)"):]
    body = body[: body.index("EXPECT_TRUE")]
    assert "5, -4, -8, -8, -8, -8" in body and "kNW" in body
    expected = "".join(re.findall(r'"([ACGT]+)"', body[body.index("valid_result"):]))
    got = ref_spoa(seqs, 5, -4, -8)
    assert got == expected, "compiled reference disagrees with its own test?"
    return {"source": "reference external/spoa/test/spoa_test.cpp:220-239 + test/data/sample.fastq",
            "m": 5, "n": -4, "g": -8, "seqs": seqs, "expected": expected}


def window_fixture():
    rng = np.random.default_rng(20261017)
    groups = []

    def add(name, specs, scores=DEFAULT_SCORES):
        batch = build_batch(specs)
        cons, acc, _ = ref_consensus(batch, scores)
        if not acc.all():
            batch = drop_rejected_arms(batch, acc)
            cons, acc, _ = ref_consensus(batch, scores)
            assert acc.all()
        wins = []
        for w in range(batch.n_win):
            s = batch.spec(w)
            wins.append({"draft": s.draft, "internal": list(s.internal), "pre": list(s.pre),
                         "suf": list(s.suf), "n_empty": s.n_empty, "wtype": s.wtype,
                         "consensus": cons[w]})
        groups.append({"name": name, "scores": list(scores), "windows": wins})

    add("edge_cases", edge_case_windows())
    for kind in ("internal", "backbone", "prefix", "suffix", "mixed"):
        specs = [random_window(rng, length=120, n_arms=30, kind=kind) for _ in range(4)]
        specs += [random_window(rng, length=int(rng.integers(5, 70)), n_arms=int(rng.integers(3, 25)),
                                kind=kind, err=float(rng.choice([0.01, 0.05, 0.15]))) for _ in range(28)]
        add(kind, specs)
    add("n_in_draft", [random_window(rng, length=40, n_arms=8, kind="backbone", draft_n=0.1) for _ in range(12)])
    add("long", [random_window(rng, length=int(rng.integers(120, 400)), n_arms=int(rng.integers(4, 16)),
                               kind=str(rng.choice(["internal", "mixed", "prefix"])), wtype=WINDOW_LONG,
                               err=float(rng.choice([0.01, 0.05]))) for _ in range(16)])
    add("alt_scores", [random_window(rng, length=40, n_arms=10, kind="mixed", err=0.1) for _ in range(16)],
        scores=(2, -3, -2, 1, -1, -1))
    add("alt_scores_long", [random_window(rng, length=150, n_arms=8, kind="mixed", err=0.05, wtype=WINDOW_LONG)
                            for _ in range(8)], scores=(1, -1, -1, 2, -2, -3))
    return {"generator_version": GENERATOR_VERSION, "oracle": "oracle/_ref/libhypo_ref.so (SISD)",
            "groups": groups}


def large_window_fixture():
    """Windows of the sizes the multi-tile and global-memory capacity tiers run (250-500 bp, up to 100 reads,
    1-6 % read error, SHORT and LONG), consensus by the compiled reference.  Kept in a file of its own."""
    rng = np.random.default_rng(20261020)
    groups = []

    def add(name, specs, scores=DEFAULT_SCORES):
        batch = build_batch(specs)
        cons, acc, _ = ref_consensus(batch, scores)
        if not acc.all():
            batch = drop_rejected_arms(batch, acc)
            cons, acc, _ = ref_consensus(batch, scores)
            assert acc.all()
        wins = []
        for w in range(batch.n_win):
            s = batch.spec(w)
            wins.append({"draft": s.draft, "internal": list(s.internal), "pre": list(s.pre),
                         "suf": list(s.suf), "n_empty": s.n_empty, "wtype": s.wtype,
                         "consensus": cons[w]})
        groups.append({"name": name, "scores": list(scores), "windows": wins})

    add("short_250", [random_window(rng, length=250, n_arms=30, kind=str(rng.choice(["internal", "mixed"])),
                                    err=float(rng.choice([0.01, 0.05]))) for _ in range(8)])
    add("short_500", [random_window(rng, length=int(rng.integers(400, 500)), n_arms=int(rng.integers(8, 30)),
                                    kind=str(rng.choice(["internal", "prefix", "suffix"])),
                                    err=float(rng.choice([0.01, 0.03, 0.06]))) for _ in range(8)])
    add("many_reads", [random_window(rng, length=int(rng.integers(60, 125)), n_arms=100, kind="mixed",
                                     err=float(rng.choice([0.01, 0.04]))) for _ in range(4)])
    add("long_300_500", [random_window(rng, length=int(rng.integers(300, 500)), n_arms=int(rng.integers(12, 30)),
                                       kind="internal", wtype=WINDOW_LONG, err=float(rng.choice([0.01, 0.03])))
                         for _ in range(8)])
    return {"generator_version": GENERATOR_VERSION, "oracle": "oracle/_ref/libhypo_ref.so (SISD)",
            "groups": groups}


def inspect_fixture():
    import ctypes as C
    import tempfile
    from tests.oracle_util import ref_lib
    rng = np.random.default_rng(20261018)
    specs = list(edge_case_windows())
    for kind in ("internal", "backbone", "prefix", "suffix", "mixed"):
        specs += [random_window(rng, length=int(rng.integers(5, 110)), n_arms=int(rng.integers(3, 36)), kind=kind,
                                err=float(rng.choice([0.01, 0.05]))) for _ in range(10)]
    specs += [random_window(rng, length=int(rng.integers(120, 300)), n_arms=int(rng.integers(4, 12)),
                            kind=str(rng.choice(["internal", "mixed"])), wtype=WINDOW_LONG) for _ in range(6)]
    batch = build_batch(specs)
    lib = ref_lib(False)
    sc = (C.c_int8 * 6)(*DEFAULT_SCORES)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "inspect_ctg.txt")
        rc = lib.hypo_ref_inspect_dump(path.encode(), b"ctg_golden", sc, batch.win.ctypes.data_as(C.c_void_p),
                                       C.c_uint64(batch.n_win), batch.arms.ctypes.data_as(C.c_void_p),
                                       batch.packed.ctypes.data_as(C.c_void_p))
        assert rc == 0
        text = open(path, "rb").read()
    with gzip.GzipFile(os.path.join(HERE, "inspect_ref.txt.gz"), "wb", mtime=0) as f:
        f.write(text)
    return batch.n_win, len(text)


def long_filter_fixture():
    """Which arms of LONG windows the reference's Window keeps (hypo::Filter::is_good, reference
    include/Filter.hpp, applied by Window::add_* at include/Window.hpp:66-101): seeded LONG windows at
    several error rates, the arms as strings, and the compiled reference's accept flag per arm."""
    from tests.oracle_util import ref_consensus
    rng = np.random.default_rng(20261019)
    specs = []
    for err in (0.02, 0.04, 0.06, 0.1):
        for kind in ("internal", "mixed"):
            specs += [random_window(rng, length=int(rng.integers(40, 320)), n_arms=int(rng.integers(4, 14)), kind=kind,
                                    err=err, wtype=WINDOW_LONG) for _ in range(8)]
    # a draft with N runs: a non-ACGT character restarts the k-mer but not the filter's minimizer window
    tail = random_window(rng, length=150, n_arms=6, kind="internal", err=0.02, wtype=WINDOW_LONG)
    d = list(tail.draft)
    for p in (17, 18, 60, 95, 96, 97):
        d[p] = "N"
    specs.append(WindowSpec("".join(d), tail.internal, tail.pre, tail.suf, tail.n_empty, tail.wtype))
    batch = build_batch(specs)
    _, acc, _ = ref_consensus(batch)
    wins = []
    for w in range(batch.n_win):
        s = batch.spec(w)
        a0 = int(batch.win[w]["first_arm"])
        n = len(s.internal) + len(s.pre) + len(s.suf)
        wins.append({"draft": s.draft, "internal": s.internal, "pre": s.pre, "suf": s.suf,
                     "accepted": [int(x) for x in acc[a0:a0 + n]]})
    return {"generator_version": GENERATOR_VERSION, "oracle": "oracle/_ref/libhypo_ref.so", "windows": wins}


def dump(name, obj):
    with gzip.GzipFile(os.path.join(HERE, name), "wb", mtime=0) as f:
        f.write(json.dumps(obj, separators=(",", ":")).encode())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "large":
        lw = large_window_fixture()
        dump("windows_large.json.gz", lw)
        print("large windows:", sum(len(g["windows"]) for g in lw["groups"]))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "filter":
        lf = long_filter_fixture()
        dump("long_filter.json.gz", lf)
        flags = [x for w in lf["windows"] for x in w["accepted"]]
        print("long filter: %d windows, %d arms, %d rejected" % (len(lf["windows"]), len(flags), flags.count(0)))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "inspect":
        print("inspect stream: %d windows, %d bytes" % inspect_fixture())
        sys.exit(0)
    dump("spoa_global_consensus.json.gz", spoa_fixture())
    wf = window_fixture()
    dump("windows.json.gz", wf)
    print("windows:", sum(len(g["windows"]) for g in wf["groups"]))
    print("inspect stream: %d windows, %d bytes" % inspect_fixture())
    dump("long_filter.json.gz", long_filter_fixture())
    dump("windows_large.json.gz", large_window_fixture())
