"""CPU-side checks of the C++ host layer (generator determinism, batch shapes)."""
import numpy as np

from hypo_b200.hostlib import synth_batch
from tests.oracle_util import oracle_consensus


def test_synth_deterministic_and_thread_independent():
    a = synth_batch(5, 300, 60, 12, "mixed", threads=1)
    b = synth_batch(5, 300, 60, 12, "mixed", threads=4)
    assert a.win.tobytes() == b.win.tobytes() and a.arms.tobytes() == b.arms.tobytes()
    assert a.packed.tobytes() == b.packed.tobytes()
    c = synth_batch(6, 300, 60, 12, "mixed")
    assert c.packed.tobytes() != a.packed.tobytes()


def test_synth_shapes():
    b = synth_batch(1, 50, 120, 30, "internal")
    assert b.n_win == 50 and b.n_arms == 1500
    assert (b.win["n_internal"] == 30).all() and (b.win["n_pre"] == 0).all()
    lens = b.arms["len"]
    assert 100 < lens.mean() < 140
    m = synth_batch(1, 50, 120, 30, "mixed")
    assert (m.win["n_internal"] == 18).all() and (m.win["n_pre"] == 6).all() and (m.win["n_suf"] == 6).all()
    # consensus of clean-ish arms recovers a ~120 bp sequence
    cons, _ = oracle_consensus(b.select(np.arange(4)))
    assert all(100 < len(c) < 140 for c in cons)


def test_concat_batches_and_pipeline_mix():
    """bench.py's pipeline shape mix is a concatenation of fixed-shape batches: descriptors must be
    re-based onto one arm table and one slab without changing any window."""
    import argparse
    import numpy as np
    import bench
    from hypo_b200.batch import concat_batches
    from hypo_b200.synth import random_batch
    a, b = random_batch(1, 5, length=20, n_arms=4), random_batch(2, 7, length=9, n_arms=6, kind="mixed")
    c = concat_batches([a, b])
    assert c.n_win == 12 and c.n_arms == a.n_arms + b.n_arms
    assert [c.spec(i) for i in range(5)] == [a.spec(i) for i in range(5)]
    assert [c.spec(5 + i) for i in range(7)] == [b.spec(i) for i in range(7)]
    args = argparse.Namespace(mix="pipeline", windows=3000, err=0.01, length=120, arms=30, kind="internal")
    m = bench.make_batch(args, 11)
    assert abs(m.n_win - 3000) <= 40
    assert 8 <= np.median(m.win["draft_len"]) <= 16 and m.win["draft_len"].max() <= 130
    m2 = bench.make_batch(args, 11)
    assert (m.win == m2.win).all() and (m.packed == m2.packed).all()   # seeded


def test_batch_packer_round_trip_is_thread_independent():
    """hypo::WindowBatch (the packer behind Window::generate_consensus_batch): windows built through the
    public add_* API flatten to the batch they came from - same descriptors, same bytes per sequence, the
    slab in container order - whatever the number of packing threads."""
    from hypo_b200.hostlib import host_pack

    def seqs(win, arms, packed):
        out = [packed[int(w["draft_off"]): int(w["draft_off"]) + (int(w["draft_len"]) + 1) // 2].tobytes() for w in win]
        out += [packed[int(a["off"]): int(a["off"]) + (int(a["len"]) + 3) // 4].tobytes() for a in arms]
        return out

    for b in (synth_batch(9, 700, 60, 14, "mixed"), synth_batch(10, 64, 300, 25, "internal", wtype=1)):
        ref = None
        for threads in (1, 3, 8):
            win, arms, packed, _ = host_pack(b, threads)
            for f in ("wtype", "n_internal", "n_pre", "n_suf", "n_empty", "draft_len", "first_arm"):
                assert (win[f] == b.win[f]).all(), f
            assert (arms["len"] == b.arms["len"]).all()
            assert seqs(win, arms, packed) == seqs(b.win, b.arms, b.packed)
            # container order, no gaps: draft, then the arms as listed
            pos = 0
            for w in win:
                assert int(w["draft_off"]) == pos
                pos += (int(w["draft_len"]) + 1) // 2
                for a in arms[int(w["first_arm"]): int(w["first_arm"]) + int(w["n_internal"]) + int(w["n_pre"]) + int(w["n_suf"])]:
                    assert int(a["off"]) == pos
                    pos += (int(a["len"]) + 3) // 4
            assert pos == packed.size
            blob = win.tobytes() + arms.tobytes() + packed.tobytes()
            assert ref is None or blob == ref
            ref = blob
    # a batch already in container order comes back byte for byte
    b = synth_batch(11, 300, 120, 30, "internal")
    win, arms, packed, _ = host_pack(b, 4)
    assert win.tobytes() == b.win.tobytes() and arms.tobytes() == b.arms.tobytes()
    assert packed.tobytes() == b.packed[:packed.size].tobytes()


def test_long_window_arm_filter_matches_reference_golden():
    """Window::use_reference_long_filter: the mirror's add_* keeps exactly the arms the reference's Window
    keeps for LONG windows (reference include/Window.hpp:66-101, include/Filter.hpp).  Golden flags from the
    compiled reference (tests/golden/make_golden.py filter), both outcomes well represented."""
    from hypo_b200.batch import WindowSpec, build_batch
    from hypo_b200.hostlib import long_arm_filter
    from tests.conftest import load_golden
    g = load_golden("long_filter.json.gz")
    specs = [WindowSpec(w["draft"], w["internal"], w["pre"], w["suf"], 0, 1) for w in g["windows"]]
    want = np.array([x for w in g["windows"] for x in w["accepted"]], dtype=bool)
    assert (~want).sum() > 100 and want.sum() > 100
    got = long_arm_filter(build_batch(specs))
    assert (got == want).all(), f"{int((got != want).sum())} of {want.size} arms decided differently"
    # SHORT windows are never filtered
    short = [WindowSpec(w["draft"], w["internal"], w["pre"], w["suf"], 0, 0) for w in g["windows"][:8]]
    assert long_arm_filter(build_batch(short)).all()


def test_long_window_arm_filter_vs_compiled_reference():
    from hypo_b200.hostlib import long_arm_filter
    from tests.oracle_util import ref_consensus, ref_lib
    if ref_lib() is None:
        import pytest
        pytest.skip("oracle/_ref not built (no /root/reference)")
    for seed, kw in ((21, dict(n_win=60, length=250, n_arms=20, kind="internal", err=0.08, wtype=1)),
                     (22, dict(n_win=60, length=90, n_arms=12, kind="mixed", err=0.04, wtype=1)),
                     (23, dict(n_win=40, length=30, n_arms=8, kind="prefix", err=0.02, wtype=1))):
        b = synth_batch(seed, **kw)
        _, acc, _ = ref_consensus(b)
        assert (long_arm_filter(b) == acc).all(), kw


def _packedseq_probe(lib, name, hts, seq_len, offset, left, right):
    import ctypes as C
    fn = getattr(lib, name)
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32] + [C.c_char_p] * 6
    bufs = [C.create_string_buffer(max(n, 1)) for n in (seq_len, seq_len) + (right - left,) * 4]
    flags = fn(hts.ctypes.data, seq_len, offset, left, right, *bufs)
    out = [b.raw[:n] for b, n in zip(bufs, (seq_len, seq_len) + (right - left,) * 4)]
    if not flags & 1:
        out[0] = out[2] = b""
    if not flags & 2:
        out[3] = b""
    return flags, out


def test_packedseq_mirror_vs_compiled_reference():
    """Data formats on the input side of the path: a read as htslib stores it -> PackedSeq<2>/<4>, and
    sub-ranges of packed sequences (how hypo::Alignment cuts arms and hypo::Contig cuts window drafts;
    reference src/PackedSeq.cpp:122-229).  The mirror must answer like the reference's own class."""
    import pytest
    from hypo_b200 import hostlib
    from tests.oracle_util import ref_lib
    ref = ref_lib()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference)")
    rng = np.random.default_rng(77)
    codes = np.array([1, 2, 4, 8], np.uint8)   # A C G T in htslib's 4-bit alphabet
    seen_invalid = seen_dirty = 0
    for trial in range(300):
        total = int(rng.integers(1, 90))
        nib = codes[rng.integers(0, 4, size=total)]
        if trial % 3 == 0:   # sprinkle N (15) and other ambiguity codes
            k = rng.integers(0, total, size=max(1, total // 15))
            nib[k] = rng.choice(np.array([15, 3, 5, 0], np.uint8), size=k.size)
        if total & 1:
            nib = np.append(nib, np.uint8(0))
        hts = ((nib[0::2] << 4) | nib[1::2]).astype(np.uint8)
        offset = int(rng.integers(0, total))
        seq_len = int(rng.integers(1, total - offset + 1))
        left = int(rng.integers(0, seq_len))
        right = int(rng.integers(left + 1, seq_len + 1))   # (the reference cannot build an empty sub-range: bad_alloc)
        want = _packedseq_probe(ref, "hypo_ref_packedseq_probe", hts, seq_len, offset, left, right)
        got = _packedseq_probe(hostlib.lib(), "hypo_host_packedseq_probe", hts, seq_len, offset, left, right)
        assert got == want, (trial, seq_len, offset, left, right)
        seen_invalid += not want[0] & 1
        seen_dirty += not want[0] & 2
    assert seen_invalid > 10 and seen_dirty > 5


def test_packedseq_kmer_search_vs_compiled_reference():
    """find_kmer / check_kmer / find_canonical_kmer / check_canonical_kmer of the mirror against the
    reference's class (reference src/PackedSeq.cpp:264-414): first and last occurrence, N restarts the
    k-mer, canonical k-mers in 32-bit registers."""
    import ctypes as C
    import pytest
    from hypo_b200 import hostlib
    from tests.oracle_util import ref_lib
    ref = ref_lib()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference)")

    def probe(lib, name, *a):
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                       C.POINTER(C.c_uint64)]
        r = C.c_uint64(0)
        return fn(*a, C.byref(r)), int(r.value)

    def kmer(s, canonical):
        f = 0
        for ch in s:
            f = f << 2 | "ACGT".index(ch)
        if not canonical:
            return f
        r = 0
        for ch in reversed(s):
            r = r << 2 | 3 - "ACGT".index(ch)
        return min(f, r)

    rng = np.random.default_rng(5)
    hits = 0
    for trial in range(600):
        n = int(rng.integers(8, 120))
        nb = 2 if trial % 2 else 4
        seq = "".join(rng.choice(list("ACGT"), size=n))
        if trial % 5 == 0:   # low complexity: repeated k-mers, first != last occurrence
            seq = (seq[:6] * 30)[:n]
        if nb == 4 and trial % 3 == 0:
            seq = list(seq)
            for p in rng.integers(0, n, size=max(1, n // 12)):
                seq[p] = "N"
            seq = "".join(seq)
        k = int(rng.integers(3, min(16, n) + 1))
        mode = int(rng.integers(0, 4))
        left = int(rng.integers(0, n - k + 1))
        right = int(rng.integers(left, n + 1)) if mode in (0, 2) else left + k
        src = int(rng.integers(0, n - k + 1))
        word = seq[src:src + k]
        if "N" in word or trial % 4 == 0:
            target = int(rng.integers(0, 4 ** k))
        else:
            target = kmer(word, mode >= 2)
        args = (seq.encode(), n, nb, mode, target, k, left, right, int(trial % 2))
        want = probe(ref, "hypo_ref_kmer_probe", *args)
        got = probe(hostlib.lib(), "hypo_host_kmer_probe", *args)
        assert got == want, (trial, seq, args)
        hits += want[0]
    assert 60 < hits < 540


def test_window_mirror_counters_vs_compiled_reference():
    """add_* / add_empty / get_num_* / get_maxlen_* / get_window_len / clear_pre_suf of the Window mirror
    against the reference's Window (reference include/Window.hpp:61-120), LONG windows filtered."""
    import ctypes as C
    import pytest
    from hypo_b200 import hostlib
    from hypo_b200.batch import build_batch
    from hypo_b200.synth import edge_case_windows
    from tests.oracle_util import ref_lib
    ref = ref_lib()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference)")
    for b in (build_batch(edge_case_windows()), synth_batch(31, 80, 70, 14, "mixed", err=0.03),
              synth_batch(32, 60, 200, 16, "mixed", err=0.06, wtype=1), synth_batch(33, 40, 90, 10, "prefix", err=0.02, wtype=1)):
        res = []
        for lib, name in ((ref, "hypo_ref_window_counts"), (hostlib.lib(), "hypo_host_window_counts")):
            fn = getattr(lib, name)
            fn.restype = None
            fn.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
            c = np.zeros(8 * b.n_win, np.uint32)
            fn(b.win.ctypes.data, b.n_win, b.arms.ctypes.data, b.packed.ctypes.data, c.ctypes.data)
            res.append(c.reshape(-1, 8))
        assert (res[0] == res[1]).all()
