"""Window batches in the layout of include/hypo_b200.h (host side, numpy).

A batch is three flat buffers — window descriptors, arm descriptors and the packed
sequence slab — exactly what crosses the C ABI.  The packing rules are those of the
reference's PackedSeq (reference src/PackedSeq.cpp:45-48,58-88):
  * arms   : PackedSeq<2>, 4 bases/byte, base i in bits 6-2*(i&3) of byte i>>2, A0 C1 G2 T3
  * drafts : PackedSeq<4>, 2 bases/byte, even base in the high nibble, code 4 = N
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterable, List, Optional, Sequence

import numpy as np

WINDOW_SHORT = 0
WINDOW_LONG = 1

WIN_DTYPE = np.dtype(
    [
        ("draft_off", "<u8"),
        ("first_arm", "<u8"),
        ("draft_len", "<u4"),
        ("n_internal", "<u4"),
        ("n_pre", "<u4"),
        ("n_suf", "<u4"),
        ("n_empty", "<u4"),
        ("wtype", "<u4"),
    ],
    align=False,
)
ARM_DTYPE = np.dtype([("off", "<u8"), ("len", "<u4"), ("reserved", "<u4")], align=False)
assert WIN_DTYPE.itemsize == 40 and ARM_DTYPE.itemsize == 16

_NT4 = np.full(256, 4, dtype=np.uint8)
for _c, _v in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("U", 3)):
    _NT4[ord(_c)] = _v
    _NT4[ord(_c.lower())] = _v


def pack2(seq: str | bytes) -> np.ndarray:
    """PackedSeq<2>(std::string) — reference src/PackedSeq.cpp:58-88."""
    b = np.frombuffer(seq.encode() if isinstance(seq, str) else seq, dtype=np.uint8)
    codes = _NT4[b]
    if codes.size and codes.max() > 3:
        raise ValueError("[Hypo::PackedSeq] Error: Wrong base (Can not pack in 2 bits)")
    n = codes.size
    pad = (-n) % 4
    c = np.concatenate([codes, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    return (c[:, 0] << 6 | c[:, 1] << 4 | c[:, 2] << 2 | c[:, 3]).astype(np.uint8)


def pack4(seq: str | bytes) -> np.ndarray:
    """PackedSeq<4>(std::string) — reference src/PackedSeq.cpp:58-88 with NB=4."""
    b = np.frombuffer(seq.encode() if isinstance(seq, str) else seq, dtype=np.uint8)
    codes = _NT4[b]
    n = codes.size
    pad = (-n) % 2
    c = np.concatenate([codes, np.zeros(pad, np.uint8)]).reshape(-1, 2)
    return (c[:, 0] << 4 | c[:, 1]).astype(np.uint8)


def unpack2(buf: np.ndarray, off: int, n: int) -> str:
    i = np.arange(n)
    v = (buf[off + (i >> 2)] >> (6 - 2 * (i & 3))) & 3
    return "".join("ACGT"[x] for x in v)


def unpack4(buf: np.ndarray, off: int, n: int) -> str:
    i = np.arange(n)
    v = (buf[off + (i >> 1)] >> np.where(i & 1, 0, 4)) & 15
    return "".join("ACGTNNNNNNNNNNNN"[x] for x in v)


@dataclass
class WindowSpec:
    """One hypo::Window as the reference holds it (reference include/Window.hpp:122-135)."""

    draft: str
    internal: Sequence[str] = ()
    pre: Sequence[str] = ()
    suf: Sequence[str] = ()
    n_empty: int = 0
    wtype: int = WINDOW_SHORT


@dataclass
class WindowBatch:
    win: np.ndarray  # WIN_DTYPE [n_win]
    arms: np.ndarray  # ARM_DTYPE [n_arms]
    packed: np.ndarray  # uint8 [packed_bytes]
    meta: dict = field(default_factory=dict)

    @property
    def n_win(self) -> int:
        return int(self.win.shape[0])

    @property
    def n_arms(self) -> int:
        return int(self.arms.shape[0])

    @property
    def polished_bp(self) -> int:
        """Metric numerator: sum of Window::get_window_len() (reference include/Window.hpp:64)."""
        return int(self.win["draft_len"].astype(np.int64).sum())

    def out_bound(self) -> np.ndarray:
        """Per-window upper bound of the bytes a window may write (same rule as hypo_gpu_window_bounds):
        SHORT sum(len) + 2 * arms + draft + 2; LONG 2 * sum(len) + draft + 2."""
        n_per = (self.win["n_internal"] + self.win["n_pre"] + self.win["n_suf"]).astype(np.int64)
        first = self.win["first_arm"].astype(np.int64)
        csum = np.concatenate([[0], np.cumsum(self.arms["len"].astype(np.int64))])
        s = csum[first + n_per] - csum[first]
        draft = self.win["draft_len"].astype(np.int64)
        return np.where(self.win["wtype"] == WINDOW_LONG, 2 * s + draft + 2, s + 2 * n_per + draft + 2)

    def algorithmic_bytes(self, consensus_bytes: int) -> int:
        """Compulsory HBM traffic of one pass: every input byte read
        once, every output byte written once."""
        return int(
            self.win.nbytes + self.arms.nbytes + self.packed.nbytes + consensus_bytes + 12 * self.n_win
        )

    def select(self, idx: np.ndarray) -> "WindowBatch":
        """Sub-batch with the given windows (used for sharding and sampling); the packed
        slab is shared, arm descriptors are re-indexed."""
        idx = np.asarray(idx, dtype=np.int64)
        win = self.win[idx].copy()
        n_per = (win["n_internal"] + win["n_pre"] + win["n_suf"]).astype(np.int64)
        first = win["first_arm"].astype(np.int64)
        new_first = np.concatenate([[0], np.cumsum(n_per)[:-1]]) if len(idx) else np.zeros(0, np.int64)
        total = int(n_per.sum())
        gather = np.repeat(first - new_first, n_per) + np.arange(total, dtype=np.int64)
        arms = self.arms[gather].copy()
        win["first_arm"] = new_first
        return WindowBatch(win, arms, self.packed, dict(self.meta))

    def compact(self) -> "WindowBatch":
        """Re-bases the descriptors onto the smallest slice of the packed slab that holds every byte they
        reference (a contiguous range of a window-ordered batch then carries only its own bytes)."""
        if self.n_win == 0:
            return self
        a_end = self.arms["off"].astype(np.int64) + (self.arms["len"].astype(np.int64) + 3) // 4
        d_end = self.win["draft_off"].astype(np.int64) + (self.win["draft_len"].astype(np.int64) + 1) // 2
        used = self.arms["len"] > 0
        lo = int(min(self.win["draft_off"].min(), self.arms["off"][used].min() if used.any() else 1 << 62))
        hi = int(max(d_end.max(), a_end[used].max() if used.any() else 0))
        win, arms = self.win.copy(), self.arms.copy()
        win["draft_off"] -= lo
        arms["off"][used] -= lo
        arms["off"][~used] = 0
        packed = np.concatenate([self.packed[lo:hi], np.zeros(16, np.uint8)])
        return WindowBatch(win, arms, packed, dict(self.meta))

    def spec(self, w: int) -> WindowSpec:
        d = self.win[w]
        a0 = int(d["first_arm"])
        ni, npre, nsuf = int(d["n_internal"]), int(d["n_pre"]), int(d["n_suf"])

        def arm(k):
            return unpack2(self.packed, int(self.arms[k]["off"]), int(self.arms[k]["len"]))

        return WindowSpec(
            draft=unpack4(self.packed, int(d["draft_off"]), int(d["draft_len"])),
            internal=[arm(a0 + i) for i in range(ni)],
            pre=[arm(a0 + ni + i) for i in range(npre)],
            suf=[arm(a0 + ni + npre + i) for i in range(nsuf)],
            n_empty=int(d["n_empty"]),
            wtype=int(d["wtype"]),
        )


def build_batch(specs: Iterable[WindowSpec]) -> WindowBatch:
    """Flatten WindowSpecs the way the host packer walks hypo::Window objects
    (container order: _internal_arms, _pre_arms, _suf_arms; reference include/Window.hpp:131-133)."""
    specs = list(specs)
    win = np.zeros(len(specs), dtype=WIN_DTYPE)
    arm_rows: List[tuple] = []
    chunks: List[np.ndarray] = []
    pos = 0

    def put(a: np.ndarray) -> int:
        nonlocal pos
        off = pos
        chunks.append(a)
        pos += a.size
        return off

    for w, s in enumerate(specs):
        win[w]["draft_off"] = put(pack4(s.draft))
        win[w]["draft_len"] = len(s.draft)
        win[w]["first_arm"] = len(arm_rows)
        win[w]["n_internal"] = len(s.internal)
        win[w]["n_pre"] = len(s.pre)
        win[w]["n_suf"] = len(s.suf)
        win[w]["n_empty"] = s.n_empty
        win[w]["wtype"] = s.wtype
        for seq in list(s.internal) + list(s.pre) + list(s.suf):
            arm_rows.append((put(pack2(seq)), len(seq), 0))
    arms = np.array(arm_rows, dtype=ARM_DTYPE) if arm_rows else np.zeros(0, dtype=ARM_DTYPE)
    packed = np.concatenate(chunks).astype(np.uint8) if chunks else np.zeros(0, np.uint8)
    # 16 bytes of slack so vectorised device loads may over-read the last arm safely
    packed = np.concatenate([packed, np.zeros(16, np.uint8)])
    return WindowBatch(win, arms, packed)


def concat_batches(batches: Sequence[WindowBatch], meta: Optional[dict] = None) -> WindowBatch:
    """One batch holding the windows of all `batches` in order (descriptors re-based onto one
    arm table and one packed slab)."""
    wins, arms, packed = [], [], []
    arm_base, byte_base = 0, 0
    for b in batches:
        w = b.win.copy()
        a = b.arms.copy()
        w["first_arm"] += arm_base
        w["draft_off"] += byte_base
        a["off"] += byte_base
        wins.append(w); arms.append(a); packed.append(b.packed)
        arm_base += b.n_arms
        byte_base += int(b.packed.size)
    return WindowBatch(np.concatenate(wins), np.concatenate(arms), np.concatenate(packed), dict(meta or {}))


def split_consensus(out: np.ndarray, out_off: np.ndarray) -> List[str]:
    raw = out.tobytes()
    return [raw[int(out_off[i]) : int(out_off[i + 1])].decode("ascii") for i in range(len(out_off) - 1)]
