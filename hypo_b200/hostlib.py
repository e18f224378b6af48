"""ctypes binding of hypo_b200/libhypo_host.so: the C++ host side of the drop-in (Window /
PackedSeq mirror + batch packer) and the fast seeded synthetic window generator."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence

import numpy as np

from .batch import ARM_DTYPE, WIN_DTYPE, WindowBatch, split_consensus
from .native import HypoGpuError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhypo_host.so")
KINDS = {"internal": 0, "backbone": 1, "prefix": 2, "suffix": 3, "mixed": 4}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HypoGpuError(-1, f"{LIB_PATH} is missing - run __graft_entry__.build()")
        L = C.CDLL(LIB_PATH)
        L.hypo_synth_packed_bound.restype = C.c_uint64
        L.hypo_synth_packed_bound.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        L.hypo_synth_generate.restype = C.c_int
        L.hypo_synth_generate.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_double,
                                          C.c_double, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                          C.POINTER(C.c_uint64), C.c_int]
        L.hypo_host_pack.restype = C.c_int
        L.hypo_host_pack.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.hypo_host_long_filter.restype = C.c_int
        L.hypo_host_long_filter.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hypo_host_run.restype = C.c_int
        L.hypo_host_run.argtypes = [C.POINTER(C.c_int8), C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_uint64, C.c_void_p]
        L.hypo_host_windows_create.restype = C.c_void_p
        L.hypo_host_windows_create.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.hypo_host_windows_free.restype = None
        L.hypo_host_windows_free.argtypes = [C.c_void_p]
        L.hypo_host_windows_run.restype = C.c_int
        L.hypo_host_windows_run.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_double)]
        L.hypo_host_windows_consensus.restype = C.c_uint64
        L.hypo_host_windows_consensus.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.hypo_host_inspect_write.restype = C.c_int
        L.hypo_host_inspect_write.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p]
        L.hypo_host_inspect_open.restype = C.c_void_p
        L.hypo_host_inspect_open.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64]
        L.hypo_host_inspect_close.restype = None
        L.hypo_host_inspect_close.argtypes = [C.c_void_p]
        L.hypo_host_inspect_sizes.restype = None
        L.hypo_host_inspect_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.hypo_host_inspect_fill.restype = None
        L.hypo_host_inspect_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hypo_host_inspect_stitch.restype = C.c_uint64
        L.hypo_host_inspect_stitch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
        L.hypo_host_inspect_replay.restype = C.c_int64
        L.hypo_host_inspect_replay.argtypes = [C.c_void_p, C.POINTER(C.c_int8), C.c_int, C.POINTER(C.c_double),
                                               C.c_char_p]
        _lib = L
    return _lib


def synth_batch(seed: int, n_win: int, length: int = 120, n_arms: int = 30, kind: str = "internal",
                err: float = 0.01, draft_err: float = 0.03, wtype: int = 0, threads: int = 0) -> WindowBatch:
    """Seeded synthetic windows: truth iid ACGT; draft = truth with draft_err
    sub/ins/del each; arms = (slices of) truth with err sub/ins/del each."""
    L = lib()
    cap = int(L.hypo_synth_packed_bound(n_win, length, n_arms))
    win = np.zeros(n_win, WIN_DTYPE)
    arms = np.zeros(n_win * n_arms, ARM_DTYPE)
    packed = np.zeros(cap, np.uint8)
    used = C.c_uint64(0)
    rc = L.hypo_synth_generate(seed, n_win, length, n_arms, KINDS[kind], err, draft_err, wtype, win.ctypes.data,
                               arms.ctypes.data, packed.ctypes.data, cap, C.byref(used), threads)
    if rc != 0:
        raise RuntimeError("hypo_synth_generate: packed buffer too small")
    packed = packed[: int(used.value) + 16].copy()   # 16 bytes of slack, zero-filled
    return WindowBatch(win, arms, packed,
                       {"seed": seed, "length": length, "n_arms": n_arms, "kind": kind, "err": err,
                        "draft_err": draft_err, "wtype": wtype})


def host_run(batch: WindowBatch, scores: Sequence[int] = (5, -4, -8, 3, -5, -4), device: int = 0) -> List[str]:
    """Builds hypo::Window objects through the mirror's public API and runs
    Window::generate_consensus_batch (the drop-in for reference src/Hypo.cpp:236-248)."""
    L = lib()
    cap = int(batch.out_bound().sum()) + 16
    out = np.empty(cap, np.uint8)
    off = np.zeros(batch.n_win + 1, np.uint64)
    sc = (C.c_int8 * 6)(*[int(x) for x in scores])
    rc = L.hypo_host_run(sc, device, batch.win.ctypes.data, batch.n_win, batch.arms.ctypes.data,
                         batch.packed.ctypes.data, out.ctypes.data, cap, off.ctypes.data)
    if rc != 0:
        raise HypoGpuError(rc, "hypo_host_run failed")
    return split_consensus(out, off)


class HostWindows:
    """hypo::Window objects (built through the mirror's public add_* API) that are polished in place by
    WindowBatch::run - pack, copies, kernels and scatter, chunked and double-buffered: the call a
    maintainer makes instead of the OpenMP loop of reference src/Hypo.cpp:236-248."""

    def __init__(self, batch: WindowBatch):
        self.n_win = batch.n_win
        self._h = lib().hypo_host_windows_create(batch.win.ctypes.data, batch.n_win, batch.arms.ctypes.data,
                                                 batch.packed.ctypes.data)

    def run(self, chunk_windows: int = 0) -> dict:
        t = (C.c_double * 5)()
        rc = lib().hypo_host_windows_run(self._h, int(chunk_windows), t)
        if rc != 0:
            raise HypoGpuError(rc, "hypo_host_windows_run failed")
        return {"pack_s": t[0], "device_s": t[1], "scatter_s": t[2], "total_s": t[3], "chunks": int(t[4])}

    def consensus_bytes(self):
        n = int(lib().hypo_host_windows_consensus(self._h, None, 0, None))
        out = np.empty(n + 1, np.uint8)
        off = np.zeros(self.n_win + 1, np.uint64)
        lib().hypo_host_windows_consensus(self._h, out.ctypes.data, n, off.ctypes.data)
        return out[:n], off

    def consensus(self) -> List[str]:
        out, off = self.consensus_bytes()
        return split_consensus(out, off)

    def close(self):
        if self._h:
            lib().hypo_host_windows_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def long_arm_filter(batch: WindowBatch) -> np.ndarray:
    """CPU only: per arm, whether the Window mirror keeps it when it filters the arms of LONG windows like
    the reference's Window (Window::use_reference_long_filter)."""
    acc = np.zeros(max(batch.n_arms, 1), np.uint8)
    rc = lib().hypo_host_long_filter(batch.win.ctypes.data, batch.n_win, batch.arms.ctypes.data,
                                     batch.packed.ctypes.data, acc.ctypes.data)
    if rc != 0:
        raise HypoGpuError(rc, "hypo_host_long_filter: Window::add_* and the filter disagree")
    return acc[: batch.n_arms].astype(bool)


def host_pack(batch: WindowBatch, threads: int = 0):
    """CPU only: batch -> hypo::Window objects (public add_* API) -> the C++ batch packer with `threads`
    OpenMP threads -> (win, arms, packed, seconds spent packing).  The packer lays the slab out in
    container order (draft, internal, prefix, suffix arms, window after window); for a batch that already
    is, the three arrays equal the batch's own byte for byte."""
    L = lib()
    # bytes in use (batches may carry slack behind the last sequence)
    n_bytes = int(max((batch.arms["off"] + (batch.arms["len"] + 3) // 4).max(initial=0),
                      (batch.win["draft_off"] + (batch.win["draft_len"] + 1) // 2).max(initial=0)))
    win = np.zeros(batch.n_win, WIN_DTYPE)
    arms = np.zeros(batch.n_arms, ARM_DTYPE)
    packed = np.zeros(n_bytes, np.uint8)
    sec = C.c_double(0.0)
    rc = L.hypo_host_pack(batch.win.ctypes.data, batch.n_win, batch.arms.ctypes.data, batch.n_arms,
                          batch.packed.ctypes.data, n_bytes, threads, win.ctypes.data, arms.ctypes.data,
                          packed.ctypes.data, C.byref(sec))
    if rc != 0:
        raise HypoGpuError(rc, "hypo_host_pack: the packer's arm / byte counts differ from the batch's")
    return win, arms, packed, float(sec.value)


# ---- window streams in the reference's inspect-file format (host/WindowStream.hpp) -------------------

def write_inspect(path: str, batch: WindowBatch, consensus: Sequence[str], contig: str = "ctg") -> None:
    """Stores a batch and the consensus strings to record in the format of the reference's
    Contig::generate_inspect_file (reference src/Contig.cpp:368-453, src/Window.cpp:63-84)."""
    blob = "".join(consensus).encode()
    off = np.zeros(batch.n_win + 1, np.uint64)
    off[1:] = np.cumsum([len(c) for c in consensus])
    buf = np.frombuffer(blob + b"\0", np.uint8).copy()
    rc = lib().hypo_host_inspect_write(path.encode(), contig.encode(), batch.win.ctypes.data, batch.n_win,
                                       batch.arms.ctypes.data, batch.packed.ctypes.data, buf.ctypes.data,
                                       off.ctypes.data)
    if rc != 0:
        raise OSError(f"cannot write {path}")


class InspectStream:
    """A parsed inspect file: its windows as a flat batch, the recorded consensus strings, and a
    replay through the Window mirror's public API (one C-ABI call on the device)."""

    def __init__(self, path: str):
        err = C.create_string_buffer(512)
        self._h = lib().hypo_host_inspect_open(path.encode(), err, 512)
        if not self._h:
            raise ValueError(err.value.decode())
        counts = (C.c_uint64 * 6)()
        lib().hypo_host_inspect_sizes(self._h, counts)
        self.n_regions, self.n_windows, n_arms, n_bytes, n_cons, self.polished_bp = [int(x) for x in counts]
        win = np.zeros(self.n_windows, WIN_DTYPE)
        arms = np.zeros(n_arms, ARM_DTYPE)
        packed = np.zeros(n_bytes + 16, np.uint8)
        cons = np.zeros(n_cons + 1, np.uint8)
        off = np.zeros(self.n_windows + 1, np.uint64)
        lib().hypo_host_inspect_fill(self._h, win.ctypes.data, arms.ctypes.data, packed.ctypes.data,
                                     cons.ctypes.data, off.ctypes.data)
        self.batch = WindowBatch(win, arms, packed, {"source": path})
        self.recorded = split_consensus(cons, off)

    def replay(self, scores: Sequence[int] = (5, -4, -8, 3, -5, -4), device: int = 0, out_path: str = ""):
        """Returns (windows whose consensus differs from the recorded one, seconds of the batch call)."""
        sec = C.c_double(0)
        sc = (C.c_int8 * 6)(*[int(x) for x in scores])
        bad = lib().hypo_host_inspect_replay(self._h, sc, device, C.byref(sec), out_path.encode())
        return int(bad), float(sec.value)

    def stitched(self, mode: int = 0) -> str:
        """The polished contig (Contig::operator<<) from the windows' current consensus strings: mode 0 on
        the host, 1 on the device (hypo_gpu_stitch), 2 on the device from the still resident result of
        the replay that just ran, 3 on the host from the RECORDED consensus strings."""
        n = int(lib().hypo_host_inspect_stitch(self._h, mode, None, 0)) if mode in (0, 3) else self.n_bases_bound()
        buf = C.create_string_buffer(n + 16)
        n = int(lib().hypo_host_inspect_stitch(self._h, mode, buf, n + 16))
        return buf.raw[:n].decode()

    def n_bases_bound(self) -> int:
        return int(lib().hypo_host_inspect_stitch(self._h, 3, None, 0)) * 2 + 1024

    def close(self):
        if self._h:
            lib().hypo_host_inspect_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
