"""ctypes binding of the C ABI in include/hypo_b200.h (hypo_b200/libhypo_b200.so).

This is the only way the Python host layer reaches the device.  There is no CPU fallback:
if the shared library is missing or no CUDA device is usable, every call raises."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .batch import WindowBatch, split_consensus

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhypo_b200.so")

ABI_SYMBOLS = (
    "hypo_gpu_init",
    "hypo_gpu_init_multi",
    "hypo_gpu_device_count",
    "hypo_gpu_set_option",
    "hypo_gpu_window_bounds",
    "hypo_gpu_out_bound",
    "hypo_gpu_consensus_batch",
    "hypo_gpu_consensus_batch_device",
    "hypo_gpu_last_fail_hist",
    "hypo_gpu_compact_device",
    "hypo_gpu_last_timing",
    "hypo_gpu_last_tier_windows",
    "hypo_gpu_extract_arms",
    "hypo_gpu_polish_alignments",
    "hypo_gpu_solid_kmer_support",
    "hypo_gpu_minimiser_support",
    "hypo_gpu_stitch",
    "hypo_gpu_last_rerouted",
    "hypo_gpu_last_cells",
    "hypo_gpu_issue_rate",
    "hypo_gpu_launch_count",
    "hypo_gpu_last_error",
    "hypo_gpu_host_alloc",
    "hypo_gpu_host_free",
    "hypo_gpu_shutdown",
    "hypo_gpu_abi_version",
)


N_TIERS = 11          # capacity tiers of the library (hypo_gpu_last_tier_windows)
TIER_QUAD, TIER_HALF, TIER_BIG = 8, 9, 10


class HypoGpuError(RuntimeError):
    """Raised for every non-zero return of the C ABI; mirrors the reference's
    `fprintf(stderr, "[Hypo::X] Error: ...")` + exit(1) convention with an exception."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[Hypo::GPU] Error: {msg} (code {code})")
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HypoGpuError(
                -1,
                f"{LIB_PATH} is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a); there is no CPU fallback",
            )
        L = C.CDLL(LIB_PATH)
        L.hypo_gpu_init.restype = C.c_int
        L.hypo_gpu_init.argtypes = [C.POINTER(C.c_int8), C.c_int]
        L.hypo_gpu_init_multi.restype = C.c_int
        L.hypo_gpu_init_multi.argtypes = [C.POINTER(C.c_int8), C.c_int]
        L.hypo_gpu_device_count.restype = C.c_int
        L.hypo_gpu_set_option.restype = C.c_int
        L.hypo_gpu_set_option.argtypes = [C.c_char_p, C.c_int64]
        L.hypo_gpu_window_bounds.restype = C.c_int
        L.hypo_gpu_window_bounds.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
        L.hypo_gpu_out_bound.restype = C.c_uint64
        L.hypo_gpu_out_bound.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.hypo_gpu_consensus_batch.restype = C.c_int
        L.hypo_gpu_consensus_batch.argtypes = [
            C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
            C.c_void_p, C.c_uint64, C.c_void_p]
        L.hypo_gpu_consensus_batch_device.restype = C.c_int
        L.hypo_gpu_consensus_batch_device.argtypes = [
            C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hypo_gpu_compact_device.restype = C.c_int
        L.hypo_gpu_compact_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                              C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]
        L.hypo_gpu_last_timing.restype = C.c_int
        L.hypo_gpu_last_timing.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.hypo_gpu_last_tier_windows.restype = C.c_int
        L.hypo_gpu_last_tier_windows.argtypes = [C.POINTER(C.c_uint32), C.c_int]
        L.hypo_gpu_last_rerouted.restype = C.c_uint64
        L.hypo_gpu_last_cells.restype = C.c_uint64
        L.hypo_gpu_issue_rate.restype = C.c_int
        L.hypo_gpu_issue_rate.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.hypo_gpu_launch_count.restype = C.c_uint64
        L.hypo_gpu_last_error.restype = C.c_char_p
        L.hypo_gpu_shutdown.restype = None
        L.hypo_gpu_abi_version.restype = C.c_int
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise HypoGpuError(rc, lib().hypo_gpu_last_error().decode(errors="replace"))


def init(scores: Sequence[int] = (5, -4, -8, 3, -5, -4), device: int = 0) -> None:
    """Window::prepare_for_poa (reference src/Window.cpp:31-42): fix the score parameters."""
    sc = (C.c_int8 * 6)(*[int(x) for x in scores])
    _check(lib().hypo_gpu_init(sc, int(device)))


def init_multi(scores: Sequence[int] = (5, -4, -8, 3, -5, -4), n_gpus: int = 0) -> None:
    """One host process driving n_gpus devices (0 = all visible): batches are cut into contiguous
    window ranges of equal estimated cost, one per device."""
    sc = (C.c_int8 * 6)(*[int(x) for x in scores])
    _check(lib().hypo_gpu_init_multi(sc, int(n_gpus)))


def device_count() -> int:
    return int(lib().hypo_gpu_device_count())


def set_option(name: str, value: int) -> None:
    """Test / measurement knobs of the C ABI (first_tier, scap, gather)."""
    _check(lib().hypo_gpu_set_option(name.encode(), int(value)))


def window_bounds(batch: WindowBatch) -> np.ndarray:
    """hypo_gpu_window_bounds: bytes every window's output slot must hold (host arithmetic)."""
    b = np.zeros(batch.n_win, np.uint64)
    _check(lib().hypo_gpu_window_bounds(batch.win.ctypes.data, batch.n_win, batch.arms.ctypes.data, batch.n_arms,
                                        b.ctypes.data))
    return b


REGION_DTYPE = np.dtype([("key0", "<u8"), ("key1", "<u8"), ("start", "<u4"), ("type", "<u4")])
CONTIG_DTYPE = np.dtype([("first_region", "<u8"), ("draft_off", "<u8"), ("n_regions", "<u4"), ("len", "<u4")])
ALN_DTYPE = np.dtype([("cigar_off", "<u8"), ("seq_off", "<u8"), ("contig", "<u4"), ("pos", "<u4"), ("n_cigar", "<u4"),
                      ("l_qseq", "<u4")])
REGION_TYPES = {"SR": 0, "MSR": 1, "SWS": 2, "WS": 3, "SW": 4, "MWM": 5, "WM": 6, "MW": 7, "SWM": 8, "MWS": 9, "OTH": 10}
assert REGION_DTYPE.itemsize == 24 and CONTIG_DTYPE.itemsize == 24 and ALN_DTYPE.itemsize == 32


def extract_arms(contigs, regions, drafts, alns, cigar, seqs, k: int):
    """hypo_gpu_extract_arms: alignments + region tables -> (WindowBatch, win_region)."""
    from .batch import ARM_DTYPE, WIN_DTYPE
    L = lib()
    L.hypo_gpu_extract_arms.restype = C.c_int
    sizes = (C.c_uint64 * 3)()
    win_cap, arm_cap, byte_cap = len(regions) + 1, int(max(1, cigar.size * 0 + 1)), 1
    # first call with empty buffers reports the sizes
    for attempt in range(2):
        win = np.zeros(win_cap, WIN_DTYPE)
        win_region = np.zeros(win_cap, np.uint64)
        arms = np.zeros(arm_cap, ARM_DTYPE)
        packed = np.zeros(byte_cap + 16, np.uint8)
        rc = L.hypo_gpu_extract_arms(
            C.c_void_p(contigs.ctypes.data), C.c_uint64(len(contigs)), C.c_void_p(regions.ctypes.data),
            C.c_uint64(len(regions)), C.c_void_p(drafts.ctypes.data), C.c_uint64(drafts.size),
            C.c_void_p(alns.ctypes.data), C.c_uint64(len(alns)), C.c_void_p(cigar.ctypes.data), C.c_uint64(cigar.size),
            C.c_void_p(seqs.ctypes.data), C.c_uint64(seqs.size), C.c_uint32(k),
            C.c_void_p(win.ctypes.data), C.c_uint64(win_cap), C.byref(sizes, 0), C.c_void_p(win_region.ctypes.data),
            C.c_void_p(arms.ctypes.data), C.c_uint64(arm_cap), C.byref(sizes, 8),
            C.c_void_p(packed.ctypes.data), C.c_uint64(byte_cap), C.byref(sizes, 16))
        n_win, n_arms, n_bytes = int(sizes[0]), int(sizes[1]), int(sizes[2])
        if rc == 4 and attempt == 0:   # HYPO_E_OUT_CAP: now the sizes are known
            win_cap, arm_cap, byte_cap = max(n_win, 1), max(n_arms, 1), max(n_bytes, 1)
            continue
        _check(rc)
        break
    return WindowBatch(win[:n_win].copy(), arms[:n_arms].copy(), packed[: n_bytes + 16].copy(), {}), win_region[:n_win].copy()


def solid_kmer_support(contig_first_kmer, solid_pos, kmer_id, alns, cigar, seqs, k: int):
    """hypo_gpu_solid_kmer_support -> (coverage, support) per solid k-mer."""
    L = lib()
    L.hypo_gpu_solid_kmer_support.restype = C.c_int
    n = int(solid_pos.size)
    cov, sup = np.zeros(n + 1, np.uint32), np.zeros(n + 1, np.uint32)
    _check(L.hypo_gpu_solid_kmer_support(
        C.c_void_p(contig_first_kmer.ctypes.data), C.c_uint64(len(contig_first_kmer) - 1), C.c_void_p(solid_pos.ctypes.data),
        C.c_void_p(kmer_id.ctypes.data), C.c_uint64(n), C.c_void_p(alns.ctypes.data), C.c_uint64(len(alns)),
        C.c_void_p(cigar.ctypes.data), C.c_uint64(cigar.size), C.c_void_p(seqs.ctypes.data), C.c_uint64(seqs.size),
        C.c_uint32(k), C.c_void_p(cov.ctypes.data), C.c_void_p(sup.ctypes.data)))
    return cov[:n], sup[:n]


def minimiser_support(contig_first_bound, contig_even, bounds, region_first_mini, mini_pos, mini_val, alns, cigar, seqs):
    """hypo_gpu_minimiser_support -> (coverage, support) per minimiser."""
    L = lib()
    L.hypo_gpu_minimiser_support.restype = C.c_int
    n = int(mini_pos.size)
    cov, sup = np.zeros(n + 1, np.uint32), np.zeros(n + 1, np.uint32)
    _check(L.hypo_gpu_minimiser_support(
        C.c_void_p(contig_first_bound.ctypes.data), C.c_void_p(contig_even.ctypes.data), C.c_uint64(len(contig_even)),
        C.c_void_p(bounds.ctypes.data), C.c_uint64(bounds.size), C.c_void_p(region_first_mini.ctypes.data),
        C.c_void_p(mini_pos.ctypes.data), C.c_void_p(mini_val.ctypes.data), C.c_uint64(n),
        C.c_void_p(alns.ctypes.data), C.c_uint64(len(alns)), C.c_void_p(cigar.ctypes.data), C.c_uint64(cigar.size),
        C.c_void_p(seqs.ctypes.data), C.c_uint64(seqs.size), C.c_void_p(cov.ctypes.data), C.c_void_p(sup.ctypes.data)))
    return cov[:n], sup[:n]


def polish_alignments(contigs, regions, drafts, alns, cigar, seqs, k: int) -> List[str]:
    """hypo_gpu_polish_alignments: alignments + region tables -> polished contigs."""
    L = lib()
    L.hypo_gpu_polish_alignments.restype = C.c_int
    cap = int(contigs["len"].astype(np.int64).sum()) * 2 + 1024
    out = np.zeros(cap, np.uint8)
    off = np.zeros(len(contigs) + 1, np.uint64)
    _check(L.hypo_gpu_polish_alignments(
        C.c_void_p(contigs.ctypes.data), C.c_uint64(len(contigs)), C.c_void_p(regions.ctypes.data),
        C.c_uint64(len(regions)), C.c_void_p(drafts.ctypes.data), C.c_uint64(drafts.size),
        C.c_void_p(alns.ctypes.data), C.c_uint64(len(alns)), C.c_void_p(cigar.ctypes.data), C.c_uint64(cigar.size),
        C.c_void_p(seqs.ctypes.data), C.c_uint64(seqs.size), C.c_uint32(k),
        C.c_void_p(out.ctypes.data), C.c_uint64(cap), C.c_void_p(off.ctypes.data)))
    return split_consensus(out, off)


def shutdown() -> None:
    lib().hypo_gpu_shutdown()


def launch_count() -> int:
    return int(lib().hypo_gpu_launch_count())


def consensus_batch_host(batch: WindowBatch, out: Optional[np.ndarray] = None,
                         out_off: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """hypo_gpu_consensus_batch on host buffers; returns (bytes, offsets[n_win+1])."""
    L = lib()
    if out is None:
        cap = int(batch.out_bound().sum()) + 16
        out = np.empty(cap, np.uint8)
    if out_off is None:
        out_off = np.zeros(batch.n_win + 1, np.uint64)
    _check(L.hypo_gpu_consensus_batch(
        batch.win.ctypes.data, batch.n_win, batch.arms.ctypes.data, batch.n_arms,
        batch.packed.ctypes.data, batch.packed.size, out.ctypes.data, out.size, out_off.ctypes.data))
    return out, out_off


def consensus(batch: WindowBatch) -> List[str]:
    out, off = consensus_batch_host(batch)
    return split_consensus(out, off)


def consensus_batch_device(d_win: int, n_win: int, d_arms: int, n_arms: int, d_packed: int,
                           packed_bytes: int, d_out: int, d_out_pos: int, d_out_len: int,
                           stream: int = 0) -> None:
    """hypo_gpu_consensus_batch_device on raw device pointers (e.g. torch tensors' data_ptr())."""
    _check(lib().hypo_gpu_consensus_batch_device(d_win, n_win, d_arms, n_arms, d_packed, packed_bytes,
                                                 d_out, d_out_pos, d_out_len, stream))


def compact_device(d_scratch: int, d_out_pos: int, d_out_len: int, n_win: int, d_compact: int,
                   compact_cap: int, d_off: int, stream: int = 0) -> int:
    """hypo_gpu_compact_device; returns the total number of consensus bytes."""
    total = C.c_uint64(0)
    _check(lib().hypo_gpu_compact_device(d_scratch, d_out_pos, d_out_len, n_win, d_compact, compact_cap,
                                         d_off, C.byref(total), stream))
    return int(total.value)


def last_timing() -> Tuple[float, int, List[int]]:
    """(POA-kernel device ms, POA launches, windows per tier) of the last batch call."""
    ms = C.c_float(0)
    n = C.c_uint32(0)
    tiers = (C.c_uint32 * 8)()
    lib().hypo_gpu_last_timing(C.byref(ms), C.byref(n), tiers)
    # all tiers: 0..7 as in hypo_gpu_last_timing, 8 / 9 = the group tiers (several windows per warp)
    all_tiers = (C.c_uint32 * N_TIERS)()
    lib().hypo_gpu_last_tier_windows(all_tiers, N_TIERS)
    return float(ms.value), int(n.value), [int(x) for x in all_tiers]


def last_rerouted() -> int:
    """Windows the tier probes of the last batch call handed to a later tier untried."""
    return int(lib().hypo_gpu_last_rerouted())


def last_cells() -> int:
    """DP cells (sum of (nodes + 1) x (read length + 1) over all fills) of the last batch call."""
    return int(lib().hypo_gpu_last_cells())


ISSUE_OPS = ("VIADDMNMX.S16x2", "VIMNMX3.S16x2", "VIADDMNMX.S32", "SHFL.UP", "PRMT", "IMAD")


def issue_rate(op: int) -> float:
    """Measured issue rate of one instruction of the fill's inner loop, 10^9 warp instructions / s."""
    v = C.c_double(0)
    _check(lib().hypo_gpu_issue_rate(int(op), C.byref(v)))
    return float(v.value)


def last_fail_hist() -> List[int]:
    """Why windows were abandoned in a capacity tier during the last batch call (by FailReason)."""
    h = (C.c_uint32 * 16)()
    lib().hypo_gpu_last_fail_hist(h)
    return [int(x) for x in h]
