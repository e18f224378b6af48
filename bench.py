#!/usr/bin/env python
"""bench.py — Mbp polished / s of the POA-consensus hot path (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic windows.  Workload at N=1:
BASELINE.json configs[1], "synthetic 1M windows (30 reads x 120 bp)", all-internal SHORT windows,
scores 5/-4/-8 (generator: hypo_b200/host/host_capi.cpp, distribution of BASELINE.md §3).
N>1: every rank polishes its own shard of the same shape (weak scaling, no data-path
collective), then the consensus bytes are gathered to rank 0 over NCCL.

  value  : whole-job Mbp/s with the batch resident in HBM (device pointers through
           hypo_gpu_consensus_batch_device + on-device compaction [+ NCCL gather]); CUDA events.
  e2e    : the same metric through the host-buffer C-ABI call hypo_gpu_consensus_batch
           (pinned host buffers; H2D, kernels, compaction, D2H all inside the timed region).
  roofline / cpu_baseline: see DESIGN.md §measurement.

`--impl reference` times the reference's own CPU implementation (oracle/_ref, compiled from the
unmodified reference sources; falls back to the C port when it was not built) on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SCORES = (5, -4, -8, 3, -5, -4)
METRIC = "Mbp polished/sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--windows", type=int, default=1_000_000, help="windows per GPU")
    ap.add_argument("--arms", type=int, default=30)
    ap.add_argument("--length", type=int, default=120)
    ap.add_argument("--kind", default="internal")
    ap.add_argument("--err", type=float, default=0.01)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--cpu-sample", type=int, default=12000, help="windows timed on the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--mix", default="", choices=["", "pipeline"],
                    help="pipeline: window shapes as the reference CLI produces them (SURVEY.md §6: draft length "
                         "p50 9 / p90 68 / max 100, 10-48 arms, 8 %% of the windows with prefix/suffix arms) "
                         "instead of one fixed shape; --windows is the total")
    return ap.parse_args()


def workload_name(a):
    if a.mix == "pipeline":
        return (f"synthetic {a.windows} windows/GPU, pipeline shape mix (draft 6-100 bp, median 9; 11-48 arms; "
                f"8 % with prefix/suffix arms; SHORT, err {a.err})")
    return f"synthetic {a.windows} windows/GPU ({a.arms} reads x {a.length} bp, {a.kind}, SHORT, err {a.err})"


# (draft length, weight) and (arms, weight): the two datasets measured in SURVEY.md §6
PIPELINE_LEN = ((6, 0.20), (9, 0.32), (14, 0.14), (25, 0.12), (40, 0.08), (68, 0.08), (100, 0.06))
PIPELINE_ARMS = ((11, 0.30), (16, 0.15), (30, 0.15), (39, 0.25), (48, 0.15))


def make_batch(a, seed):
    """The synthetic shard of one rank (one fixed shape, or the pipeline shape mix)."""
    from hypo_b200.batch import concat_batches
    from hypo_b200.hostlib import synth_batch
    if a.mix != "pipeline":
        return synth_batch(seed, a.windows, a.length, a.arms, a.kind, a.err)
    parts, k = [], 0
    for ln, wl in PIPELINE_LEN:
        for na, wa in PIPELINE_ARMS:
            for kind, wk in (("internal", 0.92), ("mixed", 0.08)):
                n = int(round(a.windows * wl * wa * wk))
                k += 1
                if n > 0:
                    parts.append(synth_batch(seed + 104729 * k, n, ln, na, kind, a.err))
    b = concat_batches(parts, {"mix": "pipeline"})
    perm = np.random.default_rng(seed).permutation(b.n_win)
    return b.select(perm)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_reference_time(batch, n_sample, schedule=1):
    """Times the reference's CPU path (oracle/_ref) or, if absent, the C port on a bounded sample.
    Returns (Mbp/s, cores, kind, sample description)."""
    from tests import oracle_util as ou
    n = min(n_sample, batch.n_win)
    sub = batch.select(np.arange(n))
    cores = os.cpu_count() or 1
    if ou.ref_lib(False) is not None:
        # the faster of the reference's two CPU engines (SISD default build, AVX2 -mavx2 build)
        best, which = None, None
        for simd in (False, True):
            if ou.ref_lib(simd) is None:
                continue
            _, _, sec = ou.ref_consensus(sub, SCORES, simd=simd, threads=cores, schedule=schedule)
            if best is None or sec < best:
                best, which = sec, ("AVX2" if simd else "SISD")
        kind = "reference"
        desc = (f"first {n} windows of the workload, unmodified reference Window::generate_consensus under OpenMP "
                f"schedule(dynamic,1), {cores} threads, faster of SISD/AVX2 builds ({which})")
        sec = best
    else:
        _, sec = ou.oracle_consensus(sub, SCORES, threads=cores)
        kind = "port"
        desc = f"first {n} windows of the workload, oracle/poa_oracle.c under OpenMP, {cores} threads"
    return sub.polished_bp / 1e6 / sec, cores, kind, desc


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from hypo_b200.hostlib import synth_batch
    n_gen = max(a.cpu_sample, 1000)
    gen = argparse.Namespace(**vars(a))
    gen.windows = n_gen   # the reference arm times a bounded sample of the same workload
    batch = make_batch(gen, a.seed)
    # bounded sample per step so that steps+warmup finish within minutes
    per_step = max(500, min(a.cpu_sample, n_gen) // 3)
    vals = []
    for s in range(a.warmup + a.steps):
        v, cores, kind, desc = cpu_reference_time(batch, per_step)
        if s >= a.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    sample = desc.replace(f"first {min(per_step, batch.n_win)}", f"{per_step}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mbp/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * (batch.select(np.arange(per_step)).polished_bp / 1e6) / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_name(a), "scores": list(SCORES)},
        "cpu_baseline": {"value": value, "unit": "Mbp/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return

    import torch
    import torch.distributed as dist
    from hypo_b200 import native
    from hypo_b200.hostlib import synth_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("[Hypo::GPU] Error: bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    native.init(SCORES, local)

    # ---- synthetic shard of this rank --------------------------------------------------
    batch = make_batch(a, a.seed + 7919 * rank)
    n_win, n_arms = batch.n_win, batch.n_arms
    bp = batch.polished_bp
    bound = batch.out_bound()
    out_pos_np = np.concatenate([[0], np.cumsum(bound)[:-1]]).astype(np.uint64)
    scratch_bytes = int(bound.sum()) + 16

    def to_dev(x):
        return torch.from_numpy(x.view(np.uint8).reshape(-1)).to(dev)

    d_win, d_arms, d_packed = to_dev(batch.win), to_dev(batch.arms), to_dev(batch.packed)
    d_out_pos = to_dev(out_pos_np)
    d_scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=dev)
    d_out_len = torch.zeros(n_win, dtype=torch.int32, device=dev)
    compact_cap = int(bp * 1.5) + 4096
    d_compact = torch.zeros(compact_cap, dtype=torch.uint8, device=dev)
    d_off = torch.zeros(n_win + 1, dtype=torch.int64, device=dev)
    gather_list = [torch.empty(compact_cap, dtype=torch.uint8, device=dev) for _ in range(world)] \
        if (world > 1 and rank == 0) else None
    stream = torch.cuda.current_stream().cuda_stream
    totals = []

    def step_device():
        native.consensus_batch_device(d_win.data_ptr(), n_win, d_arms.data_ptr(), n_arms, d_packed.data_ptr(),
                                      batch.packed.size, d_scratch.data_ptr(), d_out_pos.data_ptr(),
                                      d_out_len.data_ptr(), stream)
        total = native.compact_device(d_scratch.data_ptr(), d_out_pos.data_ptr(), d_out_len.data_ptr(), n_win,
                                      d_compact.data_ptr(), compact_cap, d_off.data_ptr(), stream)
        totals.append(total)
        if world > 1:   # final consensus gather: the only collective of the path
            dist.gather(d_compact, gather_list, dst=0)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step_device()
    sync_all()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = native.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    poa_ms = []
    ev0.record()
    for _ in range(a.steps):
        step_device()
        poa_ms.append(native.last_timing()[0])
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = native.launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    tiers = native.last_timing()[2]
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / a.steps
    value = (bp * world) / 1e6 / (ms_per_step / 1e3)
    total_cons = totals[-1]

    # parity spot-check of what was just timed: first windows against the CPU oracle
    ok = None
    if rank == 0:
        from tests.oracle_util import oracle_consensus
        k = min(256, n_win)
        off = d_off[: k + 1].cpu().numpy()
        raw = d_compact[: int(off[k])].cpu().numpy().tobytes()
        got = [raw[int(off[i]):int(off[i + 1])].decode() for i in range(k)]
        want, _ = oracle_consensus(batch.select(np.arange(k)), SCORES)
        ok = got == want

    # ---- end to end through the host-buffer C ABI ---------------------------------------
    e2e = None
    if not a.no_e2e:
        def pin(x):
            t_ = torch.from_numpy(x.view(np.uint8).reshape(-1)).pin_memory()
            return t_
        p_win, p_arms, p_packed = pin(batch.win), pin(batch.arms), pin(batch.packed)
        p_out = torch.empty(compact_cap, dtype=torch.uint8).pin_memory()
        p_off = torch.empty(n_win + 1, dtype=torch.int64).pin_memory()
        L = native.lib()

        def step_e2e():
            rc = L.hypo_gpu_consensus_batch(p_win.data_ptr(), n_win, p_arms.data_ptr(), n_arms, p_packed.data_ptr(),
                                            batch.packed.size, p_out.data_ptr(), compact_cap, p_off.data_ptr())
            if rc != 0:
                raise native.HypoGpuError(rc, L.hypo_gpu_last_error().decode())

        step_e2e()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e_total = int(p_off[n_win].item())
        e2e = {"value": (bp * world) / 1e6 / (dt / a.steps), "unit": "Mbp/s",
               "h2d_bytes_per_step": int(batch.win.nbytes + batch.arms.nbytes + batch.packed.nbytes),
               "d2h_bytes_per_step": int(e2e_total + 8 * (n_win + 1))}
        if rank == 0:
            # e2e output must equal the device-resident output
            same = bytes(p_out[:e2e_total].numpy().tobytes()) == d_compact[:total_cons].cpu().numpy().tobytes()
            ok = bool(ok and same)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (POA tier kernels) -------------------------------
    peak, peak_src = measured_peak_gbs()
    alg_bytes = batch.algorithmic_bytes(total_cons)
    k_ms = float(np.mean(poa_ms))
    achieved = alg_bytes / 1e9 / (k_ms / 1e3)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if (not a.mix and tj.get("windows") == a.windows and tj.get("arms") == a.arms
                    and tj.get("length") == a.length):
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "kernel": "poa_kernel (all tier launches of one step)", "kernel_ms": k_ms,
                "kernel_share_of_step": k_ms / ms_per_step,
                "note": "integer DP on an irregular DAG: issue/latency bound, not HBM bound (DESIGN.md §roofline)"}

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        v, cores, kind, desc = cpu_reference_time(batch, a.cpu_sample)
        cpu = {"value": v, "unit": "Mbp/s", "cores": cores, "kind": kind, "sample": desc}

    line = {
        "metric": METRIC, "value": value, "unit": "Mbp/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16", "data": "synthetic",
        "config": {"workload": workload_name(a), "scores": list(SCORES), "windows_total": a.windows * world,
                   "polished_bp_per_step": bp * world, "sharding": f"{world} rank(s), independent window shards",
                   "l2": "inputs (%.0f MB/GPU) larger than the 126 MB L2; no flush needed" %
                         ((batch.win.nbytes + batch.arms.nbytes + batch.packed.nbytes) / 1e6),
                   "tier_windows": tiers},
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "parity_spot_check": ok,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
