#!/usr/bin/env python
"""Times the fused step hypo_gpu_polish_alignments (alignments in -> arm extraction -> POA of every window ->
stitched contigs out) on a captured run of the reference CLI (tools/capture/make_n3_input.py), checks the
polished contigs against the reference's own output, and prints one JSON line with the reference's Monitor
times of the phases the call replaces.   python tools/bench_arms.py data/_scratch/n3_1mb.npz [steps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from hypo_b200 import native  # noqa: E402


def main():
    z = np.load(sys.argv[1])
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    meta = json.loads(str(z["meta"]))
    args = (z["contigs"], z["regions"], z["drafts"], z["alns"], z["cigar"], z["seqs"], int(meta["k"]))
    native.init((5, -4, -8, 3, -5, -4), 0)
    out = native.polish_alignments(*args)   # warm-up (allocations)
    same = out == [str(s) for s in z["polished"]]
    t0 = time.perf_counter()
    for _ in range(steps):
        native.polish_alignments(*args)
    fused = (time.perf_counter() - t0) / steps
    batch, _ = native.extract_arms(*args)
    t0 = time.perf_counter()
    for _ in range(steps):
        native.extract_arms(*args)
    extract = (time.perf_counter() - t0) / steps
    ref = meta["reference_phase_s"]
    ref_sum = sum(ref[p] for p in meta["replaced_phases"])
    n2 = None
    if "spos" in z:   # support counting (N2) on the same alignments, against the counters the reference dumped
        a3 = (z["alns"], z["cigar"], z["seqs"])
        c, s = native.solid_kmer_support(z["kfirst"], z["spos"], z["kid"], *a3, int(meta["k"]))
        ok_k = bool((c == z["kcov"]).all() and (s == z["ksup"]).all())
        t0 = time.perf_counter()
        for _ in range(steps):
            native.solid_kmer_support(z["kfirst"], z["spos"], z["kid"], *a3, int(meta["k"]))
        t_k = (time.perf_counter() - t0) / steps
        c, s = native.minimiser_support(z["bfirst"], z["even"], z["bounds"], z["rfirst"], z["mpos"], z["mval"], *a3)
        ok_m = bool((c == z["mcov"]).all() and (s == z["msup"]).all())
        t0 = time.perf_counter()
        for _ in range(steps):
            native.minimiser_support(z["bfirst"], z["even"], z["bounds"], z["rfirst"], z["mpos"], z["mval"], *a3)
        t_m = (time.perf_counter() - t0) / steps
        n2 = {"solid_kmers": int(z["spos"].size), "minimisers": int(z["mpos"].size),
              "solid_kmer_support_s": t_k, "minimiser_support_s": t_m,
              "counters_identical_to_reference_cli": {"solid_kmers": ok_k, "minimisers": ok_m},
              "reference_cli_s": {"Solid kmers support update": ref.get("Solid kmers support update"),
                                  "Minimisers support update": ref.get("Minimisers support update")}}
    h2d = sum(int(a.nbytes) for a in args[:6])
    print(json.dumps({
        "input": os.path.basename(sys.argv[1]), "alignments": int(len(z["alns"])), "regions": int(len(z["regions"])),
        "windows": batch.n_win, "arms": batch.n_arms, "polished_bp_in_windows": batch.polished_bp,
        "contig_bp": int(z["contigs"]["len"].astype(np.int64).sum()),
        "fused_polish_alignments_s": fused, "extract_arms_s_incl_batch_copy_back": extract,
        "h2d_bytes": h2d, "polished_contigs_identical_to_reference_cli": bool(same),
        "reference_cli": {"threads": meta["reference_cli_threads"], "phases_s": {p: ref[p] for p in meta["replaced_phases"]},
                          "sum_s": ref_sum, "host": "authoring container, 8 vCPU"},
        "speedup_vs_reference_phases": ref_sum / fused, "support_counting": n2}))


if __name__ == "__main__":
    main()
