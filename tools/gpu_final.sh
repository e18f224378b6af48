#!/bin/bash
# Round-end evidence on one GPU: the whole GPU test suite, the bench lines that go to profiles/, the fused-step bench.
R=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/${R}_gputests.txt
python bench.py 2> gpurun_out/${R}_bench.err | tail -1 > gpurun_out/${R}_bench_1gpu.json
python bench.py --stream data/captured/ecoli5mb_ctg1.inspect.gz data/captured/ecoli5mb_ctg2.inspect.gz 2>/dev/null | tail -1 > gpurun_out/${R}_bench_ecoli5mb.json
python bench.py --stream data/captured/ecoli5mb_ctg1.inspect.gz data/captured/ecoli5mb_ctg2.inspect.gz --repeat 12 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/${R}_bench_capture_x12.json
python bench.py --mix pipeline --windows 4000000 2>/dev/null | tail -1 > gpurun_out/${R}_bench_pipeline_mix.json
python bench.py --impl reference 2>/dev/null | tail -1 > gpurun_out/${R}_bench_reference_arm.json
for f in data/_scratch/n3_1mb.npz data/_scratch/n3_ecoli5mb.npz; do
  [ -f $f ] && python tools/bench_arms.py $f 5 > gpurun_out/${R}_bench_fused_$(basename $f .npz).json
done
for f in gpurun_out/${R}_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
if "value" in d:
    e = d.get("e2e") or {}
    print(d["config"]["workload"][:70] if "config" in d else "", "| value", round(d["value"], 1), "| e2e", round(e.get("value", 0), 1),
          "| packed", round((d.get("e2e_packed") or {}).get("value", 0), 1), "| cpu", (d.get("cpu_baseline") or {}).get("value"),
          "| parity", d.get("parity_spot_check"))
else:
    print({k: d[k] for k in ("alignments", "windows", "fused_polish_alignments_s", "polished_contigs_identical_to_reference_cli", "speedup_vs_reference_phases", "support_counting")})
PY
done
