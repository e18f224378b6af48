#!/bin/bash
# Round-end evidence on one GPU, most important first (everything lands in gpurun_out/, then profiles/).
R=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/${R}_gputests.txt
python bench.py 2> gpurun_out/${R}_bench.err | tail -1 > gpurun_out/${R}_bench_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-compute-roofline --no-extra > gpurun_out/${R}_launches.log 2>&1
python bench.py --mix pipeline --windows 4000000 2>/dev/null | tail -1 > gpurun_out/${R}_bench_pipeline_mix.json
python bench.py --mix pipeline --windows 4000000 --no-cpu-baseline --no-e2e --option group_tiers=0 2>/dev/null | tail -1 > gpurun_out/${R}_bench_pipeline_mix_one_warp_tiers_only.json
python bench.py --stream data/captured/ecoli5mb_ctg1.inspect.gz data/captured/ecoli5mb_ctg2.inspect.gz 2>/dev/null | tail -1 > gpurun_out/${R}_bench_ecoli5mb.json
python bench.py --stream data/captured/ecoli5mb_ctg1.inspect.gz data/captured/ecoli5mb_ctg2.inspect.gz --repeat 12 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/${R}_bench_capture_x12.json
python bench.py --stream data/captured/ecoli5mb_ctg1.inspect.gz data/captured/ecoli5mb_ctg2.inspect.gz --no-cpu-baseline --no-e2e --option group_tiers=2 2>/dev/null | tail -1 > gpurun_out/${R}_bench_ecoli5mb_group_tiers_forced.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${R}_launches_pipeline_mix.csv \
    python bench.py --mix pipeline --windows 1000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-compute-roofline > gpurun_out/${R}_launches_mix.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:poa_group_kernel -s 1 -c 1 -o gpurun_out/${R}_gprof -f \
    python bench.py --mix pipeline --windows 1000000 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-compute-roofline > gpurun_out/${R}_gncu.log 2>&1
ncu -i gpurun_out/${R}_gprof.ncu-rep --page details > gpurun_out/${R}_poa_group_kernel_Tq_details.txt 2>/dev/null
python tools/sweep.py --long --check-cells 4e9 > gpurun_out/${R}_sweep.jsonl 2> gpurun_out/${R}_sweep.err
python bench.py --impl reference 2>/dev/null | tail -1 > gpurun_out/${R}_bench_reference_arm.json
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck.log
HYPO_SANITIZE_NO_TEAMS=1 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_run.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 6 python tools/sanitize_run.py > gpurun_out/sanitizer_racecheck_teams.log 2>&1; echo "racecheck (teams) exit $?" >> gpurun_out/sanitizer_racecheck_teams.log
for f in gpurun_out/sanitizer_memcheck.log gpurun_out/sanitizer_racecheck.log gpurun_out/sanitizer_racecheck_teams.log; do tail -n 3 $f; done
timeout 400 python tools/fuzz_parity.py 240 31337 2>&1 | tail -3 | tee gpurun_out/${R}_fuzz.txt
for f in gpurun_out/${R}_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print("unreadable", e); sys.exit(0)
if "value" in d:
    e = d.get("e2e") or {}
    print((d.get("config") or {}).get("workload", "")[:70], "| value", round(d["value"], 1), "| e2e", round(e.get("value", 0), 1),
          "| packed", round((d.get("e2e_packed") or {}).get("value", 0), 1), "| cpu", (d.get("cpu_baseline") or {}).get("value"),
          "| parity", d.get("parity_spot_check"), "| tiers", (d.get("config") or {}).get("tier_windows"))
    if d.get("extra"): print("   extra:", json.dumps(d["extra"])[:900])
PY
done
