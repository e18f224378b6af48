#!/bin/bash
# Evidence for the TMA-staging decision: parity of the variant, A/B throughput, and an
# ncu comparison of the compact-tier kernel with and without staging on 100 K windows.
mkdir -p gpurun_out
echo "== parity tests, TMA variant" > gpurun_out/r2_tma_evidence.txt
bash tools/gpu_variant_tests.sh tma >> gpurun_out/r2_tma_evidence.txt 2>&1
cp hypo_b200/libhypo_b200.so /tmp/orig2.so
M=gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_no_instruction.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active
for v in base tma; do
  cp build/variants/$v.so hypo_b200/libhypo_b200.so
  echo "== ncu, $v, 100 K windows of 30 x 120 bp" >> gpurun_out/r2_tma_evidence.txt
  ncu --metrics $M --clock-control none -k regex:poa_kernel -c 1 --csv --log-file gpurun_out/r2_tma_ncu_$v.csv python bench.py --windows 100000 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-compute-roofline > /dev/null 2>&1
  python - "$v" >> gpurun_out/r2_tma_evidence.txt <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(f"gpurun_out/r2_tma_ncu_{sys.argv[1]}.csv")) if len(r) > 10]
h = rows[0]
for r in rows[1:]:
    d = dict(zip(h, r))
    print(f"{d['Metric Name']:85s} {d['Metric Unit']:14s} {d['Metric Value']}")
PY
done
cp /tmp/orig2.so hypo_b200/libhypo_b200.so
cat gpurun_out/r2_tma_evidence.txt
