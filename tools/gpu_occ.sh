#!/bin/bash
# Developer aid: throughput of the T0 tier vs warps per SM (occupancy curve).
mkdir -p gpurun_out
for cfg in "4 2" "6 2" "8 2" "8 1"; do
  set -- $cfg
  echo "wpb=$1 bps=$2" >> gpurun_out/occ.log
  HYPO_B200_T0_WPB=$1 HYPO_B200_T0_BPS=$2 python bench.py --steps 2 --warmup 1 --windows 300000 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])" >> gpurun_out/occ.log
done
cat gpurun_out/occ.log
