#!/bin/bash
# Developer aid: A/B throughput of library variants under build/variants/*.so (300K windows each).
mkdir -p gpurun_out; : > gpurun_out/ab.log
cp hypo_b200/libhypo_b200.so /tmp/orig.so
for v in build/variants/*.so; do
  cp $v hypo_b200/libhypo_b200.so
  for rep in 1 2; do
    echo -n "$(basename $v) " >> gpurun_out/ab.log
    python bench.py --steps 3 --warmup 2 --windows ${AB_WINDOWS:-300000} ${AB_ARGS} --no-cpu-baseline --no-e2e --no-compute-roofline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['parity_spot_check'], d['config']['tier_windows'][:3])" >> gpurun_out/ab.log
  done
done
cp /tmp/orig.so hypo_b200/libhypo_b200.so
cat gpurun_out/ab.log
