#!/bin/bash
# Developer aid: A/B of library variants on several workloads (kernel Mbp/s via tools/sweep.py rows + bench).
mkdir -p gpurun_out; : > gpurun_out/ab2.log
cp hypo_b200/libhypo_b200.so /tmp/orig.so
for v in build/variants/*.so; do
  cp $v hypo_b200/libhypo_b200.so
  echo "== $(basename $v)" >> gpurun_out/ab2.log
  python bench.py --steps 3 --warmup 2 --windows 500000 --no-cpu-baseline --no-e2e --no-compute-roofline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('headline', round(d['value'],1))" >> gpurun_out/ab2.log
  python bench.py --steps 3 --warmup 2 --windows 2000000 --mix pipeline --no-cpu-baseline --no-e2e --no-compute-roofline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pipeline mix', round(d['value'],1))" >> gpurun_out/ab2.log
  python bench.py --steps 3 --warmup 2 --windows 200000 --kind prefix --length 110 --no-cpu-baseline --no-e2e --no-compute-roofline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prefix-heavy 30x110', round(d['value'],1))" >> gpurun_out/ab2.log
  python bench.py --steps 3 --warmup 2 --windows 200000 --kind mixed --length 100 --no-cpu-baseline --no-e2e --no-compute-roofline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mixed 30x100', round(d['value'],1))" >> gpurun_out/ab2.log
  python tools/sweep.py --arms 30 --lengths 120,500 --errs 0.01 --long --budget 4e10 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); s=d['shape']
    if s['wtype']=='LONG': print('LONG', s['arms'], s['length'], round(d['mbp_per_s_kernel'],2), d['bit_exact'])" >> gpurun_out/ab2.log
done
cp /tmp/orig.so hypo_b200/libhypo_b200.so
cat gpurun_out/ab2.log
