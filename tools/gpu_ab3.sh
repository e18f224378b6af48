#!/bin/bash
# Developer aid: A/B of library variants (build/variants/*.so): headline, pipeline mix, a few sweep rows with tier histograms.
mkdir -p gpurun_out; : > gpurun_out/ab3.log
cp hypo_b200/libhypo_b200.so /tmp/orig.so
for v in build/variants/*.so; do
  cp $v hypo_b200/libhypo_b200.so
  echo "== $(basename $v)" >> gpurun_out/ab3.log
  [ -n "$AB_SKIP_BENCH" ] || python bench.py --steps 3 --warmup 2 --windows 500000 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('headline', round(d['value'],1), d['config']['tier_windows'])" >> gpurun_out/ab3.log
  [ -n "$AB_SKIP_BENCH" ] || python bench.py --steps 3 --warmup 2 --windows 2000000 --mix pipeline --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pipeline mix', round(d['value'],1), d['config']['tier_windows'])" >> gpurun_out/ab3.log
  python tools/sweep.py ${AB_SWEEP:---arms 30 --lengths 120,250 --errs 0.01,0.05 --budget 3e10} 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); s=d['shape']
    print(s['wtype'], s['arms'], s['length'], s['err'], round(d['mbp_per_s_kernel'],2), d['tier_windows'], d['abandoned_by_reason'], d['bit_exact'])" >> gpurun_out/ab3.log
done
cp /tmp/orig.so hypo_b200/libhypo_b200.so
cat gpurun_out/ab3.log
