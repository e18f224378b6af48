# Developer aid: A/B of the group tiers on the pipeline shape mix and the captured E. coli-sized set.
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-compute-roofline"
for opt in "group_tiers=0" "group_sort=0" "group_tiers=1"; do
  python bench.py --mix pipeline --windows 2000000 $B --option $opt 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mix   $opt', round(d['value'],1), 'Mbp/s', round(d['roofline']['kernel_ms'],2), 'ms', d['config']['tier_windows'])"
  python bench.py --stream data/captured/ecoli5mb_ctg1.inspect.gz data/captured/ecoli5mb_ctg2.inspect.gz --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-compute-roofline --option $opt 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ecoli $opt', round(d['value'],1), 'Mbp/s', round(d['roofline']['kernel_ms'],2), 'ms', d['config']['tier_windows'])"
done
