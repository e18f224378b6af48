# Developer aid: group tiers on / off by batch size (captured set 1x / 3x / 12x, pipeline mix 100 K / 400 K windows).
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-compute-roofline"
S="--stream data/captured/ecoli5mb_ctg1.inspect.gz data/captured/ecoli5mb_ctg2.inspect.gz"
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"],1), "Mbp/s", round(d["roofline"]["kernel_ms"],2), "ms", d["config"]["tier_windows"])'
for opt in group_tiers=0 group_tiers=1; do
  for r in 1 3 12; do python bench.py $S --repeat $r $B --option $opt 2>/dev/null | python -c "$P" "ecoli x$r $opt"; done
  for n in 100000 400000; do python bench.py --mix pipeline --windows $n $B --option $opt 2>/dev/null | python -c "$P" "mix $n $opt"; done
done
