# Developer aid: the team fill (several warps per window) on the large window shapes.
fmt='import json,sys
for l in sys.stdin:
    d=json.loads(l); s=d["shape"]; print(sys.argv[1], s["arms"], s["length"], s["err"], s["wtype"], d["windows"], round(d["mbp_per_s_kernel"],2), "Mbp/s", round(d["gcups"],1), "GCUPS", d["tier_windows"], d["abandoned_by_reason"], d["bit_exact"], d["bit_exact_checked"])'
for opt in ${1:-teams=1}; do
  python tools/sweep.py --only-long --lengths 250,500 --budget 1e11 --check 32 --option $opt | python -c "$fmt" $opt
  python tools/sweep.py --arms 10,30 --lengths 250,500 --errs 0.01 --budget 1e11 --check 24 --option $opt | python -c "$fmt" $opt
done
