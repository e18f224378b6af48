import json,sys
for l in sys.stdin:
    d=json.loads(l); s=d["shape"]; print(sys.argv[1], s["arms"], s["length"], s["err"], s["wtype"], d["windows"], "nodes", round(d["nodes_avg"]), round(d["mbp_per_s_kernel"],2), "Mbp/s", round(d["gcups"],1), "GCUPS", d["tier_windows"], d["abandoned_by_reason"], d["bit_exact"], d["bit_exact_checked"])
