"""Developer aid: print a sweep.jsonl (tools/sweep.py) as a table."""
import json, sys
print(f"{'arms':>4} {'len':>4} {'err':>5} {'type':>5} {'win':>7} {'nodes':>6} {'Mbp/s':>8} {'e2e':>8} {'GCUPS':>7} {'%DPX':>5} {'GB/s':>6} {'tiers':<40} {'rerouted':>8} exact")
for l in open(sys.argv[1]):
    d=json.loads(l); s=d['shape']
    print(f"{s['arms']:4d} {s['length']:4d} {s['err']:5.2f} {s['wtype']:>5} {d['windows']:7d} {d['nodes_avg']:6.0f} {d['mbp_per_s_kernel']:8.2f} {d['mbp_per_s_e2e']:8.2f} {d['gcups']:7.1f} {100*d.get('frac_of_dpx_peak',0):5.2f} {d['hbm_gbs_algorithmic']:6.2f} {str(d['tier_windows']):<40} {d.get('rerouted_by_probe',0):8d} {d['bit_exact']} ({d['bit_exact_checked']})")
