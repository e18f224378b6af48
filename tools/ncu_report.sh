#!/bin/bash
# Developer aid: summarise gpurun_out/prof.ncu-rep (per-function shares, stall reasons, key raw metrics).
cd /root/repo
ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src.csv 2>/dev/null
ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source sass > gpurun_out/sass.csv 2>/dev/null
ncu -i gpurun_out/prof.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/raw.csv
python - <<'PY'
import csv, re
rows = list(csv.reader(open('gpurun_out/raw.csv')))
h = rows[0]; u = rows[1]; v = rows[2]
want = ['gpu__time_duration.sum','smsp__inst_executed.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','sm__warps_active.avg.per_cycle_active','launch__registers_per_thread','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']
for a,b,c in zip(h,u,v):
    if a in want or ('issue_stalled' in a and a.endswith('.ratio')): print(f"{a:90s} {b:12s} {c}")
src = open('hypo_b200/csrc/poa_kernel.cu').read().split('\n')
rows = list(csv.reader(open('gpurun_out/src.csv')))
hdr_i = next(i for i, r in enumerate(rows) if len(r) > 3 and r[0] == 'Line No' and r[2] == 'Address')
hdr = rows[hdr_i]; ci={h:i for i,h in enumerate(hdr)}
regions=[]
for n,l in enumerate(src,1):
    m = re.match(r'^__device__ .*?(\w+)\(', l) or re.match(r'^__global__ .* (\w+)\(', l)
    if m: regions.append((n,m.group(1)))
def region(n):
    name='top'
    for s,nm in regions:
        if n>=s: name=nm
        else: break
    return name
cur=None; seen=set(); agg={}
cols=['Instructions Executed','# Samples','stall_no_inst','stall_long_sb','stall_short_sb','stall_wait']
for r in rows[hdr_i+1:]:
    if len(r)<len(hdr): continue
    if r[0]=='Line No': break
    if r[0]!='': cur=int(r[0]); continue
    if r[2] in ('...','-') or r[2] in seen: continue
    seen.add(r[2])
    k=region(cur); a=agg.setdefault(k,[0]*(len(cols)+1)); a[0]+=1
    for j,c in enumerate(cols):
        try: a[j+1]+=int(r[ci[c]])
        except: pass
tot=[sum(a[j] for a in agg.values()) for j in range(len(cols)+1)]
print(f"static {tot[0]}  inst {tot[1]:.3e} samples {tot[2]}  no_inst {100*tot[3]/tot[2]:.1f}% long_sb {100*tot[4]/tot[2]:.1f}% short_sb {100*tot[5]/tot[2]:.1f}% wait {100*tot[6]/tot[2]:.1f}%")
print(f"{'function':20s} static  inst%  samp%  no_inst% long_sb% short_sb% wait%   (shares of each column)")
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][2])[:26]:
    print(f"{k:20s} {a[0]:5d} {100*a[1]/tot[1]:6.1f} {100*a[2]/tot[2]:6.1f} {100*a[3]/max(tot[3],1):7.1f} {100*a[4]/max(tot[4],1):8.1f} {100*a[5]/max(tot[5],1):8.1f} {100*a[6]/max(tot[6],1):6.1f}")
PY
