"""Developer aid: aggregate an `ncu --page source --csv --print-source cuda,sass` dump per source line
and per function region of poa_kernel.cu."""
import csv, sys, re
path = sys.argv[1]
src = open('/root/repo/hypo_b200/csrc/poa_kernel.cu').read().split('\n') if len(sys.argv) < 3 else open(sys.argv[2]).read().split('\n')
rows = list(csv.reader(open(path)))
# first table: header at the row starting with "Line No","Source","Address"
hdr_i = next(i for i, r in enumerate(rows) if len(r) > 3 and r[0] == 'Line No' and r[2] == 'Address')
hdr = rows[hdr_i]
ci = {h: i for i, h in enumerate(hdr)}
inst_c = hdr.index('Instructions Executed'); samp_c = hdr.index('# Samples')
line_inst = {}; line_samp = {}
cur = None
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr): continue
    if r[0] == 'Line No': break
    if r[0] != '':
        cur = int(r[0]);
        continue   # source line summary row (aggregated) - skip, we sum SASS rows
    if cur is None: continue
    try:
        line_inst[cur] = line_inst.get(cur, 0) + int(r[inst_c]); line_samp[cur] = line_samp.get(cur, 0) + int(r[samp_c])
    except ValueError: pass
tot_i = sum(line_inst.values()); tot_s = sum(line_samp.values())
# function regions
regions = []
for n, l in enumerate(src, 1):
    m = re.match(r'^(?:template.*\n)?__device__ .*?(\w+)\(', l) or re.match(r'^__global__ .* (\w+)\(', l)
    if m: regions.append((n, m.group(1)))
regions.append((10**9, 'end'))
def region(n):
    name = 'top'
    for s, nm in regions:
        if n >= s: name = nm
        else: break
    return name
agg = {}
for n in line_inst:
    k = region(n); a = agg.setdefault(k, [0, 0]); a[0] += line_inst[n]; a[1] += line_samp.get(n, 0)
print(f"total inst {tot_i:.3e} samples {tot_s}")
for k, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:22s} inst {100*i/tot_i:5.1f}%  samples {100*s/max(tot_s,1):5.1f}%")
print("--- top lines by samples")
for n in sorted(line_samp, key=lambda n: -line_samp[n])[:25]:
    print(f"{n:4d} inst {100*line_inst[n]/tot_i:5.1f}% samp {100*line_samp[n]/max(tot_s,1):5.1f}%  {src[n-1].strip()[:100]}")
