"""Developer aid: per-function shares (instructions, stall samples) of one kernel from an
`ncu --page source --csv --print-source cuda,sass` dump, for kernels whose device functions live in one
main source file (instructions inlined from headers are attributed to the function around them)."""
import csv, re, sys
path, main_src = sys.argv[1], sys.argv[2]
src = open(main_src).read().split('\n')
regions = []
for n, l in enumerate(src, 1):
    m = re.match(r'^__device__ .*?(\w+)\(', l) or re.match(r'^__global__ .* (\w+)\(', l) or re.match(r'^static __device__ .*?(\w+)\(', l)
    if m: regions.append((n, m.group(1)))
def region(n):
    name = 'top'
    for s, nm in regions:
        if n >= s: name = nm
        else: break
    return name
rows = list(csv.reader(open(path)))
recs = {}   # address -> (is_main, line, metrics)
cur_file = None; hdr = None; cur_line = None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur_file = r[1]; continue
    if len(r) >= 2 and r[0] == 'Function Name': continue
    if len(r) > 3 and r[0] == 'Line No': hdr = r; ci = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] != '': cur_line = int(r[0]); continue
    a = r[2]
    if not a.startswith('0x'): continue
    def g(c):
        try: return int(r[ci[c]])
        except Exception: return 0
    m = dict(inst=g('Instructions Executed'), thr=g('Thread Instructions Executed'), samp=g('# Samples'), noinst=g('stall_no_inst'), lsb=g('stall_long_sb'),
             ssb=g('stall_short_sb'), wait=g('stall_wait'), br=g('stall_branch_resolving'), math=g('stall_math'))
    is_main = cur_file.endswith(main_src.split('/')[-1])
    if a not in recs or (is_main and not recs[a][0]): recs[a] = (is_main, cur_line, m, r[3])
agg = {}; fn = 'top'
for a in sorted(recs, key=lambda x: int(x, 16)):
    is_main, line, m, sass = recs[a]
    if is_main: fn = region(line)
    d = agg.setdefault(fn, dict(static=0, inst=0, thr=0, samp=0, noinst=0, lsb=0, ssb=0, wait=0, br=0, math=0))
    d['static'] += 1
    for k in m: d[k] += m[k]
tot = {k: sum(d[k] for d in agg.values()) for k in next(iter(agg.values()))}
print(f"static {tot['static']} inst {tot['inst']:.3e} thr/inst {tot['thr']/max(tot['inst'],1):.1f} samples {tot['samp']}: no_inst {100*tot['noinst']/tot['samp']:.1f}% long_sb {100*tot['lsb']/tot['samp']:.1f}% short_sb {100*tot['ssb']/tot['samp']:.1f}% wait {100*tot['wait']/tot['samp']:.1f}% branch {100*tot['br']/tot['samp']:.1f}% math {100*tot['math']/tot['samp']:.1f}%")
print(f"{'function':22s} static  inst%  thr/inst samp%  no_inst% long_sb% short_sb% wait%  branch%")
for k, d in sorted(agg.items(), key=lambda kv: -kv[1]['samp']):
    print(f"{k:22s} {d['static']:5d} {100*d['inst']/tot['inst']:6.1f} {d['thr']/max(d['inst'],1):7.1f} {100*d['samp']/tot['samp']:6.1f} {100*d['noinst']/max(tot['noinst'],1):8.1f} {100*d['lsb']/max(tot['lsb'],1):8.1f} {100*d['ssb']/max(tot['ssb'],1):8.1f} {100*d['wait']/max(tot['wait'],1):6.1f} {100*d['br']/max(tot['br'],1):7.1f}")
