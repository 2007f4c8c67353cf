"""Join ncu's per-SASS-instruction counters with nvdisasm line info: executed warp-instructions
and stall samples per CUDA source line.  usage: line_profile.py <ncu source csv> <nvdisasm -g -c txt> <mangled substr> <norm>"""
import collections, csv, re, sys
src_csv, dis, pat, norm = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
ia, iall, iex = hdr.index('Address'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
base = int(data[0][ia], 16)
per_off = {int(r[ia], 16) - base: (int(r[iex]), int(r[iall])) for r in data}
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and pat in l)
cur, agg = None, collections.defaultdict(lambda: [0, 0, 0])
for l in lines[start + 1:]:
    if l.startswith('//-------') or l.startswith('.text.'):
        break
    m = re.search(r'## File "(.*?)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', l)
    if m and cur:
        off = int(m.group(1), 16)
        e, s = per_off.get(off, (0, 0))
        a = agg[cur]
        a[0] += e; a[1] += s; a[2] += 1
tot_e = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(f'total exec/norm = {tot_e / norm:.1f}, samples = {tot_s}')
srcs = {}
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:70]:
    if f not in srcs:
        try: srcs[f] = open('pjz_b200/csrc/' + f).read().split('\n')
        except Exception: srcs[f] = []
    text = srcs[f][ln - 1].strip()[:78] if ln - 1 < len(srcs[f]) else ''
    print(f'{a[0] / norm:7.1f} {100 * a[1] / tot_s:5.1f}%  {f}:{ln}: {text}')
