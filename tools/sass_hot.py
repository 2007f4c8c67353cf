"""Execution-weighted view of an `ncu --page source --csv` dump: opcode mix of what actually runs
(normalised per warp and plane), the instructions that collect the stall samples, and the stall
reasons.   usage: sass_hot.py SOURCE.csv NORM [top]      NORM = warp-plane iterations in the launch"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
norm = float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
ex, smp, src = ix['Instructions Executed'], ix['Warp Stall Sampling (All Samples)'], ix['Source']
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
ops = collections.Counter(); tot = 0; tots = 0
reasons = collections.Counter()
for r in data:
  e = int(r[ex]); s = int(r[smp]); tot += e; tots += s
  m = re.match(r'\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', r[src])
  ops[m.group(1) if m else '?'] += e
  for h in stall_cols:
    reasons[h] += int(r[ix[h]] or 0)
print(f'executed warp-instructions: {tot:.4g} = {tot / norm:.1f} per warp and plane; stall samples {tots}')
print('opcode mix (per warp and plane):', ', '.join(f'{o} {c / norm:.1f}' for o, c in ops.most_common(45)))
print('stall reasons:', ', '.join(f'{h[6:]} {100 * c / max(tots, 1):.1f}%' for h, c in reasons.most_common(12)))
print('--- top instructions by stall samples (samples, share, exec/norm, dominant reasons, SASS)')
for r in sorted(data, key=lambda r: -int(r[smp]))[:top]:
  rs = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
  print(f'{int(r[smp]):8d} {100 * int(r[smp]) / tots:5.2f}% {int(r[ex]) / norm:6.2f}  '
        f'{rs[0][1]}:{rs[0][0]} {rs[1][1]}:{rs[1][0]}  {r[src].strip()[:90]}')
