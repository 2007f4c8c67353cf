"""Instruction-mix summary of a kernel's SASS (cuobjdump -sass): total and largest loop bodies."""
import collections, re, subprocess, sys
so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n', 1)[0]
    if pat not in name:
        continue
    ins = re.findall(r'^\s+/\*([0-9a-f]{4,6})\*/\s+((?:@!?U?P[0-9T]+ )?)([A-Z0-9_]+)([.\w]*)[\s;]', f, re.M)
    print(name[:80], 'total', len(ins))
    br = [(int(a, 16), l) for a, l in re.findall(r'^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?BRA.*?);', f, re.M)]
    back = []
    for a, l in br:
        m = re.search(r'0x([0-9a-f]+)', l)
        if m and int(m.group(1), 16) < a:
            back.append((a, int(m.group(1), 16)))
    back.sort(key=lambda x: x[0] - x[1], reverse=True)
    for a, t in back[:4]:
        body = [i for i in ins if t <= int(i[0], 16) <= a]
        print('  loop', hex(t), '->', hex(a), len(body), 'instr',
              collections.Counter(i[2] for i in body).most_common(40))
