"""SASS evidence for one kernel of libb200fdtd.so: opcode inventory of the whole function, the
memory / async-copy / cache-control / synchronisation mnemonics that characterise the design, and
an excerpt of the largest loop (the plane loop).   usage: sass_excerpt.py LIB MANGLED_SUBSTR [lines]"""
import collections, re, subprocess, sys
so, pat = sys.argv[1], sys.argv[2]
nlines = int(sys.argv[3]) if len(sys.argv) > 3 else 160
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEY = ("LDGSTS", "LDGDEPBAR", "DEPBAR", "UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS",
       "CCTL", "LDG", "STG", "LDS", "STS", "RED", "ATOMG", "ATOMS", "MEMBAR", "ERRBAR", "SHFL", "BAR",
       "NANOSLEEP", "LDC", "LDCU", "FFMA", "FADD", "FMUL", "HADD2", "F2FP", "IMAD", "UTCHMMA", "HMMA")
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n', 1)[0].strip()
    if pat not in name:
        continue
    lines = f.split('\n')
    ins = [(int(m.group(1), 16), m.group(3) + m.group(4), l) for l in lines
           for m in [re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+((?:@!?U?P[0-9T]+ )?)([A-Z0-9_]+)([.\w]*)', l)] if m]
    print(f"Function: {name}\ninstructions: {len(ins)}")
    full = collections.Counter(i[1].split('.')[0] for i in ins)
    print("opcode inventory:", ", ".join(f"{k} {v}" for k, v in full.most_common()))
    variants = collections.Counter(i[1] for i in ins if i[1].split('.')[0] in KEY)
    print("\nmemory / async / cache-control / sync mnemonics (with modifiers):")
    for k in KEY:
        vs = [(n, c) for n, c in variants.items() if n.split('.')[0] == k]
        if vs:
            print(f"  {k:10s} " + ", ".join(f"{n} x{c}" for n, c in sorted(vs, key=lambda t: -t[1])[:8]))
    for absent in ("UTMALDG", "UTMASTG", "UTCHMMA", "HMMA", "UBLKCP"):
        if not any(i[1].startswith(absent) for i in ins):
            print(f"  {absent:10s} (none)")
    back = []
    for a, op, l in ins:
        if op.startswith("BRA"):
            m = re.search(r'0x([0-9a-f]+)', l.split('BRA', 1)[1])
            if m and int(m.group(1), 16) < a:
                back.append((a - int(m.group(1), 16), int(m.group(1), 16), a))
    back.sort(reverse=True)
    if back:
        _, t, a = back[0]
        body = [i for i in ins if t <= i[0] <= a]
        print(f"\nlargest loop: {hex(t)} .. {hex(a)}, {len(body)} instructions (static; cold paths "
              f"-- snapshots, plane source, spin slow paths -- included); first {nlines}:")
        for i in body[:nlines]:
            print(re.sub(r'\s+/\* 0x[0-9a-f]+ \*/\s*$', '', i[2]).rstrip())
    break
