"""Print the SASS of the innermost loop that contains the first SHFL.UP of a kernel (the NW step loop)."""
import re, sys
txt = open(sys.argv[1]).read()
kern = sys.argv[2]
full = len(sys.argv) > 3
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n')[0]
    if kern not in name: continue
    ins = [l for l in f.split('\n') if re.match(r'\s*/\*[0-9a-f]{4}\*/', l)]
    addr = lambda l: int(re.match(r'\s*/\*([0-9a-f]{4})\*/', l).group(1), 16)
    sh = [addr(l) for l in ins if 'SHFL.UP' in l]
    best = None
    for l in ins:
        m = re.search(r'BRA\s+.*?(0x[0-9a-f]+)', l)
        if m and int(m.group(1), 16) < addr(l) and sh and int(m.group(1), 16) <= sh[0] <= addr(l):
            if best is None or addr(l) - int(m.group(1), 16) < best[1] - best[0]: best = (int(m.group(1), 16), addr(l))
    print(name[:50], 'total', len(ins), 'loop', best and (best[1] - best[0]) // 16 + 1)
    if full and best:
        for l in ins:
            if best[0] <= addr(l) <= best[1]: print(re.sub(r'/\* 0x[0-9a-f]+ \*/', '', l).rstrip()[8:])
