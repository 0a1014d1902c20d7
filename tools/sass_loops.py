"""List the loops (backward branches) of one kernel's SASS with their size and shuffle / local-memory content.
usage: cuobjdump -sass kmagpu_align.o > a.sass; python tools/sass_loops.py a.sass nw_batch_kernel [min_instr]"""
import re, sys
txt = open(sys.argv[1]).read()
want = sys.argv[2]
minn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n')[0]
    if want not in name:
        continue
    ins = []
    for m in re.finditer(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', f):
        ins.append((int(m.group(1), 16), m.group(2)))
    addr_idx = {a: i for i, (a, _) in enumerate(ins)}
    print(name[:60], len(ins), "instructions")
    for i, (a, t) in enumerate(ins):
        m = re.search(r'\bBRA\b.*?0x([0-9a-f]+)', t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_idx:
            j = addr_idx[tgt]
            body = [x for _, x in ins[j:i + 1]]
            n = len(body)
            if n < minn:
                continue
            cnt = lambda pat: sum(1 for x in body if re.search(pat, x))
            print(f"  loop {tgt:#x}..{a:#x}: {n} instr, SHFL.UP {cnt(r'SHFL.UP')}, SHFL.DOWN {cnt(r'SHFL.DOWN')}, "
                  f"LDL {cnt(r'LDL')}, STL {cnt(r'STL')}, STG {cnt(r'STG|ST.E')}, LDG {cnt(r'LDG|LD.E')}, LDS {cnt(r'LDS')}, BRA {cnt(r'BRA')}")
