"""Static count of the SASS instructions in the largest backward-branch loop that contains a given
opcode (default MUFU.EX2) of one kernel in an object file: python tools/sass_loop.py obj kernel_substr [opcode]"""
import collections, re, subprocess, sys
obj, ksub = sys.argv[1], sys.argv[2]
needle = sys.argv[3] if len(sys.argv) > 3 else "MUFU.EX2"
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, funcs = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if ksub not in name: continue
    best = None
    for addr, text in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?(0x[0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr:
                body = [(a, t) for a, t in ins if tgt <= a <= addr]
                if any(needle in t for _, t in body) and (best is None or len(body) > len(best)):
                    best = body
    if not best:
        print(name, "no loop with", needle); continue
    ops = collections.Counter((t.split()[1] if t.startswith("@") else t.split()[0]) for _, t in best)
    print(name, "loop instrs:", len(best))
    print("  ", sorted(ops.items(), key=lambda x: -x[1])[:24])
