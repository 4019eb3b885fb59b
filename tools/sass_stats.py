#!/usr/bin/env python
"""Static SASS statistics of expand_kernel in an object file: instruction counts of the kernel body and of each called
routine (emit, apply<G>, flush), local-memory instructions, and proof-of-feature mnemonics.  usage: sass_stats.py obj.o"""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if "expand_kernel" not in name:
        continue
    ins = re.findall(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", f)
    addr = [(int(a, 16), t.strip()) for a, t in ins]
    targets = sorted({int(m.group(1), 16) for _, t in addr for m in [re.search(r"CALL\.\w+(?:\.\w+)*\s+(?:.*?)(0x[0-9a-f]+)", t)] if m})
    bounds = [0] + targets + [addr[-1][0] + 16]
    print(name[:90])
    for lo, hi in zip(bounds, bounds[1:]):
        sel = [t for a, t in addr if lo <= a < hi]
        loc = sum(1 for t in sel if re.match(r"(@!?U?P\d\s+)?(LDL|STL)", t))
        print(f"  region {lo:#7x}: {len(sel):5d} instr, local ld/st {loc}")
    allt = " ".join(t for _, t in addr)
    print("  total", len(addr), {k: len(re.findall(k, allt)) for k in ["UBLKCP", "ATOMG.E.CAS.128", "BAR.SYNC", "ATOMS", "SHFL", "VOTE", "LDL", "STL"]})
