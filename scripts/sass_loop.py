"""Opcode histogram of the SASS between two addresses of one function of an object file.
   usage: sass_loop.py <object> <function substring> <lo hex> <hi hex> [print]"""
import re, subprocess, sys
from collections import Counter
obj, fun, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
on = False
c = Counter(); n = 0
for l in txt.splitlines():
    if "Function :" in l:
        on = fun in l
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
    if on and m and lo <= int(m.group(1), 16) <= hi:
        p = m.group(2).split()
        c[p[1] if p[0].startswith("@") else p[0]] += 1; n += 1
        if len(sys.argv) > 5: print(m.group(1), m.group(2))
print(n, "instructions")
for k, v in c.most_common(): print(f"{v:5d} {k}")
