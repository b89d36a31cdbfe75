"""Write the SASS of the two hot loops to profiles/ (from the object files of the current build):
the scan kernel's row-pair loop and the frame-evaluation kernel's two per-bit block loops."""
import re, subprocess, sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"


def sass(obj, fun=None):
    cmd = ["cuobjdump", "-sass"] + (["-fun", fun] if fun else []) + [str(ROOT / "build" / obj)]
    txt = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    ins = []
    for l in txt.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip(), re.sub(r"\s+/\* 0x[0-9a-f]+ \*/$", "", l)))
    return ins


def loops(ins):
    out = []
    for a, t, _ in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            out.append((int(m.group(1), 16), a, (a - int(m.group(1), 16)) // 16 + 1))
    return out


def histogram(ins, lo, hi):
    c = Counter()
    for a, t, _ in ins:
        if lo <= a <= hi:
            p = t.split()
            c[(p[1] if p[0].startswith("@") else p[0])] += 1
    return ", ".join(f"{k} {v}" for k, v in c.most_common(14))


k1 = sass("modes_scan2.o")
lo, hi, n = next(x for x in loops(k1) if x[2] > 500)
head = (f"# scan_kernel, the row-pair loop: {n} SASS instructions in the loop body (the slow path of the first/last tile\n"
        f"# included; 64 positions per lane and trip on the fast path).  sm_100a, nvcc 12.9.\n# opcodes: {histogram(k1, lo, hi)}\n")
(ROOT / "profiles" / f"{tag}_sass_k1_row_loop.txt").write_text(head + "\n".join(l for a, _, l in k1 if lo <= a <= hi) + "\n")

k2 = sass("modes_kernels.o", "_ZN5modes18eval_serial_kernelILb1EEEvNS_9BatchViewENS_12DeviceTablesEPKjPjjP15modes_candidateii")
big = [x for x in loops(k2) if 200 < x[2] < 500][:2]
parts = []
for name, (lo, hi, n) in zip(["first pass, one block of 16 bits (dump1090.c:1667-1690)",
                              "phase-corrected retry, one block of 16 bits (dump1090.c:1498-1558 and the slicing of the corrected samples)"], big):
    parts.append(f"# eval_serial_kernel<lean>, {name}: {n} instructions = {n / 16:.1f} per bit\n# opcodes: {histogram(k2, lo, hi)}\n"
                 + "\n".join(l for a, _, l in k2 if lo <= a <= hi) + "\n")
(ROOT / "profiles" / f"{tag}_sass_k2_bit_loops.txt").write_text("\n".join(parts))
print("written", tag)
