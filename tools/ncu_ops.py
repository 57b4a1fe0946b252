#!/usr/bin/env python
"""Per-source-line opcode histogram of an ncu SASS source page (see tools/ncu_lines.py)."""
import csv, re, subprocess, os, tempfile, sys
from collections import defaultdict, Counter
sass_csv, lib = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = list(csv.reader(open(sass_csv))); kname = rows[0][1]; hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
dem = subprocess.run(["cu++filt"], input=dis, capture_output=True, text=True).stdout
line_of = {}; active = False; cur = None
for raw, d in zip(dis.splitlines(), dem.splitlines()):
    m = re.match(r"^\s*\.text\.(\S+):", raw)
    if m:
        active = d.strip().rstrip(":").replace(".text.", "", 1) == kname
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', raw)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", raw)
    if m and active:
        line_of[int(m.group(1), 16)] = cur
base = None; per = defaultdict(Counter); tot = Counter(); ops = Counter(); thr = Counter()
for r in data:
    a = r[col["Address"]]; addr = int(a, 16) if a.startswith("0x") else int(a)
    if base is None: base = addr
    loc = line_of.get(addr - base, ("?", 0)); n = int(float(r[col["Instructions Executed"]] or 0))
    src = r[col["Source"]]; op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
    per[loc][op] += n; tot[loc] += n; ops[op] += n
    thr[loc] += int(float(r[col["Thread Instructions Executed"]] or 0))
norm = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
total = sum(tot.values())
print(f"total {total / norm:.1f}; opcode mix:", ", ".join(f"{k}:{v / norm:.0f}" for k, v in ops.most_common(24)))
for loc, n in tot.most_common(top):
    print(f"{loc[0]}:{loc[1]:<5d} {n / norm:8.1f} thr {thr[loc] / max(n, 1):4.1f} ",
          ", ".join(f"{k}:{v / norm:.0f}" for k, v in per[loc].most_common(7)))
