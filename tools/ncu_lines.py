#!/usr/bin/env python
"""
Attribute an ncu SASS-level source page to CUDA source lines.

    ncu -i prof.ncu-rep --page source --csv > sass.csv
    python tools/ncu_lines.py sass.csv path/to/lib.so [top_n]

Uses `cuobjdump -xelf` + `nvdisasm -g` (needs -lineinfo at compile time) to map each SASS
offset of the profiled kernel to file:line (innermost inlined location), then sums
"Instructions Executed" and stall samples per line.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    sass_csv, lib = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rows = list(csv.reader(open(sass_csv)))
    kname = rows[0][1]
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    # mangled-name-free matching: use template args to find the function section
    tmpl = re.search(r"<(.*?)>\(", kname)
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp,
                          stdout=subprocess.DEVNULL)
    # one cubin per translation unit: disassemble all of them (the kernel is matched by name below)
    dis = ""
    for f in sorted(os.listdir(tmp)):
        if f.endswith(".cubin"):
            dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    demangled = subprocess.run(["cu++filt"], input=dis, capture_output=True, text=True).stdout
    want = kname.replace("(bool)0", "false").replace("(bool)1", "true")
    want = re.sub(r"\(int\)", "", want).replace("void ", "").split("(")[0]
    want = re.sub(r"\s+", "", want)
    line_of = {}
    cur_fn, cur_loc, active = None, None, False
    fn_re = re.compile(r"^\s*\.text\.(\S+):")
    for raw, dem in zip(dis.splitlines(), demangled.splitlines()):
        m = fn_re.match(raw)
        if m:
            active = dem.strip().rstrip(":").replace(".text.", "", 1) == kname
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', raw)
        if m:
            cur_loc = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", raw)
        if m and active:
            line_of[int(m.group(1), 16)] = cur_loc
    if not line_of:
        print("could not match kernel section; functions seen differ from", want)
        sys.exit(1)
    agg = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    base = None
    for r in data:
        try:
            addr = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]])
        except ValueError:
            continue
        if base is None:
            base = addr
        off = addr - base
        loc = line_of.get(off, ("?", 0))
        ins = int(float(r[col["Instructions Executed"]] or 0))
        smp = int(float(r[col["Warp Stall Sampling (All Samples)"]] or 0))
        thr = int(float(r[col["Thread Instructions Executed"]] or 0))
        for k, v in enumerate((ins, smp, thr)):
            agg[loc][k] += v
            tot[k] += v
    print(f"kernel: {kname}\ntotal warp-instr {tot[0]:,}  stall samples {tot[1]:,}  avg active threads "
          f"{tot[2] / max(tot[0], 1):.1f}")
    print(f"{'file:line':32s} {'instr%':>7s} {'samples%':>9s} {'thr/inst':>8s}")
    for loc, (ins, smp, thr) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{loc[0] + ':' + str(loc[1]):32s} {100 * ins / max(tot[0], 1):7.2f} "
              f"{100 * smp / max(tot[1], 1):9.2f} {thr / max(ins, 1):8.1f}")


if __name__ == "__main__":
    main()
