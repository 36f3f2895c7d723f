#!/usr/bin/env python
"""Per-source-line instruction / stall-sample profile of one kernel from an ncu report.

    python tools/sass_profile.py <report.ncu-rep> <kernel-regex> [lib.so]

Joins ncu's SASS page (instructions executed and stall samples per instruction) with the line table
of the library (nvdisasm -g) by instruction offset and prints the totals per source line."""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(lib, mangled_re):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    out = "".join(subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
                  for cubin in sorted(glob.glob(os.path.join(tmp, "*.cubin"))))
    table, cur, active = {}, None, False
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            active = re.search(mangled_re, m.group(1)) is not None
            if active and table:
                break
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "fcfv_nme23_b200", "libfcfv_b200.so")
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kname = rows[0][1]
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    mangled = {"k_asm_fill": "k_asm_fillILi1ELi1ELi0ELb0EiE", "k_asm_count": "k_asm_countILi1ELi1ELi0ELb0E",
               "k_res_elem": "k_res_elemILi1E", "k_local": "k_localILi1E"}.get(kre, kre)
    table = sass_lines(lib, mangled)
    base = None
    per_line = defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    for r in rows[2:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        addr = int(r[0], 16)
        if base is None:
            base = addr
        off = addr - base
        ie = int(float(r[ci["Instructions Executed"]] or 0))
        ss = int(float(r[ci["# Samples"]] or 0))
        line = table.get(off, (("?", 0), ""))[0] or ("?", 0)
        per_line[line][0] += ie
        per_line[line][1] += ss
        per_line[line][2] += 1
        tot_i += ie
        tot_s += ss
    print(f"# {kname[:100]}\n# total warp instructions {tot_i}, stall samples {tot_s}")
    print(f"{'file:line':34s} {'sass':>5s} {'warp_insts':>12s} {'%':>6s} {'samples':>8s} {'%':>6s}")
    for line, (ie, ss, n) in sorted(per_line.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        if ie == 0 and ss == 0:
            continue
        print(f"{line[0] + ':' + str(line[1]):34s} {n:5d} {ie:12d} {100.0 * ie / max(tot_i, 1):6.2f} {ss:8d} {100.0 * ss / max(tot_s, 1):6.2f}")


if __name__ == "__main__":
    main()
