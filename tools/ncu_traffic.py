#!/usr/bin/env python
"""DRAM traffic per launch of the hot-path kernels from an ncu report -> small JSON for profiles/.

    python tools/ncu_traffic.py <report.ncu-rep> <nel> > profiles/<tag>_traffic.json

bench.py puts the entry of its dominant kernel into roofline.traffic when `nel` matches the run."""
import csv
import json
import subprocess
import sys


def main():
    rep, nel = sys.argv[1], int(sys.argv[2])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    acc = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].strip()
        rd = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
        a = acc.setdefault(name, [0.0, 0.0, 0])
        a[0] += rd; a[1] += wr; a[2] += 1
    res = {"source": rep.split("/")[-1], "nel": nel, "how": "ncu dram__bytes_read.sum + dram__bytes_write.sum, average per launch",
           "kernels": {k: {"dram_read_bytes": v[0] / v[2], "dram_write_bytes": v[1] / v[2], "dram_bytes": (v[0] + v[1]) / v[2],
                           "launches": v[2]} for k, v in acc.items()}}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
