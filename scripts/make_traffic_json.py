#!/usr/bin/env python
"""profiles/<tag>_traffic.json: ncu dram__bytes_read/write (and duration) per kernel, first captured
launch of each kernel in profiles/<tag>_full_*.raw.csv.  bench.py reports these as roofline.traffic."""
import csv
import glob
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0,
        "msecond": 1e3, "nsecond": 1e-3}


def main(tag):
    table = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", f"{tag}_full_*.raw.csv"))):
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}

        def val(r, m):
            return float(r[col[m]].replace(",", "")) * UNIT[units[col[m]]]
        for r in rows[2:]:
            name = re.sub(r"<.*", "", r[col["Kernel Name"]].split("(")[0]).split("::")[-1].replace("void ", "").strip()
            if name in table:
                continue
            table[name] = {"dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
                           "us": val(r, "gpu__time_duration.sum"), "capture": os.path.basename(path)}
    out = os.path.join(ROOT, "profiles", f"{tag}_traffic.json")
    json.dump(table, open(out, "w"), indent=1)
    print(json.dumps({k: round(v["dram_read_bytes"] + v["dram_write_bytes"]) for k, v in table.items()}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
