#!/usr/bin/env python
"""Summarises `ncu --set full` raw-page CSVs into profiles/ncu_traffic.json, the table bench.py's `roofline.traffic` reads:
per workload and kernel family, dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches) of
the family's most frequent launch shape, with the capture file named next to it.

    python tools/ncu_extract.py srgan_c3 umma_conv profiles/ncu_body_r2b_raw.csv "64->64 3x3 forward, 16x96x96" [kernel-name regex] [min us]
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    wl, family, path, label = sys.argv[1:5]
    pat = re.compile(sys.argv[5]) if len(sys.argv) > 5 else None
    min_us = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0      # keeps only launches at least this long (one shape out of a mixed capture)
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {c: i for i, c in enumerate(hdr)}
    tot, n, dur = 0.0, 0, 0.0
    for r in rows[2:]:
        if pat and not pat.search(r[ix["Kernel Name"]]):
            continue
        if float(r[ix["gpu__time_duration.sum"]].replace(",", "")) < min_us:
            continue
        b = 0.0
        for col in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[ix[col]].replace(",", "")) * UNIT[units[ix[col]]]
        tot += b; n += 1
        dur += float(r[ix["gpu__time_duration.sum"]].replace(",", ""))
    assert n, "no launch matched"
    out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    table = json.load(open(out_path)) if os.path.exists(out_path) else {}
    table.setdefault(wl, {})[family] = {"dram_bytes_per_launch": round(tot / n), "launches_averaged": n, "launch": label,
                                        "capture": os.path.relpath(path, ROOT), "ncu_us_per_launch": round(dur / n, 2)}
    json.dump(table, open(out_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(table[wl][family]))


if __name__ == "__main__":
    main()
