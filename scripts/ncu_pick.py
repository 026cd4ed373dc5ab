#!/usr/bin/env python
"""Print selected metrics of an ncu report: ncu_pick.py file.ncu-rep regex [regex ...] (last kernel instance)."""
import csv, re, subprocess, sys
rep, pats = sys.argv[1], [re.compile(p) for p in sys.argv[2:]]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("#", r[hdr.index("Kernel Name")][:60] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, r):
        if any(p.search(h) for p in pats):
            print("  %-80s %-10s %s" % (h, u, v))
