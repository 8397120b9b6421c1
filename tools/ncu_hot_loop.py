#!/usr/bin/env python
"""Per-instruction warp-stall samples of the hottest loop of one kernel, from `ncu -i rep --page source --csv --kernel-name regex:NAME`.
Usage: python tools/ncu_hot_loop.py source_page.csv [min_fraction_of_max_exec]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.9
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ix = {k: i for i, k in enumerate(hdr)}
body = [r for r in rows[h + 1:] if len(r) > 10 and r[0].startswith("0x")]
num = lambda r, k: int(float(r[ix[k]] or 0))
tot = sum(num(r, "# Samples") for r in body)
mx = max(num(r, "Instructions Executed") for r in body)
hot = [r for r in body if num(r, "Instructions Executed") >= mx * frac]
print("total samples %d, loop instructions %d, samples in loop %d, executions of the loop body %d" % (tot, len(hot), sum(num(r, "# Samples") for r in hot), mx))
keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(num(r, k) for r in hot) for k in keys}
print("loop stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(1, sum(agg.values()))) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
for r in hot:
    s = num(r, "# Samples")
    st = " ".join("%s=%d" % (k[6:], num(r, k)) for k in keys if s and num(r, k) > 0.2 * s)
    print("%6d  %-72s %s" % (s, r[ix["Source"]].strip()[:72], st))
