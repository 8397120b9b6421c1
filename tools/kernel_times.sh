#!/bin/bash
# per-kernel durations of one bench step (ncu launch list, serialized, cold-cache): tools/kernel_times.sh <workload> <out.csv>
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_r|k_df|k_field" -c 40 --csv --log-file "$2" python bench.py --workload "$1" --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - "$2" <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0][:60]
    agg.setdefault(name, []).append(float(r[-1]) / 1e6)
for k, v in agg.items():
    print("%-62s n=%d  mean %.3f ms" % (k, len(v), sum(v) / len(v)))
PY
