#!/usr/bin/env python
"""Per-kernel SASS evidence of the Blackwell-native paths: counts of tcgen05 MMA (UTC*MMA), tensor-memory loads (LDTM), bulk copies
(UBLKCP), legacy warp-level MMA (HMMA), double-precision FMA (DFMA) and MUFU in every kernel of the given object files.
Usage: python tools/sass_summary.py vae-gp-ode_b200/csrc/build/rbf_inst_16.o [more objects] > profiles/sass_rNN.txt"""
import re
import subprocess
import sys

PAT = [("UTCMMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("UBLKCP", r"\bUBLKCP"), ("HMMA", r"\bHMMA"), ("DFMA", r"\bDFMA"), ("MUFU", r"\bMUFU"),
       ("SYNCS", r"\bSYNCS"), ("STS", r"\bSTS"), ("RED", r"\bRED\b|\bREDG")]
for obj in sys.argv[1:]:
    print("== %s" % obj)
    sass = subprocess.check_output(["cuobjdump", "-sass", obj]).decode()
    names = subprocess.check_output(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)).encode()).decode().splitlines()
    blocks = re.split(r"Function : \S+", sass)[1:]
    for name, body in zip(names, blocks):
        inst = [ln for ln in body.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln)]
        cnt = {k: sum(1 for ln in inst if re.search(p, ln)) for k, p in PAT}
        name = name.replace("gpode::", "").replace("(anonymous namespace)::", "")
        name = re.sub(r"\(.*", "", name)
        print("  %-58s insts %6d  %s" % (name[:58], len(inst), "  ".join("%s %d" % (k, v) for k, v in cnt.items() if v)))
