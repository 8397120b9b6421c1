"""Recipe for ``oracle/_ref``: a verbatim, git-ignored copy of the reference's model package.  TEST / BASELINE INFRASTRUCTURE ONLY.

    python -m oracle.fetch_ref            (run by __graft_entry__.build() whenever /root/reference is present)

The reference is pure Python (no build step): "compiling" it is copying ``experiments/model/`` (and the two data helpers the
parity tests mirror) next to the oracle so that it travels to the GPU box with the repository snapshot -- ``oracle/_ref/`` is listed
in .gitignore (reference sources never enter the history) but not in .gpurunignore.  Consumers: tests/ (the reference's own
ODEGPVAE / VAE / compute_loss with the drop-in SVGP_Layer / Flow swapped in; the live reference as checker) and the CPU-baseline
legs of bench.py (``cpu_baseline.kind = "reference"``).  Nothing under vae-gp-ode_b200/ imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC_DEFAULT = "/root/reference"
WHAT = ["experiments/__init__.py", "experiments/model", "experiments/data/__init__.py", "experiments/data/utils.py"]


def fetch(src=SRC_DEFAULT, dst=DST):
    if not os.path.isdir(os.path.join(src, "experiments", "model", "core")):
        return False
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    for rel in WHAT:
        s, d = os.path.join(src, rel), os.path.join(dst, rel)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.ipynb"))
        else:
            shutil.copy2(s, d)
    with open(os.path.join(dst, "ORIGIN"), "w") as f:
        f.write("verbatim copy of %s (%s) made by oracle/fetch_ref.py; not part of the repository history\n" % (src, ", ".join(WHAT)))
    return True


if __name__ == "__main__":
    ok = fetch(sys.argv[1] if len(sys.argv) > 1 else SRC_DEFAULT)
    print("oracle/_ref %s" % ("written" if ok else "NOT written: reference tree absent"))
