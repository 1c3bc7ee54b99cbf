"""Make the reference tree available where /root/reference is not mounted (the GPU box): copy its Python sources and configs
into baseline/_ref/ucdir_reference/ -- git-ignored (never part of this repo's history), not gpurun-ignored (travels with the
snapshot like the built .so).  This is the "install" of the reference: it is not a pip package (no setup.py / pyproject), so
`pip install --target baseline/_ref /root/reference` has nothing to build; a file copy is the whole installation.
Used by tests/test_gpu_reference_caller.py to drive the reference's OWN caller (model/model.py DDPM, sr.py) against
ucdir_b200.define_G.  Run by __graft_entry__.build() when /root/reference exists."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("UCDIR_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref", "ucdir_reference")
KEEP = ("model", "core", "data", "utils", "config", "metric", "sr.py", "eval1.py", "LICENSE", "README.md", "requirements.txt")


def vendor(force=False):
    if not os.path.isdir(SRC):
        return None
    if os.path.isdir(DST) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for name in KEEP:
        s = os.path.join(SRC, name)
        if os.path.isdir(s):
            shutil.copytree(s, os.path.join(DST, name), ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.png", "*.jpg"))
        elif os.path.isfile(s):
            shutil.copy2(s, os.path.join(DST, name))
    return DST


if __name__ == "__main__":
    print(vendor(force="--force" in sys.argv))
