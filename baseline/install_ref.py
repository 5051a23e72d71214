#!/usr/bin/env python
"""Installs the UNMODIFIED reference (frotaur/PyCA) into baseline/_ref/ (git-ignored, travels to the GPU box).

    python baseline/install_ref.py [--ref /root/reference]

1. The sanctioned route: `pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target
   baseline/_ref <copy of the reference under /tmp>`.  In this image it fails: the reference's build backend (`hatchling`,
   pyproject.toml:1-3) is neither installed nor in /opt/wheelhouse.
2. Fallback, equivalent for this project: its wheel is pure Python with `[tool.hatch.build.targets.wheel] packages =
   ["pyca"]` (pyproject.toml:38-39), i.e. exactly the `pyca/` package tree -- so that tree is laid down as pip would, plus
   `demo_data/` (parameter files and NCA checkpoints the public API loads, interface/defaults.py:10-25).
Nothing is edited; bench.py imports it through baseline/ref_loader.py (inert stand-ins for the GUI packages the image lacks).
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    if not os.path.isdir(os.path.join(a.ref, "pyca")):
        print(f"no reference tree at {a.ref}: nothing installed")
        return 0
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    how = None
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(a.ref, src, ignore=shutil.ignore_patterns(".git", "__pycache__"))
        r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
                            "/opt/wheelhouse", "--target", DEST, src], capture_output=True, text=True)
        if r.returncode == 0 and os.path.isdir(os.path.join(DEST, "pyca")):
            how = "pip install --no-deps --target baseline/_ref"
        else:
            why = (r.stderr.strip().splitlines() or ["?"])[-1]
            os.makedirs(DEST, exist_ok=True)
            shutil.copytree(os.path.join(src, "pyca"), os.path.join(DEST, "pyca"), ignore=shutil.ignore_patterns("__pycache__"))
            how = f"pure-Python wheel contents laid down by hand (pip failed: {why})"
        shutil.copytree(os.path.join(src, "demo_data"), os.path.join(DEST, "demo_data"))
    with open(os.path.join(DEST, "INSTALL.txt"), "w") as f:
        f.write(how + "\n")
    print("installed:", how)
    return 0


if __name__ == "__main__":
    sys.exit(main())
