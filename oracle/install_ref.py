"""Installs the UNMODIFIED reference into baseline/_ref so that it travels to the GPU box (bench.py --impl
reference and the `tv_cuda` arm import it from there).

The base contract's recipe is `pip install --target baseline/_ref /root/reference`; charles-fox/DGOD has no
setup.py / pyproject.toml (it is a directory of scripts), so the "install" is a byte-for-byte copy of its
Python modules.  baseline/_ref is git-ignored (never committed) and not gpurun-ignored.  Run here:

    python -m oracle.install_ref

TEST / BENCH INFRASTRUCTURE ONLY — nothing under dgod_b200/ imports baseline/_ref.
"""
from __future__ import annotations

import hashlib
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference")
DST = ROOT / "baseline" / "_ref"
MODULES = ["fasterrcnn.py", "fcos.py", "DGcommon.py", "DGFRCNN.py", "DGFCOS.py"]   # the path's modules


def install(verbose: bool = True) -> bool:
    """Copies the reference modules; returns False when /root/reference is not on this machine."""
    if not SRC.exists():
        return False
    DST.mkdir(parents=True, exist_ok=True)
    lines = []
    for name in MODULES:
        shutil.copyfile(SRC / name, DST / name)
        lines.append(f"{hashlib.sha256((DST / name).read_bytes()).hexdigest()}  {name}")
    (DST / "SHA256SUMS").write_text("\n".join(lines) + "\n")
    if verbose:
        print(f"installed {len(MODULES)} reference modules into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
