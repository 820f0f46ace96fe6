"""Builds zquatev_b200/lib/libzquatev_b200.so (nvcc, sm_100a) in-tree.  No GPU needed to build."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libzquatev_b200.so")


def build(verbose: bool = False, force: bool = False) -> str:
    cmd = ["make", "-C", CSRC, "-j", str(os.cpu_count() or 4)]
    if force:
        subprocess.check_call(["make", "-C", CSRC, "clean"], stdout=subprocess.DEVNULL)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc build of libzquatev_b200.so failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
