"""Builds `libcloudsc2_b200.so` in-tree with nvcc for sm_100a (`python -m cloudsc2_b200.build`)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.normpath(os.path.join(HERE, "..", "csrc"))
OUT = os.path.join(HERE, "libcloudsc2_b200.so")
SOURCES = ["cs2_kernels.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh")) + [os.path.join("..", "..", "include", "cloudsc2_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False, experiments: bool = False, out: str = OUT, defines=()) -> str:
    """`experiments`: also compile the measured-and-rejected kernel variants of csrc/experiments/ (selected at run time with
    CS2_NL_PIPE / CS2_NL_SPLIT / CS2_NL_BULK; profiles/README.md).  The shipped library is built without them."""
    if not force and out == OUT and up_to_date():
        return OUT
    cmd = [nvcc(), *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), *(["-DCS2_EXPERIMENTS"] if experiments else []), *[f"-D{d}" for d in defines],
           "-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    variant = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    if variant:  # A/B build with extra -D flags: --variant=NAME:FLAG1,FLAG2 -> build/libcloudsc2_b200_NAME.so (select with CS2_LIB)
        os.makedirs(os.path.join(HERE, "..", "..", "build"), exist_ok=True)
        for v in variant:
            name, _, flags = v.partition(":")
            print(build(force=True, verbose="-v" in sys.argv, defines=[f for f in flags.split(",") if f],
                        out=os.path.normpath(os.path.join(HERE, "..", "..", "build", f"libcloudsc2_b200_{name}.so"))))
    elif "--experiments" in sys.argv:  # a second library for A/B runs: CS2_LIB=build/libcloudsc2_b200_experiments.so
        os.makedirs(os.path.join(HERE, "..", "..", "build"), exist_ok=True)
        print(build(force=True, verbose="-v" in sys.argv, experiments=True,
                    out=os.path.normpath(os.path.join(HERE, "..", "..", "build", "libcloudsc2_b200_experiments.so"))))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
