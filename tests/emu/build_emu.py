"""Build the CPU-emulated copy of the C-ABI library (TEST INFRASTRUCTURE ONLY).

The same sources nvcc compiles for sm_100a are compiled with g++ against tests/emu/cuda_emu.h.
The product package never loads this library."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build", "libwfm_emu.so")
SRCS = [os.path.join(ROOT, "microtipi_b200", "csrc", "wfm_api.cu")]
CSRC = os.path.join(ROOT, "microtipi_b200", "csrc")
DEPS = SRCS + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".inl"))] + [
    os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "wfm_b200.h")]


def build(force=False, sanitize=False):
    out = OUT.replace(".so", "_asan.so") if sanitize else OUT
    extra = os.environ.get("WFM_EMU_FLAGS", "").split()       # experiment knobs (-DWFM_...=...): a separate library
    if extra:
        out = out.replace(".so", "_" + "".join(ch if ch.isalnum() else "_" for ch in "".join(extra)) + ".so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in DEPS):
        return out
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-fno-strict-aliasing", "-fvisibility=hidden", "-Wl,-Bsymbolic",
           "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-unused-variable",
           "-DWFM_EMU", "-include", os.path.join(HERE, "cuda_emu.h")] + extra + ["-x", "c++"] + SRCS + \
          ["-o", out, "-lpthread"]
    if sanitize:
        cmd[1:1] = ["-fsanitize=address", "-fno-omit-frame-pointer"]
        cmd[cmd.index("-O2")] = "-O1"          # the instrumented build is slow to compile at -O2
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, sanitize="--asan" in sys.argv))
