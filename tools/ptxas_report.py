"""Compact register / spill table of every kernel (nvcc -Xptxas -v), no GPU needed."""
import re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "microtipi_b200", "csrc", "wfm_api.cu")
out = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                      "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"] + os.environ.get("WFM_BUILD_FLAGS", "").split() + [src, "-o", "/tmp/_ptxas_report.so"],
                     capture_output=True, text=True).stderr
pat = sys.argv[1] if len(sys.argv) > 1 else "pipeline"
cur = None
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\w+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        spill = None
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur:
        spill = m.groups()
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        if re.search(pat, cur):
            print(f"{cur:55s} regs={m.group(1):>3s} stack={spill[0]:>4s} spill_st={spill[1]:>4s} spill_ld={spill[2]:>4s}")
        cur = None
