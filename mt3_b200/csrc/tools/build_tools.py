"""Builds the standalone kernel bring-up binaries next to this file (sm_100a)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TOOLS = ["gemm_tc_test", "attn_tc_test"]


def main():
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    for t in TOOLS:
        src, out = os.path.join(HERE, t + ".cu"), os.path.join(HERE, t)
        deps = [src] + [os.path.join(HERE, "..", f) for f in os.listdir(os.path.join(HERE, "..")) if f.endswith((".cuh", ".cu"))]
        if os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(d) for d in deps):
            continue
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "--expt-relaxed-constexpr", "-o", out, src]
        subprocess.run(cmd, check=True)
        print("built", out)


if __name__ == "__main__":
    sys.exit(main())
