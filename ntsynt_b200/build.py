"""Build libntsynt_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libntsynt_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--extended-lambda",
    "-shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-ccbin", "/usr/bin/g++",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    "Compile every .cu under csrc/ into one shared library; returns its path."
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, *sources(), "-ldl"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose:
        sys.stderr.write(res.stdout)
    elif res.returncode != 0:
        sys.stderr.write("\n".join(l for l in res.stdout.splitlines() if "ptxas info" not in l and "bytes stack" not in l
                                   and "Compile time" not in l and "registers" not in l) + "\n")
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libntsynt_b200.so")
    with open(os.path.join(HERE, "build.log"), "w", encoding="utf-8") as fh:
        fh.write(" ".join(cmd) + "\n" + res.stdout)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
