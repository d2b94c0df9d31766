"""Builds csrc/libnbgrad_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
LIB = os.environ.get("NBGRAD_B200_LIB") or os.path.join(CSRC, "libnbgrad_b200.so")  # env override: A/B builds on the GPU box
SOURCES = ["nbg_b200.cu"]
# No --split-compile: it cuts the build from 2 min to 45 s, but the partitioning changes the register allocation of the hottest kernel from
# build to build (jac_rx_kernel<8,4>: 250 registers / 220 ms per bench window in one build, 204 registers / 237 ms in the next, same source;
# profiles/r02h_ab.jsonl).  NBGRAD_FAST_BUILD=1 turns it on for development builds.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]
if os.environ.get("NBGRAD_FAST_BUILD") == "1":
    NVCC_FLAGS += ["--split-compile", "0"]
# NBGRAD_EXPERIMENTS=1 also compiles the measured-and-rejected kernel variants (DMMA Jacobian kernel, pivot-block / lockstep variants of
# jac_rx_kernel: DESIGN.md 5); they double the build time and are off by default
if os.environ.get("NBGRAD_EXPERIMENTS") == "1":
    NVCC_FLAGS.append("-DNBG_EXPERIMENTS")


def dependencies():
    """Everything the library is compiled from: every .cu / .cuh / .inc in csrc/ and the public header."""
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".inc"))]
    deps.append(os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "nbgrad.h"))
    return deps


def source_hash():
    """sha1 over the sources the library is compiled from: profiles that quote per-kernel numbers record it, and bench.py refuses to
    scale a profile taken from other sources."""
    import hashlib
    h = hashlib.sha1()
    for f in dependencies():
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # the hash of the sources is compiled in (nbg_source_hash): a library that does not match the sources next to it is detected at load
    cmd = [nvcc] + NVCC_FLAGS + ['-DNBG_SRC_HASH="%s"' % source_hash()] + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB + ".tmp"] + SOURCES
    subprocess.run(cmd, cwd=CSRC, check=True)
    os.replace(LIB + ".tmp", LIB)
    return LIB


def built_hash():
    """Source hash compiled into the library on disk, or None."""
    import ctypes
    if not os.path.exists(LIB):
        return None
    try:
        L = ctypes.CDLL(LIB)
        L.nbg_source_hash.restype = ctypes.c_char_p
        return L.nbg_source_hash().decode()
    except (OSError, AttributeError):
        return None


def needs_build():
    return built_hash() != source_hash()


if __name__ == "__main__":
    import sys
    if "--if-stale" in sys.argv:
        print("up to date" if not needs_build() else build(force=True))
    else:
        print(build(force=True, verbose="-v" in sys.argv))
