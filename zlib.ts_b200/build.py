"""Builds the native libraries in-tree (sm_100a only; nvcc cross-compiles without a GPU).

    libzlibts_b200.so  -- the C-ABI engine (include/zlibts_b200.h): CUDA kernels + host glue
    libzts_synth.so    -- synthetic input generators, plain C
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libzlibts_b200.so")
SYNTH_LIB = os.path.join(HERE, "libzts_synth.so")

CU_SOURCES = ["zts_ctx.cu", "zts_hoststage.cu", "zts_checksum.cu", "zts_inflate.cu", "zts_lz77.cu", "zts_huffman.cu", "zts_deflate.cu",
              "zts_container.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
] + (["-DZTS_USE_MATCH"] if os.environ.get("ZTS_USE_MATCH") else []) + os.environ.get("ZTS_NVCC_EXTRA", "").split() + [
    "-Xcompiler", "-fPIC", "-shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_synth(force=False):
    """libzts_synth.so alone: plain-C generators, no CUDA (bench.py's reference arm builds nothing else of this package)."""
    synth_src = os.path.join(CSRC, "zts_synth.c")
    if force or not _newer(SYNTH_LIB, [synth_src]):
        gcc = shutil.which("gcc") or "gcc"
        tmp = SYNTH_LIB + ".tmp%d" % os.getpid()
        subprocess.run([gcc, "-O2", "-fPIC", "-shared", "-std=c11", "-o", tmp, synth_src], check=True)
        os.replace(tmp, SYNTH_LIB)
    return SYNTH_LIB


def build(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "zlibts_b200.h"))
    if force or not _newer(LIB, deps):
        tmp = LIB + ".tmp%d" % os.getpid()  # built aside and renamed: concurrent ranks never see a partial file
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + srcs
        subprocess.run(cmd, check=True, cwd=CSRC)
        os.replace(tmp, LIB)
    build_synth(force)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
