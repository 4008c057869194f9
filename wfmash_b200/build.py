"""Builds libwfmash_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", f) for f in ("wfa_host.cu", "sketch.cu", "minmer_host.cu", "index_host.cu", "epilogue.cu", "chain_host.cu", "filter_host.cu", "stats_host.cu", "ani_host.cu", "phases_host.cu", "index_file_host.cu")]
HDR = [os.path.join(HERE, "csrc", f) for f in ("wfa_kernels.h", "wfb_rt.h", "minmer_kernels.h", "sketch_kernels.h", "index_kernels.h", "l2_kernels.h", "ani_kernels.h")] + [os.path.join(ROOT, "include", "wfmash_b200.h")]
OUT = os.path.join(HERE, "libwfmash_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(f) > t for f in SRC + HDR if os.path.exists(f))


# tuned defaults of the breakpoint kernel (see DESIGN.md / profiles/): __launch_bounds__ and call style
DEFAULT_DEFS = ["-DWFB_BREAK_MAXTHREADS=256", "-DWFB_BREAK_MINBLOCKS=2", "-DWFB_PREFETCH_DIST=0"]


def build_variant(out, defs, verbose=False):
    """Build the library with other compile-time knobs (used by tuning sweeps only)."""
    srcs = [s for s in SRC if os.path.exists(s)]
    cmd = [nvcc_path()] + NVCC_FLAGS + list(defs) + ["-I", os.path.join(ROOT, "include"), "-o", out] + srcs
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return out


def build(force=False, verbose=False):
    srcs = [s for s in SRC if os.path.exists(s)]
    if not force and not needs_build():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + DEFAULT_DEFS + ["-I", os.path.join(ROOT, "include"), "-o", OUT] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(OUT)
