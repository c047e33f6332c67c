"""Compile the sm_100a CUDA library in-tree: myrtle-vision_b200/csrc/libmv_b200.so.

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with
the gpurun snapshot.  Usage: python myrtle-vision_b200/csrc/build.py [--force] [--verbose]
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libmv_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def sources():
    return sorted(glob.glob(os.path.join(HERE, "*.cu")))


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + glob.glob(os.path.join(HERE, "*.cuh")) + glob.glob(
        os.path.join(HERE, "..", "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, report=None):
    """-> path of the .so.  `report` (a list) receives one line saying what was done."""
    if not force and not needs_build():
        if report is not None:
            report.append("csrc: %s is newer than every source (%d .cu files) — nothing to compile "
                          "(MV_FORCE_BUILD=1 or build.py --force recompiles)" % (os.path.basename(SO), len(sources())))
        return SO
    objs = []
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
             "-Xcompiler", "-fPIC", "--use_fast_math=false"]
    flags = [f for f in flags if f != "--use_fast_math=false"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    flags += os.environ.get("MV_NVCC_FLAGS", "").split()      # e.g. -DMV_SN_TRACE for tools/trace_attn_bwd.py
    procs = []
    for src in sources():
        obj = src[:-3] + ".o"
        objs.append(obj)
        procs.append((src, subprocess.Popen([NVCC] + flags + ["-c", src, "-o", obj])))
    for src, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed on " + src)
    subprocess.check_call([NVCC, "-shared", "-o", SO] + objs +
                          ["-gencode", "arch=compute_100a,code=sm_100a"])
    if report is not None:
        report.append("csrc: compiled %d .cu files with nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 "
                      "and linked %s" % (len(objs), os.path.basename(SO)))
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
