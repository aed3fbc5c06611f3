"""Builds libfd_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

-fmad=false: the reference (Rust, no fast-math) never contracts a*b+c; bit-exact NMS keep lists and fixed-point
resize/warp tables depend on separately rounded operations.  -lineinfo keeps ncu's source page usable.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfd_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["fd_ctx.cu", "fd_ops.cu", "fd_nms.cu", "fd_decode.cu", "fd_detect_fused.cu", "fd_select.cu", "fd_preprocess.cu", "fd_align.cu", "fd_pipeline.cu", "fd_jpeg.cu"]
EXTRA = os.environ.get("FD_NVCC_EXTRA", "").split()   # experiments only, e.g. FD_NVCC_EXTRA=-DFUSED_MAXNREG=48
FLAGS = EXTRA + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
         "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2", "--shared", "-cudart", "shared"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "fd_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [NVCC] + [f for f in FLAGS if f != "--shared"] + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("== %s ==\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([NVCC, "--shared", "-cudart", "shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
