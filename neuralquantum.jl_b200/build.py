"""Build libnqcuda.so in-tree for sm_100a with nvcc (no GPU needed: cross-compiles).

    python neuralquantum.jl_b200/build.py [--force] [--verbose]

Objects go to neuralquantum.jl_b200/build/, the library to neuralquantum.jl_b200/libnqcuda.so
(git-ignored, but it travels to the GPU box with the gpurun snapshot).
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libnqcuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-DNQ_BUILDING"]


def _nccl_include():
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            inc = os.path.join(list(spec.submodule_search_locations)[0], "include")
            if os.path.exists(os.path.join(inc, "nccl.h")):
                return inc
    except Exception:
        pass
    return "/usr/include"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = (glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.inc")) +
            glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    inc = ["-I", _nccl_include()]
    # the linked library is newer than every source: nothing to do (object files do not travel to the GPU box)
    if not force and os.path.exists(LIB) and not _stale(LIB, srcs + hdrs):
        build_cabi_smoke()
        return LIB
    jobs = []
    for s in srcs:
        o = os.path.join(BUILD, os.path.basename(s)[:-3] + ".o")
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC] + ARCH + FLAGS + inc + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append((s, cmd))
    def run(job):
        s, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed on %s" % s)
    objs = [os.path.join(BUILD, os.path.basename(s)[:-3] + ".o") for s in srcs]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    build_cabi_smoke()
    return LIB


def build_cabi_smoke():
    """tests/cabi_smoke: a plain-C host of the ABI (gcc, host buffers only), run by tests/test_gpu_cabi_c.py."""
    src = os.path.join(HERE, "..", "tests", "cabi_smoke.c")
    exe = os.path.join(HERE, "..", "tests", "cabi_smoke")
    if not os.path.exists(src) or (os.path.exists(exe) and not _stale(exe, [src, LIB])):
        return exe
    cmd = ["gcc", "-O2", "-std=c11", src, "-I", os.path.join(HERE, "..", "include"), "-L", HERE, "-lnqcuda", "-lm",
           "-Wl,-rpath,$ORIGIN/../neuralquantum.jl_b200", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("gcc failed on tests/cabi_smoke.c")
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
