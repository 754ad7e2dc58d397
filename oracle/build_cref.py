"""Build oracle/_ref/libcref.so (the C restatement used as CPU baseline and second oracle).
gcc only; outputs stay under oracle/_ref/ (git-ignored, but shipped to the GPU box)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libcref.so")


def build(force=False):
    src = os.path.join(HERE, "cref.c")
    os.makedirs(OUT, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    cmd = ["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-o", LIB, src, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
