"""Builds libsoftmac_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libsoftmac_b200.so")
SOURCES = [os.path.join(CSRC, "smx_api.cu")]
HEADERS = [os.path.join(CSRC, h) for h in ("smx_math.cuh", "smx_contact.cuh", "smx_kernels.cuh", "smx_sdf.cuh", "smx_rigid.cuh")] + \
          [os.path.join(os.path.dirname(HERE), "include", "softmac_b200.h")]
# -prec-div=false -prec-sqrt=false -ftz=true: approximate reciprocal / square root (2 ulp) and flushed denormals; measured +0.6 % (rest) /
# +1.6 % (stressed) on cube-1M with the whole GPU parity suite unchanged and green (gpurun_out r2q).  NOT --use_fast_math: fmad is
# already on and the __sinf / __expf substitutions are not wanted.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-prec-div=false", "-prec-sqrt=false", "-ftz=true",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-g", "-Xcompiler", "-fopenmp", "-shared", "--extended-lambda", "-Xptxas", "-v"]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force=False, verbose=False, extra=(), out=None):
    """extra / out: ablation builds (`python -m softmac_b200.build --out lib/variant.so -DSMX_...`), loaded with SMX_LIB=..."""
    if out is None and not force and not stale():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    flags = list(NVCC_FLAGS) + list(extra)
    target = out or LIB
    cmd = [nvcc()] + flags + ["-o", target] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(HERE, "lib", "build.log") if out is None else target + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libsoftmac_b200.so")
    if verbose:
        print(log)
    return target


if __name__ == "__main__":
    a = sys.argv[1:]
    out = a[a.index("--out") + 1] if "--out" in a else None
    print(build(force="--force" in a, verbose=out is None, extra=[x for x in a if x.startswith("-") and x not in ("--out", "--force")], out=out))
