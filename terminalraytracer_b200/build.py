"""Builds libtrt_b200.so (the product: CUDA kernels + C ABI + host C) in-tree with nvcc for sm_100a.

Run as `python -m terminalraytracer_b200.build` or through __graft_entry__.build().
The render TU is compiled with -fmad=false: the reference's x86-64 build never fuses a*b+c, and the
output must be bit-identical ("Exactness")."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(HERE, "libtrt_b200.so")
OBJ = os.path.join(HERE, "_obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
               "-I", INCLUDE, "-I", CSRC]
# per-TU extra flags
CU_SOURCES = {
    "trt_render.cu": ["-fmad=false", "-Xptxas", "-v"] + os.environ.get("TRT_EXTRA_NVCC", "").split(),   # TRT_EXTRA_NVCC: experiment builds only
    "trt_encode.cu": ["-fmad=false"],
    "trt_peak.cu": [],
    "trt_api.cu": ["-fmad=false"],
}
C_SOURCES = {"trt_host.c": ["-O3", "-ffp-contract=off", "-fPIC", "-I", INCLUDE]}


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libtrt_b200 cannot be built (there is no CPU fallback)")
    return nvcc


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, defines=(), out=None):
    """defines/out: experiment builds (e.g. defines=["TRT_MIN_CTAS_PER_SM=6"], out="libtrt_b200_x.so");
    the product build uses neither."""
    global OBJ, LIB
    if out:
        LIB = os.path.join(HERE, out)
        OBJ = os.path.join(HERE, "_obj_" + os.path.splitext(out)[0])
        force = True
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    headers.append(os.path.abspath(__file__))
    objs = []
    log = []
    for src, extra in CU_SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + ARCH + NVCC_COMMON + extra + ["-D" + d for d in defines] + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log.append(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + log[-1])
    for src, flags in C_SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [os.environ.get("CC", "gcc")] + flags + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log.append(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("cc failed:\n" + log[-1])
    if force or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + log[-1])
    if verbose:
        print("\n".join(log))
    with open(os.path.join(OBJ, "build.log"), "a") as f:
        f.write("\n".join(log))
    return LIB


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build_library(force="--force" in sys.argv, verbose="--quiet" not in sys.argv, defines=defs, out=outs[0] if outs else None))
