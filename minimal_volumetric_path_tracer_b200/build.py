"""Build libvpt_b200.so (CUDA kernels for sm_100a + the C-ABI of include/vpt.h) and the C++ host `rt`, in-tree.

nvcc cross-compiles without a GPU.  The fp64 REF-mode translation unit is compiled with -fmad=false (strict IEEE, the
reference's rounding); the fp32 unit keeps precise libdevice transcendentals (no -use_fast_math) but uses the fast
division / square root (<= 2 ulp), see DESIGN.md "Math precision".
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libvpt_b200.so")
RT = os.path.join(HERE, "bin", "rt")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
# -Xcicc -O2 on the FP32 unit: at its default -O3 cicc 12.9 dies with SIGSEGV on this translation unit (about one run in three since the
# fused next-event scan, every run with some equivalent spellings of it); -O2 compiles every time and the kernels time the same
# (8647 against 8652 Mpaths/s on C2).  _run() additionally repeats a command that ends with SIGSEGV.
UNITS = [
    ("vpt_kernels_f32.cu", ["-prec-div=false", "-prec-sqrt=false", "-ftz=true", "-Xcicc", "-O2"]),
    ("vpt_kernels_hbm.cu", ["-prec-div=false", "-prec-sqrt=false", "-ftz=true"]),
    ("vpt_kernels_f64.cu", ["-fmad=false"]),
    ("vpt_api.cpp", []),
]
DEPS = sorted(f for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))) + [os.path.join("..", "..", "include", "vpt.h")]  # every header: a stale .so is a silent bug


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    for attempt in range(4):  # cicc 12.9 has been seen to die with SIGSEGV now and then on these sources: the same command succeeds when repeated
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode not in (139, -11):
            break
        print("build: %s crashed (SIGSEGV), attempt %d of 4" % (os.path.basename(cmd[0]), attempt + 1), file=sys.stderr, flush=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build_library(force=False, verbose=False, ptxas_info=False):
    os.makedirs(OBJ, exist_ok=True)
    deps = [os.path.join(CSRC, d) for d in DEPS] + [os.path.abspath(__file__)]
    objs, log = [], ""
    for src, flags in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + deps):
            cmd = [_nvcc()] + ARCH + COMMON + flags + (["-Xptxas", "-v"] if ptxas_info else []) + ["-x", "cu", "-c", s, "-o", o]
            log += _run(cmd, verbose)
    if force or _stale(LIB, objs):
        log += _run([_nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-fvisibility=hidden"], verbose)
    return log


def build_host(force=False, verbose=False):
    """The C++ host that keeps the reference's CLI (./rt <spp> -> image.ppm) and calls vpt_render() through the C-ABI."""
    src = os.path.join(HOST, "rt_main.cpp")
    os.makedirs(os.path.dirname(RT), exist_ok=True)
    if force or _stale(RT, [src, LIB, os.path.join(HERE, "..", "include", "vpt.h")]):
        cxx = "g++"
        _run([cxx, "-O2", "-std=c++17", "-I", os.path.join(HERE, "..", "include"), src, "-o", RT, "-L", HERE, "-lvpt_b200",
              "-Wl,-rpath,$ORIGIN/.."], verbose)


def build_all(force=False, verbose=False):
    log = build_library(force, verbose)
    if os.path.exists(os.path.join(HOST, "rt_main.cpp")):
        build_host(force, verbose)
    return log


if __name__ == "__main__":
    out = build_library(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv)
    if "--ptxas" in sys.argv:
        print(out)
    if os.path.exists(os.path.join(HOST, "rt_main.cpp")):
        build_host(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
