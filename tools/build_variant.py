"""Build an experimental copy of the library with extra -D flags on the FP32 translation unit (development only):
    python tools/build_variant.py <name> -DVPT_SM_THREADS=896 ...      -> tools/_variants/<name>.so
    (a name starting with "d_" applies the flags to the FP64 translation unit instead, e.g. d_t384 -DVPT_SMD_THREADS=384)
    VPT_LIB=tools/_variants/<name>.so python tools/gpu_time.py 1024 smwave        (on the GPU box)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from minimal_volumetric_path_tracer_b200 import build as b
name, defs = sys.argv[1], sys.argv[2:]
b.build_library()
out = os.path.join(ROOT, "tools", "_variants"); os.makedirs(out, exist_ok=True)
objs = []
for src, flags in b.UNITS:
    if src == ("vpt_kernels_f64.cu" if name.startswith("d_") else "vpt_kernels_f32.cu"):
        o = os.path.join(out, "%s.%s.o" % (os.path.splitext(src)[0], name))
        subprocess.run([b._nvcc()] + b.ARCH + b.COMMON + flags + defs + ["-Xptxas", "-v", "-x", "cu", "-c", os.path.join(b.CSRC, src), "-o", o], check=True,
                       stderr=open(os.path.join(out, name + ".ptxas.txt"), "w"))
    else:
        o = os.path.join(b.OBJ, os.path.splitext(src)[0] + ".o")
    objs.append(o)
subprocess.run([b._nvcc()] + b.ARCH + ["-shared", "-o", os.path.join(out, name + ".so")] + objs, check=True)
print("built", os.path.join(out, name + ".so"))
