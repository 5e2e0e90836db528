"""In-kernel cycle breakdown of the SM-wide wavefront kernel.  Needs a library built with -DVPT_SMWAVE_PROFILE:
    python tools/smwave_timing.py --build            (here: compiles tools/_variants/prof.so)
    VPT_LIB=tools/_variants/prof.so python tools/smwave_timing.py [method] [spp]     (on the GPU box)"""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "--build" in sys.argv:
    from minimal_volumetric_path_tracer_b200 import build as b
    out = os.path.join(ROOT, "tools", "_variants"); os.makedirs(out, exist_ok=True)
    objs = []
    for src, flags in b.UNITS:
        o = os.path.join(out, os.path.splitext(src)[0] + ".prof.o"); objs.append(o)
        subprocess.run([b._nvcc()] + b.ARCH + b.COMMON + flags + ["-DVPT_SMWAVE_PROFILE", "-x", "cu", "-c", os.path.join(b.CSRC, src), "-o", o], check=True)
    subprocess.run([b._nvcc()] + b.ARCH + ["-shared", "-o", os.path.join(out, "prof.so")] + objs, check=True)
    print("built", os.path.join(out, "prof.so")); sys.exit(0)
import minimal_volumetric_path_tracer_b200 as v
if os.environ.get("VPT_LIB"):
    v.api.LIB_PATH = os.path.abspath(os.environ["VPT_LIB"])
method = int(sys.argv[1]) if len(sys.argv) > 1 else 1
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
p = v.default_params(spp=spp, method=method, kernel=v.KERNEL_WAVEFRONT_SM)
v.render(p)
hdr, st = v.render(p, stats=True)
lib = v.load_library()
buf = (C.c_ulonglong * 32)()
lib.vpt_debug_counters(buf, 32)
d = list(buf)
tot = d[19] or 1
names = ["PRIMARY", "MED_POINT", "MED_AREA", "SURF_P", "SURF_L", "SURF_F", "GEN", "GEN(tail)"]
print("method %d spp %d: %.1f Mpaths/s, kernel %.2f ms, rounds/CTA %.0f" % (method, spp, st.paths / st.kernel_ms / 1e3, st.kernel_ms, d[20] / 148))
for i, n in enumerate(names):
    if d[8 + i]:
        print("  %-10s %5.1f%% of warp time, %8d batches, %7.0f cycles/batch" % (n, 100 * d[i] / tot, d[8 + i], d[i] / d[8 + i]))
print("  arrival spread at (A) (idle warp time) %5.1f%%   last arrival -> (B) passed %5.1f%% = %.0f cycles per round; round = %.0f cycles" % (
    100 * d[23] / tot, 100 * d[22] / tot, d[22] / 24 / max(d[20], 1), d[19] / (148 * 24) / max(d[20] / 148, 1)))
print("  barrier(A) wait %5.1f%%   plan (A->B) %5.1f%%   claim loop overhead %5.1f%%   flush %5.1f%%   other %5.1f%%" % (
    100 * d[16] / tot, 100 * d[17] / tot, 100 * d[18] / tot, 100 * d[21] / tot, 100 * (tot - sum(d[0:8]) - d[16] - d[17] - d[18] - d[21]) / tot))
