"""time vpt_render() (host buffers) call by call: usage e2e_probe.py [spp] [method]"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import minimal_volumetric_path_tracer_b200 as v
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
method = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lib = v.load_library()
p = v.default_params(spp=spp, method=method)
scene = v.default_scene()
host = np.empty((p.height, p.width, 3), dtype=np.float32)
for i in range(8):
    t0 = time.perf_counter()
    rc = lib.vpt_render(C.byref(p), scene, len(scene), host.ctypes.data_as(C.POINTER(C.c_float)), None)
    t1 = time.perf_counter()
    print("call %d: %.2f ms rc %d" % (i, (t1 - t0) * 1e3, rc), flush=True)
st = v.Stats()
lib.vpt_render(C.byref(p), scene, len(scene), host.ctypes.data_as(C.POINTER(C.c_float)), C.byref(st))
print("kernel %.2f ms total %.2f ms" % (st.kernel_ms, st.total_ms))
