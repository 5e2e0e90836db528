import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import minimal_volumetric_path_tracer_b200 as v
w, h, spp, method = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
p = v.default_params(width=w, height=h, spp=spp, method=method, kernel=v.KERNEL_WAVEFRONT_SM if len(sys.argv) > 5 and sys.argv[5] == "sm" else v.KERNEL_WAVEFRONT)
hdr, st = v.render(p, stats=True)
print("ok", w, h, spp, method, hdr.mean(axis=(0, 1)), st.paths, st.events, st.scene_scans, st.nonfinite, "%.1f Mpaths/s" % (st.paths / st.kernel_ms / 1e3))
