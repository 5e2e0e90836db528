"""Device-timed Mpaths/s of the kernel variants on the C2 frame (1024x768, default scene), all four shade methods.
usage: gpu_time.py [spp] [mega,smwave,hbm,f64,f64mega]      VPT_LIB=tools/_variants/<name>.so selects an experimental build;
VPT_W / VPT_H = another frame size, VPT_METHODS=1,2 = a subset of the shade methods"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import minimal_volumetric_path_tracer_b200 as v
if os.environ.get("VPT_LIB"):  # development only: time an experimental build of the library
    v.api.LIB_PATH = os.path.abspath(os.environ["VPT_LIB"])
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kernels = {"mega": dict(kernel=v.KERNEL_MEGA), "smwave": dict(kernel=v.KERNEL_WAVEFRONT_SM), "hbm": dict(kernel=v.KERNEL_WAVEFRONT_HBM),
           "f64": dict(kernel=v.KERNEL_AUTO, precision=v.PRECISION_FP64_REF, quirks=v.QUIRKS_REFERENCE),
           "f64mega": dict(kernel=v.KERNEL_MEGA, precision=v.PRECISION_FP64_REF, quirks=v.QUIRKS_REFERENCE)}
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["mega", "smwave", "hbm"]
for name in names:
    for method in [int(m) for m in os.environ.get("VPT_METHODS", "0,1,2,4").split(",")]:
        p = v.default_params(spp=spp, method=method, **kernels[name])
        if os.environ.get("VPT_W"): p.width, p.height = int(os.environ["VPT_W"]), int(os.environ["VPT_H"])
        v.render(p)
        best = 0
        for _ in range(3):
            hdr, st = v.render(p, stats=True)
            best = max(best, st.paths / st.kernel_ms / 1e3)
        print("%s method %d spp %d: %.1f Mpaths/s  (events/path %.3f scans/path %.3f) mean %s" % (name, method, spp, best, st.events / st.paths, st.scene_scans / st.paths, hdr.mean(axis=(0, 1))), flush=True)
