"""Small fixed workload for ncu: the C2 frame (1024x768, default scene) at reduced spp -- per-sample behaviour is identical to
the benchmark's, the launch is just shorter.  usage: profile_target.py [--method 1] [--spp 64] [--precision fp32|fp64ref] [--reps 2] [--kernel auto|mega|smwave|hbm] [--config default|c4]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import minimal_volumetric_path_tracer_b200 as v
if os.environ.get("VPT_LIB"):  # development only: time an experimental build of the library
    v.api.LIB_PATH = os.path.abspath(os.environ["VPT_LIB"])
ap = argparse.ArgumentParser()
ap.add_argument("--method", type=int, default=1); ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--precision", default="fp32"); ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--width", type=int, default=1024); ap.add_argument("--height", type=int, default=768)
ap.add_argument("--config", default="default"); ap.add_argument("--kernel", default="auto")
a = ap.parse_args()
kw = dict(sigma_a=0.0005, sigma_s=0.0495, continue_prob=0.95, max_depth=64) if a.config == "c4" else {}
p = v.default_params(width=a.width, height=a.height, spp=a.spp, method=a.method, seed=1, kernel={"auto": v.KERNEL_AUTO, "mega": v.KERNEL_MEGA, "smwave": v.KERNEL_WAVEFRONT_SM, "hbm": v.KERNEL_WAVEFRONT_HBM}[a.kernel], **kw)
if a.precision != "fp32":
    p.precision = v.PRECISION_FP64_REF; p.quirks = v.QUIRKS_REFERENCE
for _ in range(a.reps):
    hdr, st = v.render(p, stats=True)
    print("method %d %s spp %d: kernel %.3f ms, %.1f Mpaths/s, events/path %.3f, scans/path %.3f" %
          (a.method, a.precision, a.spp, st.kernel_ms, st.paths / st.kernel_ms / 1e3, st.events / st.paths, st.scene_scans / st.paths), flush=True)
