"""Python twin of the C++ host (host/rt_main.cpp): the reference's command line `rt <spp>` -> image.ppm (src/rt.cpp:744-830)
with the hard-coded choices of the reference exposed as flags.  The pixel loop runs in libvpt_b200.so on the GPU."""
import argparse
import sys
import time

from . import api

METHODS = {"free": api.METHOD_FREE_FLIGHT, "equi": api.METHOD_EQUIANGULAR, "mis": api.METHOD_MIS, "march": api.METHOD_RAYMARCH,
           "mis-distance": api.METHOD_MIS_DISTANCE}  # not in the reference: one-sample MIS of the free-flight and equi-angular distance techniques


def parse_args(argv):
    ap = argparse.ArgumentParser(prog="rt", description="B200 volumetric path tracer: rt <spp> -> image.ppm")
    ap.add_argument("spp", type=int, help="samples per pixel (the reference's only argument, rt.cpp:784)")
    ap.add_argument("--method", choices=sorted(METHODS), default="free", help="shade method (rt.cpp:791-796 picks by recompiling)")
    ap.add_argument("--size", default="1024x768", help="WxH (rt.cpp:752)")
    ap.add_argument("--sigma-a", type=float, default=0.001)
    ap.add_argument("--sigma-s", type=float, default=0.009)
    ap.add_argument("--march-step", type=float, default=0.1, help="--method march: step length (rt.cpp:791)")
    ap.add_argument("--march-source", type=int, default=7, help="--method march: source sphere (rt.cpp:791 passes 7; 8 is the point light)")
    ap.add_argument("--scene", default=None, help="scene file (scenes/*.txt); default: the reference's active scene, Sphere.cpp:11-22")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-depth", type=int, default=0)
    ap.add_argument("--continue-prob", type=float, default=0.6)
    ap.add_argument("--ref", action="store_true", help="FP64 REF mode with the reference's rounding-decided behaviours")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("-o", "--output", default="image.ppm")
    ap.add_argument("--pfm", default=None, help="also write the frame before the tonemap as a binary PFM (little-endian floats)")
    a = ap.parse_args(argv)
    try:
        w, h = (int(v) for v in a.size.lower().split("x"))
    except ValueError:
        ap.error("--size must look like 1024x768")
    if a.spp <= 0 or w <= 0 or h <= 0:
        ap.error("spp and size must be positive")
    a.width, a.height = w, h
    return a


def params_from_args(a):
    p = api.default_params(width=a.width, height=a.height, spp=a.spp, method=METHODS[a.method], sigma_a=a.sigma_a, sigma_s=a.sigma_s,
                           seed=a.seed, max_depth=a.max_depth, continue_prob=a.continue_prob, march_step=a.march_step, march_source=a.march_source)
    if a.ref:
        p.precision = api.PRECISION_FP64_REF
        p.quirks = api.QUIRKS_REFERENCE
    return p


def main(argv=None):
    a = parse_args(sys.argv[1:] if argv is None else argv)
    start = time.time()
    p = params_from_args(a)
    scene = api.load_scene(a.scene) if a.scene else api.default_scene()
    if a.gpus > 1:
        hdr, st = api.render_multi(p, scene, list(range(a.gpus)), stats=True)
    else:
        hdr, st = api.render(p, scene, stats=True)
    sys.stderr.write("\r%5.2f%%\n" % 100.0)
    api.write_ppm(hdr, a.output)
    if a.pfm:
        api.write_pfm(hdr, a.pfm)
    sys.stderr.write("paths %d  events %d  scans %d  kernel %.3f ms  (%.1f Mpaths/s)\n" % (st.paths, st.events, st.scene_scans, st.kernel_ms,
                                                                                      st.paths / (st.kernel_ms * 1e3)))
    print("elapsed time: %gs" % (time.time() - start))
    return 0


if __name__ == "__main__":
    sys.exit(main())
