"""Compare an FP32 kernel variant with the megakernel on the same Philox streams (same decisions, sums equal to fp32 rounding).
usage: variant_check.py <kernel name> [w h spp]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import minimal_volumetric_path_tracer_b200 as v
if os.environ.get("VPT_LIB"):  # development only: time an experimental build of the library
    v.api.LIB_PATH = os.path.abspath(os.environ["VPT_LIB"])
kern = {"mega": v.KERNEL_MEGA, "smwave": v.KERNEL_WAVEFRONT_SM, "hbm": v.KERNEL_WAVEFRONT_HBM}[sys.argv[1]]
w, h, spp = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (160, 120, 16)
ok = True
for method in (0, 1, 2):
    for extra in ({}, dict(tile_rank=1, tile_count=3), dict(sigma_a=0.0005, sigma_s=0.0495, continue_prob=0.95, max_depth=64)):
        p = v.default_params(width=w, height=h, spp=spp, method=method, seed=12, output=v.OUTPUT_SUM, **extra)
        a, sa = v.render(p.copy(kernel=v.KERNEL_MEGA), stats=True)
        b, sb = v.render(p.copy(kernel=kern), stats=True)
        b2 = v.render(p.copy(kernel=kern))
        err = np.abs(a - b) / np.maximum(np.abs(a), 1e-3)
        good = (sb.paths == sa.paths and abs(int(sa.events) - int(sb.events)) <= 3e-4 * sa.events and abs(int(sa.scene_scans) - int(sb.scene_scans)) <= 2e-3 * sa.scene_scans
                and np.median(err) < 1e-6 and np.mean(err > 1e-3) < 0.015 and np.array_equal(b, b2) and sb.nonfinite == sa.nonfinite)
        ok &= bool(good)
        print("method %d %s: paths %d/%d events %d/%d scans %d/%d median err %.2e frac>1e-3 %.4f max %.3e reproducible %s -> %s" % (
            method, extra, sb.paths, sa.paths, sb.events, sa.events, sb.scene_scans, sa.scene_scans, np.median(err), np.mean(err > 1e-3), err.max(), np.array_equal(b, b2), "ok" if good else "FAIL"), flush=True)
print("ALL OK" if ok else "FAILED")
