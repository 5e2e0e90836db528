import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import minimal_volumetric_path_tracer_b200 as v
from oracle_lib import L1, DEFAULT_SCENE, scene_without
l1 = L1()
def bm(img, block=16):
    h, w, _ = img.shape
    return img.reshape(h // block, block, w // block, block, 3).mean(axis=(1, 3))
# 1. full-size CRN, low spp, also at a high sample offset
for prec in (v.PRECISION_FP64_REF, v.PRECISION_FP32):
    for (b, e) in ((0, 4), (4090, 4094)):
        p = v.default_params(width=1024, height=768, spp=8192, sample_begin=b, sample_end=e, method=0, precision=prec, seed=123, output=v.OUTPUT_SUM)
        hdr, st = v.render(p, stats=True)
        ref, _, rst = l1.render(DEFAULT_SCENE, 0, 0, 0.001, 0.009, 1024, 768, 123, e - b, sample_begin=b, want_sumsq=False)
        err = np.abs(hdr - ref) / np.maximum(np.abs(ref), 1e-3)
        print("CRN full size prec", prec, (b, e), "events", st.events, rst["events"], "median err", np.median(err), "frac>1e-3", np.mean(err > 1e-3),
              "mean dev", hdr.mean(axis=(0, 1)), "mean ref", ref.mean(axis=(0, 1)), flush=True)
# 2. sign tests at equal spp against the goldens
for name, prec, quirks, scene in (("no8_m0", v.PRECISION_FP32, 0, scene_without([8])), ("no8_m0", v.PRECISION_FP64_REF, 0, scene_without([8])),
                                  ("robust_m0", v.PRECISION_FP32, 0, None), ("strict_m0", v.PRECISION_FP64_REF, 3, None), ("strict_m1", v.PRECISION_FP64_REF, 3, None),
                                  ("no8_m1", v.PRECISION_FP32, 0, scene_without([8])), ("robust_m2", v.PRECISION_FP32, 0, None)):
    g = np.load(os.path.join(ROOT, "tests", "golden", "image_%s.npz" % name))
    spp = int(g["spp"]); B = g["block_mean"].astype(np.float64)
    fr = []
    for seed in (11, 12, 13):
        p = v.default_params(spp=spp, method=int(g["method"]), precision=prec, quirks=quirks, seed=seed)
        A = bm(v.render(p, v.scene_from_rows(scene) if scene is not None else None).astype(np.float64))
        fr.append([float(np.mean((A - B)[..., c] > 0)) for c in range(3)] + [float(A.mean() / B.mean())])
    print("sign", name, "prec", prec, np.round(np.array(fr), 4).tolist(), flush=True)
# 3. GPU vs GPU sign test (different seeds), as the null
p = v.default_params(spp=256, method=0, seed=21); q = p.copy(seed=22)
A = bm(v.render(p).astype(np.float64)); B = bm(v.render(q).astype(np.float64))
print("sign gpu-gpu", [float(np.mean((A - B)[..., c] > 0)) for c in range(3)])
