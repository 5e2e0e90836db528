import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import minimal_volumetric_path_tracer_b200 as v
from oracle_lib import scene_without
def bm(img, block=16):
    h, w, _ = img.shape
    return img.reshape(h // block, block, w // block, block, 3).mean(axis=(1, 3))
def summary(z):
    return np.array([np.mean(np.clip(z, -6, 6)), np.median(z), np.median(np.abs(z)), np.mean(np.abs(z) > 3)])
for name, prec, quirks, scene, spp in (("no8_m1", 0, 0, scene_without([8]), 4096), ("strict_m0", 1, 3, None, 512), ("robust_m2", 0, 0, None, 4096), ("no8_m0", 0, 0, scene_without([8]), 4096)):
    g = np.load(os.path.join(ROOT, "tests", "golden", "image_%s.npz" % name))
    spp_ref = int(g["spp"]); ref = g["block_mean"].astype(np.float64); method = int(g["method"])
    sc = v.scene_from_rows(scene) if scene is not None else None
    p = v.default_params(spp=spp, method=method, precision=prec, quirks=quirks, seed=77)
    main = bm(v.render(p, sc).astype(np.float64))
    batches = 16; per = max(spp // (4 * batches), 4)
    q = p.copy(spp=per * batches, seed=79, output=v.OUTPUT_SUM)
    parts = np.stack([bm(v.render(q.copy(sample_begin=b * per, sample_end=(b + 1) * per), sc).astype(np.float64)) / per for b in range(batches)])
    sigma = np.sqrt(np.maximum(parts.var(axis=0, ddof=1) * per * (1.0 / spp_ref + 1.0 / spp), 1e-300))
    print(name, "ref    ", np.round(summary(((main - ref) / sigma).ravel()), 4), flush=True)
    for seed in range(100, 108):
        s = bm(v.render(p.copy(spp=spp_ref, seed=seed), sc).astype(np.float64))
        print(name, "standin", np.round(summary(((main - s) / sigma).ravel()), 4), "sign vs ref", np.round(np.mean((s - ref) > 0, axis=(0, 1)), 4), flush=True)
