"""Where do two FP32 kernel variants differ?  (development aid)  usage: kernel_diff.py  -> per scene / method: pixels that differ, max relative difference"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import minimal_volumetric_path_tracer_b200 as v
from oracle_lib import DEFAULT_SCENE
if os.environ.get("VPT_LIB"):
    v.api.LIB_PATH = os.path.abspath(os.environ["VPT_LIB"])

def big_scene(n_big):
    rng = np.random.default_rng(40 + n_big)
    rows = [r.copy() for r in DEFAULT_SCENE]
    z = [0.0] * 7
    for k in range(n_big):
        rows.append(np.array([70.0 + 10 * k, -80 + 160 * k, -95 - 10 * k, 20, .4, .5, .6, 0, 0, 0, 0, *z]))
    while len(rows) < 32:
        c = [rng.uniform(-40, 40), rng.uniform(-35, 20), rng.uniform(-70, 120)]
        if rng.random() < 0.25:
            rows.append(np.array([rng.uniform(2, 5), *c, 0, 0, 0, 0, 0, 0, 1, 0.2, 0.92, 1.1, 3.9, 2.45, 2.14, rng.uniform(0.1, 0.3)]))
        else:
            rows.append(np.array([rng.uniform(2, 6), *c, *rng.uniform(0.2, 0.9, 3), 0, 0, 0, 0, *z]))
    return np.array(rows)

gold = np.load(os.path.join(ROOT, "tests", "golden", "dielectric.npz"))
scenes = {"default": DEFAULT_SCENE, "big1": big_scene(1), "big2": big_scene(2)}
for k in gold.files:
    if k.startswith("rows_"):
        scenes[k[5:]] = gold[k]
for name, rows in scenes.items():
    sc = v.scene_from_rows(rows)
    for method in (0, 1):
        p = v.default_params(width=192, height=144, spp=1, method=method, seed=17, output=v.OUTPUT_SUM)
        a, sa = v.render(p, sc, stats=True)
        for other, kern in (("hbm", v.KERNEL_WAVEFRONT_HBM), ("mega", v.KERNEL_MEGA)):
            b, sb = v.render(p.copy(kernel=kern), sc, stats=True)
            diff = (a != b).any(axis=2)
            fin = np.isfinite(a).all(axis=2) & np.isfinite(b).all(axis=2)
            rel = np.abs(a - b).max(axis=2) / np.maximum(np.abs(a).max(axis=2), 1e-6)
            print("%-8s method %d sm vs %-4s: events %d/%d scans %d/%d nonfinite %d/%d pixels differing %d of %d, max rel %.3e, median rel of differing %.3e" % (
                name, method, other, sa.events, sb.events, sa.scene_scans, sb.scene_scans, sa.nonfinite, sb.nonfinite, int(diff.sum()), diff.size,
                float(rel[fin].max()), float(np.median(rel[diff & fin])) if (diff & fin).any() else 0.0), flush=True)
