"""development aid: FP64 device vs oracle, per path, VPT_METHOD_VOLUME_SPHERES"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import minimal_volumetric_path_tracer_b200 as v
import oracle_lib as ol
l1 = ol.L1()
g = np.load(os.path.join(ROOT, "tests", "golden", "volume_spheres.npz"))
rows = g["rows"]; sc = v.scene_from_rows(rows)
n = 4000
rng = np.random.default_rng(1)
o = np.tile(np.array(ol.CAM_O), (n, 1)); d = np.array([l1.camera_ray(256, 192, int(rng.integers(256)), int(rng.integers(192)), rng.random(), rng.random()) for _ in range(n)])
pix = rng.integers(0, 2 ** 20, n).astype(np.uint32); smp = rng.integers(0, 2 ** 14, n).astype(np.uint32)
inp = np.concatenate([o, d, pix[:, None].astype(float), smp[:, None].astype(float)], axis=1)
for quirks in (0, 3):
    want, ev = l1.radiance_philox(rows, quirks, 5, 0.001, 0.009, 42, o, d, pix, smp)
    got = v.unit(v.UNIT.RADIANCE, inp, v.default_params(method=5, precision=v.PRECISION_FP64_REF, quirks=quirks, seed=42), sc)
    e = np.abs(got[:, :3] - want).max(axis=1) / np.maximum(np.abs(want).max(axis=1), 1e-30)
    bad = np.argsort(-e)[:12]
    print("quirks", quirks, "events equal", np.mean(got[:, 3] == ev), "frac > 1e-9", np.mean(e > 1e-9), "max", e.max())
    for i in bad:
        hit, t, idx = l1.intersect(rows, quirks, o[i], d[i])
        print("  ray %d first hit %d events %d/%d err %.3e got %s want %s" % (i, idx, got[i, 3], ev[i], e[i], got[i, :3], want[i]))
