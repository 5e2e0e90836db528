"""Scratch GPU probe (development aid, not a test): device-vs-oracle per-path parity, small-render parity, rough timings."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import minimal_volumetric_path_tracer_b200 as v
from oracle_lib import L1, DEFAULT_SCENE, CAM_O

out = {}
l1 = L1()
print(v.version(), "devices", v.device_count(), flush=True)
# philox KAT vs oracle
rng = np.random.default_rng(1)
ctr = rng.integers(0, 2**32, size=(1000, 4), dtype=np.uint64).astype(np.uint32); key = rng.integers(0, 2**32, size=(1000, 2), dtype=np.uint64).astype(np.uint32)
dev = v.philox(ctr, key)
ref = np.array([l1.philox(c, k) for c, k in zip(ctr, key)])
out["philox_equal"] = bool(np.array_equal(dev, ref)); print("philox equal:", out["philox_equal"], flush=True)

# per-path radiance
n = 20000
o = np.zeros((n, 3)); d = np.zeros((n, 3))
for i in range(n):
    if i % 2 == 0:
        o[i] = CAM_O; d[i] = l1.camera_ray(1024, 768, int(rng.integers(1024)), int(rng.integers(768)), rng.random(), rng.random())
    else:
        o[i] = [rng.uniform(-45, 45), rng.uniform(-38, 38), rng.uniform(-75, 150)]; x = rng.normal(size=3); d[i] = x / np.linalg.norm(x)
pix = rng.integers(0, 2**20, n).astype(np.uint32); smp = rng.integers(0, 2**12, n).astype(np.uint32)
rows = np.concatenate([o, d, pix[:, None].astype(float), smp[:, None].astype(float)], axis=1)
scene = v.scene_from_rows(DEFAULT_SCENE)
for prec, quirks in ((v.PRECISION_FP64_REF, 3), (v.PRECISION_FP64_REF, 0), (v.PRECISION_FP32, 0)):
    for method in (0, 1, 2):
        p = v.default_params(method=method, precision=prec, quirks=quirks, seed=7)
        t0 = time.time(); got = v.unit(v.UNIT.RADIANCE, rows, p, scene); t1 = time.time()
        want, ev = l1.radiance_philox(DEFAULT_SCENE, quirks, method, 0.001, 0.009, 7, o, d, pix, smp)
        L = got[:, :3]
        den = np.maximum(np.abs(want).max(axis=1), 1e-12)
        rel = np.abs(L - want).max(axis=1) / den
        both_zero = (np.abs(want).max(axis=1) == 0) & (np.abs(L).max(axis=1) == 0)
        rel[both_zero] = 0
        ev_eq = (got[:, 3] == ev)
        tag = "p%d_q%d_m%d" % (prec, quirks, method)
        out[tag] = dict(median_rel=float(np.median(rel[~both_zero])), p99_rel=float(np.quantile(rel, 0.99)), frac_gt_1e3=float((rel > 1e-3).mean()),
                        frac_gt_1e5=float((rel > 1e-5).mean()), events_equal=float(ev_eq.mean()), nan=int(np.isnan(L).any(axis=1).sum()),
                        mean_dev=L.mean(axis=0).tolist(), mean_ref=want.mean(axis=0).tolist())
        print(tag, out[tag], flush=True)

# small render parity: device vs L1, CRN
w, h, spp = 128, 96, 16
for prec, quirks in ((v.PRECISION_FP64_REF, 3), (v.PRECISION_FP32, 0)):
    for method in (0, 1, 2):
        p = v.default_params(width=w, height=h, spp=spp, method=method, precision=prec, quirks=quirks, seed=3, output=v.OUTPUT_SUM)
        hdr, st = v.render(p, scene, stats=True)
        ref, _, rst = l1.render(DEFAULT_SCENE, quirks, method, 0.001, 0.009, w, h, 3, spp, want_sumsq=False)
        den = np.maximum(np.abs(ref), 1e-3)
        rel = np.abs(hdr - ref) / den
        tag = "render_p%d_m%d" % (prec, method)
        out[tag] = dict(mean_dev=hdr.mean(axis=(0, 1)).tolist(), mean_ref=ref.mean(axis=(0, 1)).tolist(), median_rel=float(np.median(rel)),
                        frac_gt_1e3=float((rel > 1e-3).mean()), events_dev=int(st.events), events_ref=rst["events"], scans_dev=int(st.scene_scans), scans_ref=rst["scans"],
                        nonfinite=int(st.nonfinite), kernel_ms=st.kernel_ms)
        print(tag, out[tag], flush=True)

# timings
peak, clk = v.measure_fp32_peak(0); out["fp32_peak_tflops"] = peak; out["sm_clock_mhz_max"] = clk; print("fp32 peak", peak, clk, flush=True)
for prec, quirks, spp in ((v.PRECISION_FP32, 0, 256), (v.PRECISION_FP64_REF, 3, 32)):
    for method in (0, 1, 2):
        p = v.default_params(width=1024, height=768, spp=spp, method=method, precision=prec, quirks=quirks, seed=1)
        v.render(p, scene)
        hdr, st = v.render(p, scene, stats=True)
        tag = "time_p%d_m%d" % (prec, method)
        out[tag] = dict(spp=spp, kernel_ms=st.kernel_ms, total_ms=st.total_ms, mpaths_s=st.paths / st.kernel_ms / 1e3, events_per_path=st.events / st.paths,
                        scans_per_path=st.scene_scans / st.paths, mean=hdr.mean(axis=(0, 1)).tolist(), nonfinite=int(st.nonfinite))
        print(tag, out[tag], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
