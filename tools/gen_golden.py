"""Generate tests/golden/* from the UNMODIFIED reference (oracle/_ref/libvpt_l0.so, built from /root/reference by
oracle/Makefile).  Run in the build container only (the reference tree does not travel); the outputs are committed.

  python tools/gen_golden.py units     -> tests/golden/units.npz      unit-function vectors (inputs, explicit uniforms, outputs)
  python tools/gen_golden.py paths     -> tests/golden/paths.npz      per-path radiance on seeded erand48 streams
  python tools/gen_golden.py images    -> tests/golden/image_*.npz    16x16-block statistics of whole renders
  python tools/gen_golden.py scenes    -> tests/golden/scenes.npz     per-path radiance on the reference's commented alternate scenes (scenes/*.txt)
  python tools/gen_golden.py march     -> tests/golden/march.npz      rayMarching3 (rayMarchingMethods.h:330) on fixed rays
  python tools/gen_golden.py dielectric -> tests/golden/dielectric.npz per-path radiance with material-2 (dielectric) spheres in the scene
  python tools/gen_golden.py volume    -> tests/golden/volume_spheres.npz explicitPathRecursive2 (vptShadeMethods.h:398) with material-3 spheres
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import L0, L1, DEFAULT_SCENE, CAM_O, scene_without  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
SA, SS = 0.001, 0.009


def u24(rng, *shape):
    """uniforms on the 24-bit grid the product's Philox stream uses (exact in fp32 and fp64)"""
    return rng.integers(0, 1 << 24, size=shape).astype(np.float64) / float(1 << 24)


def unit_vec(rng, n):
    v = rng.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def rand_rays(rng, l1, n):
    o = np.zeros((n, 3)); d = np.zeros((n, 3))
    for i in range(n):
        if i % 2 == 0:
            o[i] = CAM_O; d[i] = l1.camera_ray(1024, 768, int(rng.integers(1024)), int(rng.integers(768)), rng.random(), rng.random())
        else:
            o[i] = [rng.uniform(-45, 45), rng.uniform(-38, 38), rng.uniform(-75, 150)]; d[i] = unit_vec(rng, 1)[0]
    return o, d


def surface_points(rng, l0, n):
    """points ON scene surfaces (first hits of random rays) with normal, incoming direction and object id"""
    pts = []
    while len(pts) < n:
        o = np.array([rng.uniform(-45, 45), rng.uniform(-38, 38), rng.uniform(-75, 150)]); d = unit_vec(rng, 1)[0]
        hit, t, i = l0.intersect(o, d)
        if not hit or i in (7, 8, 9):
            continue
        x = o + d * t
        nrm = x - DEFAULT_SCENE[i, 1:4]; nrm /= np.linalg.norm(nrm)
        pts.append((i, x, nrm, d))
    return pts


def gen_units():
    l0 = L0(); l1 = L1(); rng = np.random.default_rng(20261018)
    l0.reset_scene(); l0.set_quirks(3)
    out = {}
    N = 256
    # Sphere::intersect (Sphere.h:27) and intersect (pathTracingUtilities.h:12)
    o, d = rand_rays(rng, l1, N)
    idx = rng.integers(0, 10, N)
    out["sphere_intersect_in"] = np.concatenate([idx[:, None].astype(float), o, d], axis=1)
    out["sphere_intersect_out"] = np.array([[l0.sphere_intersect(int(i), a, b)] for i, a, b in zip(idx, o, d)])
    out["intersect_in"] = np.concatenate([o, d], axis=1)
    out["intersect_out"] = np.array([[*map(float, l0.intersect(a, b))] for a, b in zip(o, d)])
    # visibility of points in free space (pathTracingUtilities.h:39): rounding plays no role there
    x = np.stack([rng.uniform(-45, 45, N), rng.uniform(-38, 38, N), rng.uniform(-75, 150, N)], axis=1)
    lights = DEFAULT_SCENE[rng.choice([7, 8, 9], N), 1:4]
    out["visibility_in"] = np.concatenate([lights, x], axis=1)
    out["visibility_out"] = np.array([[float(l0.visibility(a, b))] for a, b in zip(lights, x)])
    # transmitance (volumetricBasicFunctions.h:14)
    x2 = x + rng.normal(size=(N, 3)) * 60; st = rng.choice([0.01, 0.05, 0.1, 0.002], N)
    out["transmittance_in"] = np.concatenate([x, x2, st[:, None]], axis=1)
    out["transmittance_out"] = np.array([[l0.transmitance(a, b, s)] for a, b, s in zip(x, x2, st)])
    # free flight (vptSamplingFunctions.h:11-31)
    xi = u24(rng, N)
    rows = []
    for s, u in zip(st, xi):
        dd = l0.freeFlightSample(s, [u]); rows.append([dd, l0.freeFlightProb(s, dd), l0.pdfSuccess(s, dd), l0.pdfFailure(s, dd)])
    out["free_flight_in"] = np.stack([st, xi], axis=1); out["free_flight_out"] = np.array(rows)
    # isotropic phase (vptSamplingFunctions.h:34)
    u2 = u24(rng, N, 2)
    out["phase_sample_in"] = u2; out["phase_sample_out"] = np.array([l0.isotropicPhaseSample(u) for u in u2])
    # equi-angular (volumetricBasicFunctions.h:209, vptSamplingFunctions.h:60)
    src = rng.choice([7, 8, 9], N); tmax = rng.uniform(5, 400, N); xi = u24(rng, N)
    rows = []
    for s, tm, a, b, u in zip(src, tmax, o, d, xi):
        r, o4 = l0.equiAngularParams2(int(s), tm, a, b, [u]); rows.append([*o4, r, l0.equiAngularProb(o4[0], o4[1], o4[2], o4[3])])
    out["equiangular_in"] = np.concatenate([src[:, None].astype(float), tmax[:, None], o, d, xi[:, None]], axis=1); out["equiangular_out"] = np.array(rows)
    # power heuristic (misSamplingFunctions.h:12)
    fg = np.exp(rng.uniform(-6, 6, (N, 2)))
    out["power_heuristic_in"] = fg; out["power_heuristic_out"] = np.array([[l0.powerHeuristics(f, g)] for f, g in fg])
    # cosine hemisphere (samplingFunctions.h:47, :92)
    nrm = unit_vec(rng, N); u2 = u24(rng, N, 2)
    rows = []
    for n_, u in zip(nrm, u2):
        w = l0.cosineHemispheric(n_, u); rows.append([*w, l0.hemiCosineProb(float(n_ @ w))])
    out["cosine_hemisphere_in"] = np.concatenate([nrm, u2], axis=1); out["cosine_hemisphere_out"] = np.array(rows)
    # cone sampling (samplingFunctions.h:65, :85)
    wc = unit_vec(rng, N); r = rng.choice([2.0, 16.5, 0.5], N); dist = r * np.exp(rng.uniform(0.2, 5, N)); u2 = u24(rng, N, 2)
    rows = []
    for w_, r_, d_, u in zip(wc, r, dist, u2):
        cm = np.sqrt(1 - (r_ / d_) * (r_ / d_)); rows.append([*l0.solidAngleDir(w_, cm, u), l0.solidAngleProb(cm)])
    out["cone_sample_in"] = np.concatenate([wc, r[:, None], dist[:, None], u2], axis=1); out["cone_sample_out"] = np.array(rows)
    # microfacet model (microFacetUtilities.h): local frame, wh = normalize(wi + wo)
    eta = DEFAULT_SCENE[5, 11:14]; kap = DEFAULT_SCENE[5, 14:17]
    rows_in, rows_out = [], []
    while len(rows_in) < N:
        alpha = float(rng.choice([0.09, 0.3, 0.03]))
        wo = unit_vec(rng, 1)[0]; wo[2] = abs(wo[2]) + 0.05; wo /= np.linalg.norm(wo)
        wh = l0.vectorFacet(alpha, u24(rng, 2))
        wi = -wo + wh * 2 * (wh @ wo); wi /= np.linalg.norm(wi)
        if wi[2] <= 0.02:
            continue
        whn = (wi + wo) / np.linalg.norm(wi + wo); nl = [0, 0, 1]
        fr = l0.frMicroFacet(eta, kap, wi, whn, wo, alpha, nl)
        rows_in.append([*eta, *kap, alpha, *wi, *wo])
        rows_out.append([*fr, l0.microFacetProb(wo, whn, alpha, nl), l0.NDF(float(whn[2]), alpha), l0.G_smith(nl, wi, wo, whn, alpha)])
    out["microfacet_in"] = np.array(rows_in); out["microfacet_out"] = np.array(rows_out)
    # vectorFacet (microFacetUtilities.h:71)
    al = rng.choice([0.09, 0.3, 0.03], N); u2 = u24(rng, N, 2)
    out["facet_normal_in"] = np.concatenate([al[:, None], u2], axis=1); out["facet_normal_out"] = np.array([l0.vectorFacet(a, u) for a, u in zip(al, u2)])
    # in-medium NEE on the AREA lights (volumetricBasicFunctions.h:284 / :225); the point light is rounding-decided -> robust variant below
    srcA = rng.choice([7, 9], N); u2 = u24(rng, N, 2); T = np.where(rng.random(N) < 0.5, -1.0, rng.uniform(0.05, 1, N))
    rows = []
    for x_, s_, u, t_ in zip(x, srcA, u2, T):
        rows.append(l0.freeSingleScattering(x_, int(s_), 0.01, 1 / 3, u) if t_ < 0 else l0.singleScattering(x_, int(s_), 0.01, SS, t_, 1 / 3, u))
    out["medium_nee_in"] = np.concatenate([x, srcA[:, None].astype(float), np.full((N, 1), 0.01), np.full((N, 1), SS), T[:, None], np.full((N, 1), 1 / 3), u2], axis=1)
    out["medium_nee_out"] = np.array(rows)
    # ... and on the point light with both hooks on ("robust" semantics = the fp32 product semantics)
    l0.set_quirks(0)
    rows = []
    for x_, u, t_ in zip(x, u2, T):
        rows.append(l0.freeSingleScattering(x_, 8, 0.01, 1 / 3, u) if t_ < 0 else l0.singleScattering(x_, 8, 0.01, SS, t_, 1 / 3, u))
    out["medium_nee_point_robust_in"] = np.concatenate([x, np.full((N, 1), 8.0), np.full((N, 1), 0.01), np.full((N, 1), SS), T[:, None], np.full((N, 1), 1 / 3), u2], axis=1)
    out["medium_nee_point_robust_out"] = np.array(rows)
    # surface functions at real surface points; robust hooks (the point-light term is rounding-decided otherwise)
    pts = surface_points(rng, l0, N)
    rows_in, rows_out = [], []
    for (i, xs, nn, wray) in pts:
        rows_in.append([i, *xs, *nn, *wray, 8]); rows_out.append(l0.pLight(i, xs, nn, wray, DEFAULT_SCENE[8, 7:10], DEFAULT_SCENE[8, 1:4], DEFAULT_SCENE[i, 17]))
    out["point_light_robust_in"] = np.array(rows_in, dtype=float); out["point_light_robust_out"] = np.array(rows_out)
    rows_in, rows_out = [], []
    for (i, xs, nn, wray) in pts:
        u = u24(rng, 8)
        rows_in.append([i, *xs, *nn, *wray, 0.01, *u]); rows_out.append(l0.MISv2(i, xs, nn, wray, DEFAULT_SCENE[i, 17], 0.01, u))
    out["surface_mis_in"] = np.array(rows_in, dtype=float); out["surface_mis_out"] = np.array(rows_out)
    rows_in, rows_out = [], []
    for (i, xs, nn, wray) in pts:
        u = u24(rng, 2)
        fs, wi, pr = l0.bdsf(wray, nn, i, u)
        win = wi / np.linalg.norm(wi)
        rows_in.append([i, *nn, *wray, *u]); rows_out.append([*(fs * float(nn @ win) / pr), *win])
    out["bsdf_sample_in"] = np.array(rows_in, dtype=float); out["bsdf_sample_out"] = np.array(rows_out)
    l0.set_quirks(3)
    # camera rays (rt.cpp:787)
    xy = np.stack([rng.integers(0, 1024, N), rng.integers(0, 768, N)], axis=1).astype(float); u2 = u24(rng, N, 2)
    out["camera_ray_in"] = np.concatenate([xy, u2], axis=1)
    out["camera_ray_out"] = np.array([l0.camera_ray(1024, 768, int(a), int(b), u[0], u[1]) for (a, b), u in zip(xy, u2)])
    # tonemap (mathUtilities.h:43)
    tv = np.concatenate([rng.uniform(-0.2, 1.3, 500), [0, 1, 0.5, 0.001, 2, 1e-9, 0.999999]])
    out["tonemap_in"] = tv; out["tonemap_out"] = np.array([l0.toDisplayValue(t) for t in tv])
    np.savez_compressed(os.path.join(GOLD, "units.npz"), **out)
    print("units.npz:", {k: v.shape for k, v in out.items()})


def gen_paths():
    l0 = L0(); l1 = L1(); rng = np.random.default_rng(77)
    out = {}
    N = 400
    o, d = rand_rays(rng, l1, N)
    seeds = rng.integers(0, 65536, (N, 3))
    out["o"], out["d"], out["seeds"] = o, d, seeds
    for quirks in (3, 0):
        l0.reset_scene(); l0.set_quirks(quirks)
        for method in (0, 1, 2):
            res = np.zeros((N, 4))
            for i in range(N):
                L, nd = l0.radiance(method, o[i], d[i], SA, SS, seed3=tuple(int(s) for s in seeds[i]))
                res[i, :3] = L; res[i, 3] = nd
            out["q%d_m%d" % (quirks, method)] = res
    # scene variant: no point light (hazards cannot arise)
    l0.set_scene(scene_without([8])); l0.set_quirks(3)
    for method in (0, 1, 2):
        res = np.zeros((N, 4))
        for i in range(N):
            L, nd = l0.radiance(method, o[i], d[i], SA, SS, seed3=tuple(int(s) for s in seeds[i]))
            res[i, :3] = L; res[i, 3] = nd
        out["no8_m%d" % method] = res
    l0.reset_scene()
    # known-answer values listed in SURVEY.md section 8c
    v = np.array([-.3, .2, -.9]); v /= np.linalg.norm(v)
    out["kat_r2_o"] = np.array([10., -5., 50.]); out["kat_r2_d"] = v
    out["kat_free_123"] = l0.radiance(0, [10, -5, 50], v, SA, SS, seed3=(1, 2, 3))[0]
    out["kat_free_567"] = l0.radiance(0, [10, -5, 50], v, SA, SS, seed3=(5, 6, 7))[0]
    out["kat_equi_567"] = l0.radiance(1, [10, -5, 50], v, SA, SS, seed3=(5, 6, 7))[0]
    out["kat_mis_567"] = l0.radiance(2, [10, -5, 50], v, SA, SS, seed3=(5, 6, 7))[0]
    np.savez_compressed(os.path.join(GOLD, "paths.npz"), **out)
    print("paths.npz written")


def block_stats(total, sq, spp, block=16):
    h, w, _ = total.shape
    mean_px = total / spp
    var_px = np.maximum(sq / spp - mean_px ** 2, 0) / max(spp - 1, 1)  # variance of each pixel's mean
    bh, bw = h // block, w // block
    m = mean_px[:bh * block, :bw * block].reshape(bh, block, bw, block, 3).mean(axis=(1, 3))
    v = var_px[:bh * block, :bw * block].reshape(bh, block, bw, block, 3).sum(axis=(1, 3)) / (block * block) ** 2
    return m.astype(np.float32), v.astype(np.float32)


ALT_SCENES = ["scene2_sigma", "scene3_near_camera", "scene4_area_light", "scene5_infinite", "scene6_two_points"]


def gen_scenes():
    """the UNMODIFIED reference on its five commented alternate scenes (include/Sphere.cpp:27-106, as data in scenes/*.txt): 120 seeded paths per
    scene and method, as shipped (quirks 3) and with the robust hooks (quirks 0)"""
    l0 = L0(); l1 = L1(); rng = np.random.default_rng(909)
    N = 120
    o, d = rand_rays(rng, l1, N)
    seeds = rng.integers(0, 65536, (N, 3))
    out = {"o": o, "d": d, "seeds": seeds}
    for name in ALT_SCENES:
        rows = np.loadtxt(os.path.join(ROOT, "scenes", name + ".txt"), comments="#")
        out["rows_" + name] = rows
        l0.set_scene(rows)
        for quirks in (3, 0):
            l0.set_quirks(quirks)
            for method in (0, 1, 2):
                res = np.zeros((N, 4))
                for i in range(N):
                    L, nd = l0.radiance(method, o[i], d[i], SA, SS, seed3=tuple(int(s) for s in seeds[i]))
                    res[i, :3] = L; res[i, 3] = nd
                out["%s_q%d_m%d" % (name, quirks, method)] = res
    l0.reset_scene(); l0.set_quirks(3)
    np.savez_compressed(os.path.join(GOLD, "scenes.npz"), **out)
    print("scenes.npz written:", len(out), "arrays")


def dielectric_scenes():
    """material 2 appears in no scene of Sphere.cpp; its code (bdsf vptShadeMethods.h:26-46, softDielectric samplingFunctions.h:209, the
    dielectric branches of MISv2 misSamplingFunctions.h:110-118,144-152) is reachable by giving a sphere material 2: the blue ball, and both balls"""
    a = DEFAULT_SCENE.copy(); a[6, 10] = 2
    b = DEFAULT_SCENE.copy(); b[5, 10] = 2; b[6, 10] = 2
    return {"glass6": a, "glass56": b}


def gen_dielectric():
    """the UNMODIFIED reference with dielectric spheres: 240 seeded paths (half aimed at the ball) per scene, method and quirk setting"""
    l0 = L0(); l1 = L1(); rng = np.random.default_rng(2024)
    N = 240
    o, d = rand_rays(rng, l1, N)
    for i in range(0, N, 2):
        v = DEFAULT_SCENE[6 if i % 4 == 0 else 5, 1:4] + rng.normal(size=3) * 8 - o[i]; d[i] = v / np.linalg.norm(v)
    seeds = rng.integers(0, 65536, (N, 3))
    out = {"o": o, "d": d, "seeds": seeds}
    for name, rows in dielectric_scenes().items():
        out["rows_" + name] = rows
        l0.set_scene(rows)
        for quirks in (3, 0):
            l0.set_quirks(quirks)
            for method in (0, 1, 2):
                res = np.zeros((N, 4))
                for i in range(N):
                    L, nd = l0.radiance(method, o[i], d[i], SA, SS, seed3=tuple(int(s) for s in seeds[i]))
                    res[i, :3] = L; res[i, 3] = nd
                out["%s_q%d_m%d" % (name, quirks, method)] = res
    l0.reset_scene(); l0.set_quirks(3)
    np.savez_compressed(os.path.join(GOLD, "dielectric.npz"), **out)
    print("dielectric.npz written:", len(out), "arrays")


def gen_volume():
    """explicitPathRecursive2 of the UNMODIFIED reference (vptShadeMethods.h:398-497, as shipped: quirks 3) on scenes/scene_volume_spheres.txt:
    320 seeded camera paths, a third of them aimed at the two volumetric (material 3) spheres, with the number of erand48 draws; plus the
    16x16-block statistics of a 256x192 render at 64 spp (the legacy estimator takes no medium parameters: its sigma_a = 0.05, sigma_s = 0.009
    and roulette q = 0.1 are literals)"""
    l0 = L0(); l1 = L1(); rng = np.random.default_rng(777)
    rows = np.loadtxt(os.path.join(ROOT, "scenes", "scene_volume_spheres.txt"), comments="#")
    N = 320
    o = np.tile(np.array(CAM_O), (N, 1)); d = np.zeros((N, 3))
    vols = rows[rows[:, 10] == 3]
    for i in range(N):
        if i % 3 == 0:
            v = vols[(i // 3) % len(vols)]
            aim = v[1:4] + rng.normal(size=3) * 0.5 * v[0] - o[i]; d[i] = aim / np.linalg.norm(aim)
        else:
            d[i] = l1.camera_ray(1024, 768, int(rng.integers(1024)), int(rng.integers(768)), rng.random(), rng.random())
    seeds = rng.integers(0, 65536, (N, 3))
    l0.set_scene(rows); l0.set_quirks(3)
    res = np.zeros((N, 4))
    for i in range(N):
        L, nd = l0.radiance(5, o[i], d[i], SA, SS, seed3=tuple(int(s) for s in seeds[i]))
        res[i, :3] = L; res[i, 3] = nd
    w, h, spp = 256, 192, 64
    t0 = time.time()
    total, sq, draws = l0.render(w, h, spp, 5, SA, SS, seed=31, want_sumsq=True)
    mean, var = block_stats(total, sq, spp)
    l0.reset_scene(); l0.set_quirks(3)
    np.savez_compressed(os.path.join(GOLD, "volume_spheres.npz"), rows=rows, o=o, d=d, seeds=seeds, q3=res, block_mean=mean.astype(np.float32),
                        block_var=var.astype(np.float32), width=w, height=h, spp=spp, image_mean=(total / spp).mean(axis=(0, 1)))
    print("volume_spheres.npz: %d paths, %d with light, render %.1fs, image mean %s" % (N, int((res[:, :3].max(axis=1) > 0).sum()), time.time() - t0, (total / spp).mean(axis=(0, 1))))


def gen_march():
    """rayMarching3 of the UNMODIFIED reference (as shipped: quirks 3) and with the robust hooks (quirks 0) on 96 rays: the literals of the
    commented call rt.cpp:791 (sigma 0.001 / 0.0125, step 0.1, source 7), the point light (source 8) and a coarser step"""
    l0 = L0(); l1 = L1(); rng = np.random.default_rng(4242)
    l0.reset_scene()
    o, d = rand_rays(rng, l1, 96)
    out = {"o": o, "d": d}
    cases = [("src7_step01", 0.001, 0.0125, 0.1, 7), ("src8_step01", 0.001, 0.0125, 0.1, 8), ("src8_step05", 0.001, 0.009, 0.5, 8), ("src9_step1", 0.002, 0.02, 1.0, 9)]
    out["cases"] = np.array([[sa, ss, step, src] for _, sa, ss, step, src in cases])
    for q in (3, 0):
        l0.set_quirks(q)
        for name, sa, ss, step, src in cases:
            out["q%d_%s" % (q, name)] = np.array([l0.rayMarching3(o[i], d[i], sa, ss, step, src) for i in range(len(o))])
    l0.set_quirks(3)
    np.savez_compressed(os.path.join(GOLD, "march.npz"), **out)
    print("march.npz:", {k: v.shape for k, v in out.items()})


def gen_images(which=None):
    l0 = L0()
    jobs = [  # name, scene, quirks, method, spp
        ("strict_m0", None, 3, 0, 512), ("strict_m1", None, 3, 1, 256), ("strict_m2", None, 3, 2, 256),
        ("robust_m0", None, 0, 0, 256), ("robust_m1", None, 0, 1, 256), ("robust_m2", None, 0, 2, 256),
        ("no8_m0", [8], 3, 0, 256), ("no8_m1", [8], 3, 1, 256), ("no8_m2", [8], 3, 2, 256),
        # north_star's correctness render: the "MIS" method at 4096 spp (14 minutes of 8 CPU threads for the 1024x768 frame)
        ("robust_m2_4096", None, 0, 2, 4096),
        # BASELINE.json config 3 as written: 1920x1080 at 4096 spp (half an hour of 8 CPU threads)
        ("robust_m2_c3", None, 0, 2, 4096, 1920, 1080),
    ]
    for name, drop, quirks, method, spp, *size in jobs:
        w, h = size if size else (1024, 768)
        if which and name not in which:
            continue
        path = os.path.join(GOLD, "image_%s.npz" % name)
        if os.path.exists(path) and not which:
            print("skip", name); continue
        l0.reset_scene()
        if drop:
            l0.set_scene(scene_without(drop))
        l0.set_quirks(quirks)
        t0 = time.time()
        total, sq, draws = l0.render(w, h, spp, method, SA, SS, seed=1000 + method, nthreads=int(os.environ.get("GOLD_THREADS", "0")))
        dt = time.time() - t0
        m, v = block_stats(total, sq, spp)
        np.savez_compressed(path, block_mean=m, block_var=v, width=w, height=h, spp=spp, method=method, quirks=quirks,
                            dropped=np.array(drop or [], dtype=np.int64), image_mean=(total / spp).mean(axis=(0, 1)), seconds=dt,
                            mpaths_per_s=w * h * spp / dt / 1e6, draws_per_path=draws / (w * h * spp))
        print(name, "%.1fs" % dt, "%.2f Mpaths/s" % (w * h * spp / dt / 1e6), "mean", (total / spp).mean(axis=(0, 1)), flush=True)
    l0.reset_scene(); l0.set_quirks(3)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("units", "all"):
        gen_units()
    if what in ("paths", "all"):
        gen_paths()
    if what in ("scenes", "all"):
        gen_scenes()
    if what in ("dielectric", "all"):
        gen_dielectric()
    if what in ("march", "all"):
        gen_march()
    if what in ("volume", "all"):
        gen_volume()
    if what in ("images", "all"):
        gen_images(sys.argv[2:] or None)
