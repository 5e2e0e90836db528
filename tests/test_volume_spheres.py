"""VPT_METHOD_VOLUME_SPHERES = explicitPathRecursive2 (vptShadeMethods.h:398-497), the reference's only estimator that handles material 3
(volumetric spheres): multipleT volumetricBasicFunctions.h:26-58, Sphere::intersectVPT Sphere.h:39-45, intersectV2
:109-134, punctualVolumetric rayMarchingMethods.h:12-32, pLight's visibilityVPT branch vptShadeMethods.h:70-74.
  CPU: the FP64 restatement against the UNMODIFIED reference on its own erand48 sequences (tests/golden/volume_spheres.npz; live where
       oracle/_ref exists), and the C-ABI's acceptance rules for material 3;
  GPU: the FP64 reference-mode kernels against those vectors, against the oracle on Philox streams, and against the reference's render."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle_lib import CAM_O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENE_FILE = os.path.join(ROOT, "scenes", "scene_volume_spheres.txt")
M = 5  # VPT_METHOD_VOLUME_SPHERES


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLDEN, "volume_spheres.npz")))


def test_oracle_matches_the_reference_vectors(l1, gold):
    """same operations in the same order (the recursion unrolled into throughput form): agreement to summation rounding, identical draw counts"""
    rows, want = gold["rows"], gold["q3"]
    assert (rows[:, 10] == 3).sum() == 2 and (rows[:, 0] == 0).sum() == 2
    worst = 0.0
    for i in range(len(want)):
        got, draws = l1.radiance_erand48(rows, 3, M, 0.001, 0.009, gold["o"][i], gold["d"][i], tuple(int(s) for s in gold["seeds"][i]))
        assert draws == want[i, 3]
        ref = want[i, :3]
        worst = max(worst, float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300) * (ref != got))))
    assert worst < 1e-13
    assert (want[:, :3].max(axis=1) > 0).mean() > 0.9          # the scene is lit


def test_oracle_live_against_the_reference(l0, l1, gold):
    rows = gold["rows"]
    rng = np.random.default_rng(12)
    l0.set_scene(rows); l0.set_quirks(3)
    try:
        for _ in range(300):
            d = l1.camera_ray(1024, 768, int(rng.integers(1024)), int(rng.integers(768)), rng.random(), rng.random())
            seed = tuple(int(x) for x in rng.integers(0, 65536, 3))
            a, na = l0.radiance(M, np.array(CAM_O), d, 0.001, 0.009, seed3=seed)
            b, nb = l1.radiance_erand48(rows, 3, M, 0.001, 0.009, np.array(CAM_O), d, seed)
            assert na == nb
            np.testing.assert_allclose(b, a, rtol=1e-13, atol=0)
    finally:
        l0.reset_scene(); l0.set_quirks(3)


def test_the_volumetric_spheres_do_something(l1, gold):
    """rays aimed at a material-3 sphere collect in-scattered light inside it and are attenuated behind it: removing the two spheres changes them"""
    rows = gold["rows"]
    bare = rows[rows[:, 10] != 3]
    changed = 0
    aimed = range(0, 60, 3)                                                       # the generator aims every third ray at a volumetric sphere
    for i in aimed:
        seed = tuple(int(s) for s in gold["seeds"][i])
        a, _ = l1.radiance_erand48(rows, 3, M, 0.001, 0.009, gold["o"][i], gold["d"][i], seed)
        b, _ = l1.radiance_erand48(bare, 3, M, 0.001, 0.009, gold["o"][i], gold["d"][i], seed)
        changed += not np.allclose(a, b, rtol=1e-6)
    assert changed >= 0.8 * len(aimed)


def test_material_3_is_accepted_only_by_its_own_method(vpt):
    scene = vpt.load_scene(SCENE_FILE)
    assert len(scene) == 9 and sorted(s.material for s in scene).count(3) == 2
    lib = vpt.load_library()
    buf = np.zeros((8, 8, 3), dtype=np.float32)

    def rc(**kw):
        p = vpt.default_params(width=8, height=8, spp=1, **kw)
        return lib.vpt_render(C.byref(p), scene, len(scene), buf.ctypes.data_as(C.POINTER(C.c_float)), None)
    for method in (0, 1, 2, 4):                                                   # bdsf leaves pdf and direction unset for material 3 there
        assert rc(method=method) == -3 and rc(method=method, precision=vpt.PRECISION_FP64_REF, quirks=3) == -3
    assert rc(method=M) == -3                                                     # the legacy estimator exists in reference precision only
    assert rc(method=M, precision=vpt.PRECISION_FP64_REF, quirks=3) in (0, -4, -5)
    assert rc(method=6) == -1


# ---- GPU -------------------------------------------------------------------------------------------------------------------------------
def erand48_stream(seed3, n):
    x = int(seed3[0]) | (int(seed3[1]) << 16) | (int(seed3[2]) << 32)
    out = np.empty(n)
    for i in range(n):
        x = (0x5DEECE66D * x + 0xB) & ((1 << 48) - 1)
        out[i] = x / float(1 << 48)
    return out


def test_the_estimator_has_a_rounding_decided_branch_of_its_own(l1, gold):
    """Hazard C (as hazards A and B of the active methods): the first march point of a volumetric sphere lies ON the
    sphere, so in multipleT (volumetricBasicFunctions.h:42-49) the far root of that very sphere is 0 up to rounding -- `t_aux2 < 0` then
    multiplies the step's transmittance by exp(+sigma_t chord) instead of exp(-sigma_t chord).  A 1e-13 change of the camera direction flips
    it: per-path agreement between two correct FP64 implementations (glibc / CUDA libm) can only be asked of the paths that are not affected."""
    rows = gold["rows"]
    o = np.tile(np.array(CAM_O), (60, 1))
    d = gold["d"][:180:3][:60]                                                     # the rays aimed at the volumetric spheres
    pix = np.arange(60, dtype=np.uint32); smp = np.zeros(60, dtype=np.uint32)
    a, _ = l1.radiance_philox(rows, 0, M, 0.001, 0.009, 5, o, d, pix, smp)
    d2 = d + 1e-13; d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    b, _ = l1.radiance_philox(rows, 0, M, 0.001, 0.009, 5, o, d2, pix, smp)
    e = np.abs(a - b).max(axis=1) / np.maximum(np.abs(a).max(axis=1), 1e-30)
    assert np.median(e) < 1e-9                                                     # smooth for most rays ...
    assert 1e-6 < e.max() < 0.05                                                   # ... and a jump of the size of one march step for some


def _per_path(got, want):
    den = np.maximum(np.abs(want).max(axis=1), 1e-30)
    return np.abs(got - want).max(axis=1) / den


@pytest.mark.gpu
def test_gpu_fp64_reproduces_the_reference_vectors(gpu, gold):
    """the reference's own erand48 sequences (as shipped: quirks 3) through VPT_UNIT_RADIANCE_LIST: same number of draws; identical radiance
    except where one of the rounding-decided branches (hazards A, B: point-light visibility; C: above) falls differently with CUDA's libm"""
    n = len(gold["o"])
    rows = np.zeros((n, 127))
    rows[:, 0:3] = gold["o"]; rows[:, 3:6] = gold["d"]; rows[:, 6] = 120
    for i in range(n):
        rows[i, 7:] = erand48_stream(gold["seeds"][i], 120)
    p = gpu.default_params(method=M, precision=gpu.PRECISION_FP64_REF, quirks=3)
    got = gpu.unit(gpu.UNIT.RADIANCE_LIST, rows, p, gpu.scene_from_rows(gold["rows"]))
    want = gold["q3"]
    ok = got[:, 3] >= 0
    assert ok.mean() > 0.9                                                        # roulette q = 0.1: a few paths need more than 120 draws
    same = got[ok, 3] == want[ok, 3]
    assert same.mean() > 0.97
    e = _per_path(got[ok, :3], want[ok, :3])[same]
    assert np.median(e) < 1e-12 and np.mean(e > 1e-9) < 0.2, (float(np.median(e)), float(np.mean(e > 1e-9)))


@pytest.mark.gpu
def test_gpu_fp64_paths_equal_the_oracle_on_philox_streams(gpu, l1, gold):
    """robust semantics (quirks 0): only hazard C is left -- at most 1 % of the paths, by at most a march step's worth"""
    rows = gold["rows"]
    n = 6000
    rng = np.random.default_rng(1)
    o = np.tile(np.array(CAM_O), (n, 1))
    d = np.array([l1.camera_ray(256, 192, int(rng.integers(256)), int(rng.integers(192)), rng.random(), rng.random()) for _ in range(n)])
    pix = rng.integers(0, 2 ** 20, n).astype(np.uint32); smp = rng.integers(0, 2 ** 14, n).astype(np.uint32)
    inp = np.concatenate([o, d, pix[:, None].astype(float), smp[:, None].astype(float)], axis=1)
    want, ev = l1.radiance_philox(rows, 0, M, 0.001, 0.009, 42, o, d, pix, smp)
    got = gpu.unit(gpu.UNIT.RADIANCE, inp, gpu.default_params(method=M, precision=gpu.PRECISION_FP64_REF, quirks=0, seed=42), gpu.scene_from_rows(rows))
    assert np.array_equal(got[:, 3], ev)
    e = _per_path(got[:, :3], want)
    assert np.mean(e > 1e-9) < 0.01 and e.max() < 0.05 and np.median(e) < 1e-12, (float(np.mean(e > 1e-9)), float(e.max()))


@pytest.mark.gpu
def test_gpu_fp64_render_equals_the_oracle_on_philox_streams(gpu, l1, gold):
    w, h, spp = 96, 72, 4
    scene = gpu.scene_from_rows(gold["rows"])
    p = gpu.default_params(width=w, height=h, spp=spp, method=M, precision=gpu.PRECISION_FP64_REF, quirks=0, seed=6, output=gpu.OUTPUT_SUM)
    img, st = gpu.render(p, scene, stats=True)
    ref, _, rst = l1.render(gold["rows"], 0, M, 0.001, 0.009, w, h, 6, spp, want_sumsq=False)
    assert st.paths == w * h * spp and st.nonfinite == 0 and st.events == rst["events"]
    e = np.abs(img - ref).max(axis=2) / np.maximum(np.abs(ref).max(axis=2), 1e-6)
    assert np.mean(e > 3e-6) < 0.03 and e.max() < 0.05                            # (hazard C pixels)
    np.testing.assert_allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=1e-4)
    assert np.array_equal(img, gpu.render(p, scene))
    img3, st3 = gpu.render(p.copy(quirks=3), scene, stats=True)                   # as shipped: hazards A and B on top (point-light visibility)
    ref3, _, rst3 = l1.render(gold["rows"], 3, M, 0.001, 0.009, w, h, 6, spp, want_sumsq=False)
    assert abs(int(st3.events) - rst3["events"]) <= 0.01 * rst3["events"]
    np.testing.assert_allclose(img3.mean(axis=(0, 1)), ref3.mean(axis=(0, 1)), rtol=0.05)


@pytest.mark.gpu
def test_gpu_render_matches_the_reference_render(gpu, gold):
    """256x192 at 64 spp by the unmodified reference (block statistics in the golden file) against 1024 spp on the GPU: whole-image mean z < 4.5"""
    w, h = int(gold["width"]), int(gold["height"])
    p = gpu.default_params(width=w, height=h, spp=1024, method=M, precision=gpu.PRECISION_FP64_REF, quirks=3, seed=2)
    img = gpu.render(p, gpu.scene_from_rows(gold["rows"])).astype(np.float64)
    bm = img[:h // 16 * 16, :w // 16 * 16].reshape(h // 16, 16, w // 16, 16, 3).mean(axis=(1, 3))
    ref, var = gold["block_mean"].astype(np.float64), gold["block_var"].astype(np.float64)
    sigma = np.sqrt((var * (1.0 + float(gold["spp"]) / 1024)).sum(axis=(0, 1))) / (var.shape[0] * var.shape[1])
    z = (bm.mean(axis=(0, 1)) - ref.mean(axis=(0, 1))) / sigma
    assert np.all(np.abs(z) < 4.5), z
    assert bm.mean() > 0.01
