"""GPU: whole camera paths (the three shade methods) through the C-ABI against
 (a) the unmodified reference's own per-path results on its seeded erand48 sequences (tests/golden/paths.npz), and
 (b) the FP64 CPU oracle on the product's Philox streams (common random numbers)."""
import numpy as np
import pytest

from oracle_lib import DEFAULT_SCENE, CAM_O, scene_without

pytestmark = pytest.mark.gpu
SA, SS = 0.001, 0.009


def erand48_stream(seed3, n):
    """POSIX erand48 (the reference's generator, Vector.h:38): X' = (0x5DEECE66D X + 0xB) mod 2^48, returns X' / 2^48"""
    x = int(seed3[0]) | (int(seed3[1]) << 16) | (int(seed3[2]) << 32)
    out = np.empty(n)
    for i in range(n):
        x = (0x5DEECE66D * x + 0xB) & ((1 << 48) - 1)
        out[i] = x / float(1 << 48)
    return out


def list_rows(paths):
    n = len(paths["o"])
    rows = np.zeros((n, 127))
    rows[:, 0:3] = paths["o"]; rows[:, 3:6] = paths["d"]; rows[:, 6] = 120
    for i in range(n):
        rows[i, 7:] = erand48_stream(paths["seeds"][i], 120)
    return rows


def rel(got, want):
    den = np.maximum(np.abs(want).max(axis=1), 1e-30)
    e = np.abs(got - want).max(axis=1) / den
    e[(np.abs(want).max(axis=1) == 0) & (np.abs(got).max(axis=1) == 0)] = 0
    return e


@pytest.mark.parametrize("method", [0, 1, 2])
def test_fp64_ref_mode_reproduces_reference_paths(gpu, paths, method):
    """REF mode, robust semantics: nothing is rounding-decided, so the device must give the reference's numbers"""
    rows = list_rows(paths)
    p = gpu.default_params(method=method, precision=gpu.PRECISION_FP64_REF, quirks=0)
    got = gpu.unit(gpu.UNIT.RADIANCE_LIST, rows, p)
    want = paths["q0_m%d" % method]
    ok = got[:, 3] >= 0
    assert ok.mean() > 0.97                                   # a few paths need more than 120 draws
    assert np.array_equal(got[ok, 3], want[ok, 3])            # consumed exactly the reference's number of draws
    assert rel(got[ok, :3], want[ok, :3]).max() < 1e-9        # libm (CUDA vs glibc) differences only


@pytest.mark.parametrize("method", [0, 1, 2])
def test_fp64_ref_mode_with_reference_quirks(gpu, paths, method):
    """quirks on: outcomes that hinge on FP64 rounding can differ once a CUDA libm result differs by an ulp from glibc's;
    everything else must be identical, and in the scene without the point light nothing may differ at all"""
    rows = list_rows(paths)
    p = gpu.default_params(method=method, precision=gpu.PRECISION_FP64_REF, quirks=3)
    got = gpu.unit(gpu.UNIT.RADIANCE_LIST, rows, p)
    want = paths["q3_m%d" % method]
    ok = got[:, 3] >= 0
    assert np.array_equal(got[ok, 3], want[ok, 3])
    e = rel(got[ok, :3], want[ok, :3])
    assert np.mean(e > 1e-9) < 0.08
    assert np.all(rel(got[ok, 1:3], want[ok, 1:3]) < 1e-9)    # the point light is red only: green and blue never hinge on it
    no8 = gpu.scene_from_rows(scene_without([8]))
    got = gpu.unit(gpu.UNIT.RADIANCE_LIST, rows, p, no8)
    want = paths["no8_m%d" % method]
    ok = got[:, 3] >= 0
    assert np.array_equal(got[ok, 3], want[ok, 3]) and rel(got[ok, :3], want[ok, :3]).max() < 1e-9


@pytest.mark.parametrize("method", [0, 1, 2])
def test_fp32_paths_against_reference_vectors(gpu, paths, method):
    """the performance path on the reference's own random sequences (robust-hook vectors and the no-point-light scene)"""
    rows = list_rows(paths)
    p = gpu.default_params(method=method)
    for scene, key in ((None, "q0_m%d" % method), (gpu.scene_from_rows(scene_without([8])), "no8_m%d" % method)):
        got = gpu.unit(gpu.UNIT.RADIANCE_LIST, rows, p, scene)
        want = paths[key]
        ok = got[:, 3] >= 0
        same = got[ok, 3] == want[ok, 3]
        assert same.mean() > 0.985                            # a flipped hit/miss decision changes the path
        e = rel(got[ok, :3], want[ok, :3])[same]
        assert np.median(e[e > 0]) < 2e-6 and np.mean(e > 1e-4) < 0.01


def random_rays(l1, n, seed):
    rng = np.random.default_rng(seed)
    o = np.zeros((n, 3)); d = np.zeros((n, 3))
    for i in range(n):
        if i % 2 == 0:
            o[i] = CAM_O; d[i] = l1.camera_ray(1024, 768, int(rng.integers(1024)), int(rng.integers(768)), rng.random(), rng.random())
        else:
            o[i] = [rng.uniform(-45, 45), rng.uniform(-38, 38), rng.uniform(-75, 150)]; x = rng.normal(size=3); d[i] = x / np.linalg.norm(x)
    pix = rng.integers(0, 2 ** 20, n).astype(np.uint32); smp = rng.integers(0, 2 ** 14, n).astype(np.uint32)
    return o, d, pix, smp


@pytest.mark.parametrize("method", [0, 1, 2, 4])
def test_philox_paths_common_random_numbers(gpu, l1, method):
    n = 30000
    o, d, pix, smp = random_rays(l1, n, 11 + method)
    rows = np.concatenate([o, d, pix[:, None].astype(float), smp[:, None].astype(float)], axis=1)
    want, ev = l1.radiance_philox(DEFAULT_SCENE, 0, method, SA, SS, 42, o, d, pix, smp)
    # FP64 REF, robust: identical streams, identical decisions
    got = gpu.unit(gpu.UNIT.RADIANCE, rows, gpu.default_params(method=method, precision=gpu.PRECISION_FP64_REF, quirks=0, seed=42))
    assert np.array_equal(got[:, 3], ev)
    assert rel(got[:, :3], want).max() < 1e-9
    # FP32: same streams; decisions flip only at fp32-resolution boundaries
    got = gpu.unit(gpu.UNIT.RADIANCE, rows, gpu.default_params(method=method, seed=42))
    assert not np.isnan(got).any()
    assert np.mean(got[:, 3] == ev) > 0.995
    e = rel(got[:, :3], want)
    assert np.median(e[e > 0]) < 1e-6 and np.quantile(e, 0.99) < 1e-4 and np.mean(e > 1e-3) < 0.006
    # the sample mean agrees far inside the Monte Carlo noise because the samples are shared
    np.testing.assert_allclose(got[:, :3].mean(axis=0), want.mean(axis=0), rtol=2e-3)


def test_parameters_not_in_the_reference(gpu, l1):
    """continue_prob and max_depth are literals in the reference (vptShadeMethods.h:1275); config 4 of BASELINE.json needs
    them as parameters, and only the oracle can say what they should do"""
    n = 8000
    o, d, pix, smp = random_rays(l1, n, 5)
    rows = np.concatenate([o, d, pix[:, None].astype(float), smp[:, None].astype(float)], axis=1)
    kw = dict(sigma_a=0.0005, sigma_s=0.0495, continue_prob=0.95, max_depth=64)
    for method in (0, 2):
        want, ev = l1.radiance_philox(DEFAULT_SCENE, 0, method, kw["sigma_a"], kw["sigma_s"], 3, o, d, pix, smp, cp=0.95, max_depth=64)
        got = gpu.unit(gpu.UNIT.RADIANCE, rows, gpu.default_params(method=method, precision=gpu.PRECISION_FP64_REF, seed=3, **kw))
        assert np.array_equal(got[:, 3], ev) and ev.max() <= 64 and ev.mean() > 10
        assert rel(got[:, :3], want).max() < 1e-8
        got = gpu.unit(gpu.UNIT.RADIANCE, rows, gpu.default_params(method=method, seed=3, **kw))
        assert np.mean(got[:, 3] == ev) > 0.97               # long paths: more chances for one fp32 decision to flip
        np.testing.assert_allclose(got[:, :3].mean(axis=0), want.mean(axis=0), rtol=0.02)
