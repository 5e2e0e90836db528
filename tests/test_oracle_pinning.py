"""CPU: per-path radiance of the FP64 restatement against the unmodified reference on the SAME erand48 streams
(tests/golden/paths.npz; live against oracle/_ref where it exists).  Free-flight is the same sequence of operations and must
be bit-exact; the two recursive estimators are evaluated in throughput form here, so they agree to summation rounding."""
import numpy as np
import pytest

from oracle_lib import DEFAULT_SCENE, CAM_O, scene_without

SA, SS = 0.001, 0.009


def check(l1, scene, quirks, method, o, d, seeds, want):
    worst = 0.0
    for i in range(len(o)):
        got, draws = l1.radiance_erand48(scene, quirks, method, SA, SS, o[i], d[i], tuple(int(s) for s in seeds[i]))
        assert draws == want[i, 3], "draw count differs: the restatement consumed the stream differently"
        ref = want[i, :3]
        if method == 0:
            assert np.array_equal(got, ref)
        else:
            err = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300) * (ref != got))
            worst = max(worst, err)
    assert worst < 1e-13


@pytest.mark.parametrize("method", [0, 1, 2])
@pytest.mark.parametrize("quirks", [3, 0])
def test_paths_match_reference_vectors(l1, paths, quirks, method):
    check(l1, DEFAULT_SCENE, quirks, method, paths["o"], paths["d"], paths["seeds"], paths["q%d_m%d" % (quirks, method)])


@pytest.mark.parametrize("method", [0, 1, 2])
def test_paths_without_point_light(l1, paths, method):
    sc = scene_without([8])
    want = paths["no8_m%d" % method]
    check(l1, sc, 3, method, paths["o"], paths["d"], paths["seeds"], want)
    # without the r = 0 sphere nothing is rounding-decided: both quirk settings give the same numbers
    for i in range(0, len(want), 7):
        a, _ = l1.radiance_erand48(sc, 0, method, SA, SS, paths["o"][i], paths["d"][i], tuple(int(s) for s in paths["seeds"][i]))
        b, _ = l1.radiance_erand48(sc, 3, method, SA, SS, paths["o"][i], paths["d"][i], tuple(int(s) for s in paths["seeds"][i]))
        assert np.array_equal(a, b)


def test_survey_known_answers(l1, paths):
    o, d = paths["kat_r2_o"], paths["kat_r2_d"]
    assert np.array_equal(l1.radiance_erand48(DEFAULT_SCENE, 3, 0, SA, SS, o, d, (1, 2, 3))[0], paths["kat_free_123"])
    assert paths["kat_free_123"][0] == pytest.approx(0.45655814784796867, rel=1e-14) and paths["kat_free_123"][2] == 0
    assert np.array_equal(l1.radiance_erand48(DEFAULT_SCENE, 3, 0, SA, SS, o, d, (5, 6, 7))[0], paths["kat_free_567"])
    np.testing.assert_allclose(l1.radiance_erand48(DEFAULT_SCENE, 3, 1, SA, SS, o, d, (5, 6, 7))[0], paths["kat_equi_567"], rtol=1e-13)
    np.testing.assert_allclose(l1.radiance_erand48(DEFAULT_SCENE, 3, 2, SA, SS, o, d, (5, 6, 7))[0], paths["kat_mis_567"], rtol=1e-13)
    # "MIS" is numerically the equi-angular estimator
    np.testing.assert_allclose(paths["kat_equi_567"], paths["kat_mis_567"], rtol=1e-13)


def test_live_against_reference(l0, l1):
    """fresh random rays, all quirk combinations, straight against the compiled reference (only where it was built)"""
    rng = np.random.default_rng(5)
    for quirks in (3, 0, 1, 2):
        l0.set_quirks(quirks)
        for method in (0, 1, 2):
            for _ in range(150):
                if rng.random() < 0.5:
                    o = np.array(CAM_O); d = l1.camera_ray(1024, 768, int(rng.integers(1024)), int(rng.integers(768)), rng.random(), rng.random())
                else:
                    o = np.array([rng.uniform(-45, 45), rng.uniform(-38, 38), rng.uniform(-75, 150)]); d = rng.normal(size=3); d /= np.linalg.norm(d)
                seed = tuple(int(x) for x in rng.integers(0, 65536, 3))
                a, na = l0.radiance(method, o, d, SA, SS, seed3=seed)
                b, nb = l1.radiance_erand48(DEFAULT_SCENE, quirks, method, SA, SS, o, d, seed)
                assert na == nb
                np.testing.assert_allclose(b, a, rtol=1e-13, atol=0)
    l0.set_quirks(3)


def test_philox_stream_properties(l1):
    """the Philox-driven oracle: same (seed, pixel, sample) -> same path; different sample -> different stream"""
    o = np.tile(np.array(CAM_O), (64, 1)); d = np.array([l1.camera_ray(64, 48, x, 20, 0.5, 0.5) for x in range(64)])
    pix = np.arange(64, dtype=np.uint32); smp = np.zeros(64, dtype=np.uint32)
    a, ev = l1.radiance_philox(DEFAULT_SCENE, 0, 0, SA, SS, 9, o, d, pix, smp)
    b, _ = l1.radiance_philox(DEFAULT_SCENE, 0, 0, SA, SS, 9, o, d, pix, smp)
    c, _ = l1.radiance_philox(DEFAULT_SCENE, 0, 0, SA, SS, 9, o, d, pix, smp + 1)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert ev.min() == 0 and ev.max() >= 1  # 40 % of paths die at the first roulette draw
    # max_depth truncates, continue_prob = 1 never terminates by roulette
    deep, ev2 = l1.radiance_philox(DEFAULT_SCENE, 0, 0, SA, SS, 9, o, d, pix, smp, cp=1.0, max_depth=5)
    assert ev2.max() <= 5 and ev2.min() >= 1
