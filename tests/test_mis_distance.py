"""VPT_METHOD_MIS_DISTANCE: one-sample MIS, balance heuristic, of the reference's two distance techniques
(free flight, vptSamplingFunctions.h:11-20; equi-angular, volumetricBasicFunctions.h:209-223 + vptSamplingFunctions.h:60).  The reference
names such a method (MISVPTTracerRecursive, vptShadeMethods.h:1345) but implements the equi-angular estimator again, so there is no
reference vector for it: the FP64 oracle restatement is checked through what the construction guarantees (it samples a NORMALISED
mixture density and reproduces the expectation of the reference's own methods), and the CUDA kernels are checked against that oracle."""
import numpy as np
import pytest

from oracle_lib import DEFAULT_SCENE

SA, SS = 0.001, 0.009


def _rows(rng, n, tmax, src, o, d, sigma_t):
    rows = np.zeros((n, 11))
    rows[:, 0] = src; rows[:, 1] = tmax; rows[:, 2:5] = o; rows[:, 5:8] = d; rows[:, 8] = sigma_t
    rows[:, 9] = rng.random(n); rows[:, 10] = rng.random(n)
    return rows


@pytest.mark.parametrize("case", [
    dict(o=(10, -5, 50), d=(-.3, .2, -.9), t=141.77704970529885, src=8, st=0.01),   # 141.777 is the reference's own hit distance
    dict(o=(0, 11.2, 214), d=(0, -0.042573365542992951, -0.99909334325994914), t=295.86826069832023, src=9, st=0.01),  # camera centre ray
    dict(o=(-20, 20, 5), d=(0, 0, -1), t=80.0, src=8, st=0.1),                       # passes 3 units from the point light, dense fog
    dict(o=(0, 0, 0), d=(1, 0, 0), t=3.4e38, src=7, st=0.05),                        # a miss: tMax = MAXFLOAT (vptShadeMethods.h:1287)
])
def test_oracle_mixture_density_is_normalised(l1, case):
    """importance-sampling identities of the construction: P(surface) = Tr; E[1{medium} g(s) / p(s)] = int_0^t g for any g -- checked with
    g = sigma_t exp(-sigma_t s) (integral 1 - Tr), g = equi-angular pdf (integral 1, times 1 - Tr in p), and the sample range [0, t)"""
    rng = np.random.default_rng(5)
    n = 400000
    d = np.array(case["d"], dtype=float); d /= np.linalg.norm(d)
    rows = _rows(rng, n, case["t"], case["src"], case["o"], d, case["st"])
    out = l1.mis_distance(DEFAULT_SCENE, rows)
    surf = out[:, 0] == 1
    Tr = np.exp(-case["st"] * case["t"])
    assert abs(surf.mean() - Tr) < 4 * np.sqrt(max(Tr * (1 - Tr), 1e-12) / n) + 1e-12
    assert np.array_equal(surf, rows[:, 10] < Tr)
    s, p = out[~surf, 1], out[~surf, 2]
    assert np.all(s >= 0) and np.all(s <= case["t"] * (1 + 1e-12)) and np.all(p > 0) and np.isfinite(p).all()
    # technique split: half of the medium samples each
    free = rows[~surf, 10] < 0.5 + 0.5 * Tr
    assert abs(free.mean() - 0.5) < 0.005
    g_free = case["st"] * np.exp(-case["st"] * s)
    est = np.where(surf, 0.0, 0.0); est[~surf] = g_free / p
    assert abs(est.mean() - (1 - Tr)) < 5 * est.std() / np.sqrt(n)
    g_equi = 2 * p - g_free                                         # = equiAngularProb(s) (1 - Tr)
    assert np.all(g_equi > -1e-15)
    est[~surf] = g_equi / p
    assert abs(est.mean() - (1 - Tr)) < 5 * est.std() / np.sqrt(n)
    # balance heuristic: the weight g / p of either technique never exceeds 2 (bounded weights are what removes both estimators' fireflies)
    assert (g_free / p).max() <= 2 + 1e-12 and (g_equi / p).max() <= 2 + 1e-12


def test_oracle_mis_distance_has_the_expectation_of_the_reference_methods(l1):
    """whole renders through the FP64 oracle (robust semantics): method 4's image mean equals free flight's and equi-angular's inside the
    Monte Carlo noise (z test on ~1.8e6 paths each), and its per-pixel variance is below both"""
    w, h, spp = 96, 72, 256
    res = {}
    for m in (0, 1, 4):
        tot, sq, st = l1.render(DEFAULT_SCENE, 0, m, SA, SS, w, h, 1, spp)
        mean = tot / spp
        res[m] = (mean.mean(axis=(0, 1)), (sq / spp - mean ** 2).mean(axis=(0, 1)), st)
        assert abs(st["events"] / st["paths"] - 1.5) < 0.01
    n = w * h * spp
    for other in (0, 1):
        z = (res[4][0] - res[other][0]) / np.sqrt((res[4][1] + res[other][1]) / n)
        assert np.all(np.abs(z) < 4), (other, z)
    assert np.all(res[4][1] < res[0][1]) and np.all(res[4][1] < res[1][1])


def test_oracle_mis_distance_wiring_against_the_reference_method(l1):
    """single-vertex paths on explicit draw lists (every draw after the decision is < 0.4, so the second roulette ends the path): with the
    decision draw below Tr both method 2 (the reference's MISVPTTracerRecursive, pinned on reference vectors) and method 4 shade the same
    surface vertex -- identical radiance, identical number of draws; with the decision draw in the equi-angular half of the medium range
    both place the SAME medium vertex and differ exactly by the density ratio equiAngularProb (1 - Tr) / mixture"""
    rng = np.random.default_rng(3)
    o = np.array([10., -5., 50.]); d = np.array([-.3, .2, -.9]); d /= np.linalg.norm(d)
    t, st = 141.77704970529885, SA + SS                       # the reference's hit distance for this ray
    Tr = np.exp(-st * t)
    n_surface = n_medium = 0
    for _ in range(300):
        u = rng.random(120) * 0.4
        u[0] = 0.4 + 0.6 * rng.random()                        # survive the first roulette
        u[1] = rng.random()
        u[2] = rng.random()
        u[3] = Tr * rng.random() if rng.random() < 0.5 else 0.5 + 0.5 * Tr + (0.5 - 0.5 * Tr) * rng.random()
        L4, used4 = l1.radiance_list(DEFAULT_SCENE, 0, 4, SA, SS, o, d, u)
        L2, used2 = l1.radiance_list(DEFAULT_SCENE, 0, 2, SA, SS, o, d, u)
        assert used4 == used2 and used4 > 4
        if u[3] < Tr:
            assert np.array_equal(L4, L2); n_surface += 1
        else:
            emitters = [7, 8, 9]
            src = emitters[int(u[1] * 3)]
            m = l1.mis_distance(DEFAULT_SCENE, np.array([[src, t, *o, *d, st, u[2], u[3]]]))[0]
            assert m[0] == 0
            ratio = (2 * m[2] - st * np.exp(-st * m[1])) / m[2]
            np.testing.assert_allclose(L4, L2 * ratio, rtol=1e-12, atol=0); n_medium += 1
    assert n_surface > 50 and n_medium > 50


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [0, 1])
def test_gpu_unit_mis_distance(gpu, l1, precision):
    """the device's distance decision against the oracle on identical inputs: FP64 to rounding, FP32 within 1e-5 relative (north_star's unit
    tolerance) wherever the discrete choices (surface / technique) agree -- they can only differ when xd falls within fp32 rounding of Tr or
    (1 + Tr) / 2"""
    rng = np.random.default_rng(17)
    n = 20000
    rows = np.zeros((n, 11))
    rows[:, 0] = rng.choice([7, 8, 9], n)
    rows[:, 2:5] = np.c_[rng.uniform(-45, 45, n), rng.uniform(-38, 38, n), rng.uniform(-75, 150, n)]
    v = rng.normal(size=(n, 3)); rows[:, 5:8] = v / np.linalg.norm(v, axis=1, keepdims=True)
    rows[:, 1] = rng.uniform(5, 300, n); rows[::7, 1] = 3.4e38                    # every seventh row: a miss (tMax = MAXFLOAT)
    rows[:, 8] = rng.choice([0.01, 0.05], n); rows[:, 9] = rng.random(n); rows[:, 10] = rng.random(n)
    if precision == 0:  # the kernel's inputs are fp32: round them first so that both sides see the same numbers
        rows[:, 1:] = rows[:, 1:].astype(np.float32).astype(np.float64)
    want = l1.mis_distance(DEFAULT_SCENE, rows)
    q = gpu.QUIRKS_NONE
    got = gpu.unit(gpu.UNIT.MIS_DISTANCE, rows, gpu.default_params(precision=precision, quirks=q))
    if precision == 1:
        assert np.array_equal(got[:, 0], want[:, 0])
        np.testing.assert_allclose(got[:, 1:], want[:, 1:], rtol=1e-10)
        return
    same = got[:, 0] == want[:, 0]
    assert same.mean() > 0.9995
    med = same & (want[:, 0] == 0)
    # distances along the ray are sums of the light's projection and a local offset: relative to the larger of the distance and the light's
    # distance from the ray origin (as tests/test_gpu_units.py::test_equiangular_sample_and_pdf); pdf at that test's 4e-5
    light = DEFAULT_SCENE[rows[:, 0].astype(int), 1:4]
    scale = np.maximum(np.abs(want[:, 1]), np.linalg.norm(light - rows[:, 2:5], axis=1))
    e_d = (np.abs(got[:, 1] - want[:, 1]) / scale)[med]
    e_p = np.abs(got[med, 2] - want[med, 2]) / want[med, 2]
    # a flipped technique choice (xd within an ulp of (1 + Tr) / 2) gives another, equally valid, sample: allow a handful
    assert np.mean(e_d < 1e-5) > 0.9995 and np.mean(e_p < 4e-5) > 0.9995, (np.mean(e_d < 1e-5), np.mean(e_p < 4e-5))
    assert np.median(e_d) < 5e-7 and np.median(e_p) < 5e-7
