"""Material 2 (dielectric) exactly as the reference writes it: bdsf vptShadeMethods.h:26-46, softDielectric
samplingFunctions.h:209-235, the dielectric branches of MISv2 misSamplingFunctions.h:110-118,144-152, refraxDielectric / reflexDielectric /
fresnelDie microFacetUtilities.h:107-141.  No scene of Sphere.cpp uses it; the goldens (tests/golden/dielectric.npz, tools/gen_golden.py
dielectric) come from the UNMODIFIED reference with the blue ball / both balls switched to material 2.  The reference's formulas are not
Snell's law (cos_t = sqrt(..) - 1, local x, y scaled by -1.5) and every refraction multiplies the throughput by 1.5^2: renders with such a
sphere have no finite expectation, so parity is per path, not statistical."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from test_oracle_pinning import check
from test_gpu_paths import erand48_stream, rel

SA, SS = 0.001, 0.009
SCENES = ["glass6", "glass56"]


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLDEN, "dielectric.npz")))


@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("quirks", [3, 0])
def test_oracle_matches_the_reference_with_dielectric_spheres(l1, gold, name, quirks):
    """the FP64 restatement against the unmodified reference: same number of erand48 draws, free flight bit for bit"""
    for method in (0, 1, 2):
        res = gold["%s_q%d_m%d" % (name, quirks, method)]
        assert np.mean(np.abs(res[:, :3]).max(axis=1) > 0) > 0.3
        check(l1, gold["rows_" + name], quirks, method, gold["o"], gold["d"], gold["seeds"], res)


def test_dielectric_scene_file_loads_through_the_abi(vpt, gold):
    """scenes/glass_ball.txt = the golden scene `glass6`; material 3 stays rejected, other values are malformed"""
    from conftest import ROOT
    rows = vpt.scene_to_rows(vpt.load_scene(os.path.join(ROOT, "scenes", "glass_ball.txt")))
    assert np.array_equal(rows, gold["rows_glass6"])


def _dielectric_rows(n, seed):
    rng = np.random.default_rng(seed)
    nn = rng.normal(size=(n, 3)); nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    wo = rng.normal(size=(n, 3)); wo /= np.linalg.norm(wo, axis=1, keepdims=True)
    return np.hstack([nn, wo])


def test_oracle_dielectric_functions_are_the_references(l1, l0):
    """refraxDielectric / reflexDielectric / fresnelDie of the restatement against the unmodified reference functions"""
    rows = _dielectric_rows(300, 1)
    got = l1.dielectric(rows)
    for i in range(len(rows)):
        n, wo = rows[i, :3], rows[i, 3:]
        wt = l0.refraxDielectric(1.0, 1.5, wo, n); wt = wt / np.sqrt(wt @ wt)
        wr = l0.reflexDielectric(wo, n); wr = wr / np.sqrt(wr @ wr)
        F = l0.fresnelDie(1.0, 1.5, float(n @ wt), float(n @ wo))
        np.testing.assert_allclose(got[i, :3], wt, rtol=0, atol=1e-15); np.testing.assert_allclose(got[i, 3:6], wr, rtol=0, atol=1e-15)
        assert abs(got[i, 6] - F) <= 1e-12 * max(abs(F), 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [0, 1])
def test_gpu_unit_dielectric(gpu, l1, precision):
    """unit level: the device's material-2 directions and Fresnel term against the oracle on identical inputs -- FP64 to rounding, FP32 within
    north_star's 1e-5 (directions: absolute on unit vectors; F: relative, away from the two poles of fresnelDie where the reference's own value
    runs through +-inf)"""
    rows = _dielectric_rows(20000, 2)
    if precision == 0:
        rows = rows.astype(np.float32).astype(np.float64)
        rows[:, :3] /= np.linalg.norm(rows[:, :3], axis=1, keepdims=True); rows[:, 3:] /= np.linalg.norm(rows[:, 3:], axis=1, keepdims=True)
        rows = rows.astype(np.float32).astype(np.float64)
    want = l1.dielectric(rows)
    got = gpu.unit(gpu.UNIT.DIELECTRIC, rows, gpu.default_params(precision=precision))
    if precision == 1:
        np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-13)
        return
    assert np.abs(got[:, :6] - want[:, :6]).max() < 1e-5
    ci = np.einsum("ij,ij->i", rows[:, :3], rows[:, 3:]); cn = np.einsum("ij,ij->i", rows[:, :3], want[:, :3])
    away = (np.abs(1.5 * ci + cn) > 0.05) & (np.abs(ci + 1.5 * cn) > 0.05)
    assert away.mean() > 0.85
    assert (np.abs(got[away, 6] - want[away, 6]) / np.abs(want[away, 6])).max() < 1e-5


def _list_rows(gold):
    n = len(gold["o"])
    rows = np.zeros((n, 127))
    rows[:, 0:3] = gold["o"]; rows[:, 3:6] = gold["d"]; rows[:, 6] = 120
    for i in range(n):
        rows[i, 7:] = erand48_stream(gold["seeds"][i], 120)
    return rows


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCENES)
def test_gpu_fp64_reproduces_the_reference_with_dielectric_spheres(gpu, gold, name):
    """REF mode on the reference's own erand48 sequences (robust semantics: nothing rounding-decided): the reference's numbers and draw counts"""
    rows = _list_rows(gold)
    scene = gpu.scene_from_rows(gold["rows_" + name])
    for method in (0, 1, 2):
        got = gpu.unit(gpu.UNIT.RADIANCE_LIST, rows, gpu.default_params(method=method, precision=gpu.PRECISION_FP64_REF, quirks=0), scene)
        want = gold["%s_q0_m%d" % (name, method)]
        ok = got[:, 3] >= 0
        assert ok.mean() > 0.9
        assert np.array_equal(got[ok, 3], want[ok, 3])
        assert rel(got[ok, :3], want[ok, :3]).max() < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCENES)
def test_gpu_fp32_paths_with_dielectric_spheres(gpu, gold, name):
    """the performance path on the same sequences.  fresnelDie has poles (1.5 cos_i + n.wt = 0 near cos_i = 0.11) where the reference's F
    runs through +-inf: paths that touch them, and paths whose decision flips in fp32, differ; the rest agree to fp32 rounding"""
    rows = _list_rows(gold)
    scene = gpu.scene_from_rows(gold["rows_" + name])
    for method in (0, 1, 2):
        got = gpu.unit(gpu.UNIT.RADIANCE_LIST, rows, gpu.default_params(method=method), scene)
        want = gold["%s_q0_m%d" % (name, method)]
        ok = got[:, 3] >= 0
        same = got[ok, 3] == want[ok, 3]
        assert same.mean() > 0.97, same.mean()
        e = rel(got[ok, :3], want[ok, :3])[same]
        assert np.median(e[e > 0]) < 5e-6 and np.mean(e > 1e-3) < 0.03, (np.median(e[e > 0]), np.mean(e > 1e-3))


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCENES)
def test_gpu_product_kernel_per_path_parity_with_dielectric_spheres(gpu, l1, gold, name):
    """spp = 1 renders: every pixel is one path of the product kernel (AUTO) on its Philox stream against the FP64 oracle; the HBM wavefront
    and the megakernel give the same image; the two superseded kernels refuse the scene"""
    rows = gold["rows_" + name]
    scene = gpu.scene_from_rows(rows)
    w, h = 192, 144
    for method in (0, 1):
        p = gpu.default_params(width=w, height=h, spp=1, method=method, seed=13, output=gpu.OUTPUT_SUM)
        img, st = gpu.render(p, scene, stats=True)
        ref, _, rst = l1.render(rows, 0, method, SA, SS, w, h, 13, 1, want_sumsq=False)
        assert st.paths == w * h
        assert abs(int(st.events) - int(rst["events"])) <= 1e-3 * rst["events"] + 2
        fin = np.isfinite(ref).all(axis=2) & np.isfinite(img).all(axis=2)
        err = (np.abs(img - ref).max(axis=2) / np.maximum(np.abs(ref).max(axis=2), 1e-4))[fin]
        assert np.median(err) < 1e-6 and np.mean(err < 1e-4) > 0.985, (float(np.median(err)), float(np.mean(err < 1e-4)))
        hbm = gpu.render(p.copy(kernel=gpu.KERNEL_WAVEFRONT_HBM), scene)
        assert np.array_equal(img, hbm)
        mega = gpu.render(p.copy(kernel=gpu.KERNEL_MEGA), scene)
        e2 = (np.abs(img - mega).max(axis=2) / np.maximum(np.abs(mega).max(axis=2), 1e-4))[fin]
        assert np.median(e2) < 1e-6 and np.mean(e2 < 1e-4) > 0.985
        for kern in (gpu.KERNEL_MEGA_SCAN, gpu.KERNEL_WAVEFRONT):
            with pytest.raises(gpu.VptError) as e:
                gpu.render(p.copy(kernel=kern), scene)
            assert e.value.status == -3
    # the dielectric ball shows in the image: pixels covered by it differ from the Lambert render
    lam = gpu.render(gpu.default_params(width=w, height=h, spp=16, method=0, seed=13))
    die = gpu.render(gpu.default_params(width=w, height=h, spp=16, method=0, seed=13), scene)
    assert np.mean(np.abs(lam - die).max(axis=2) > 1e-3) > 0.02
