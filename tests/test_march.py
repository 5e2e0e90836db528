"""rayMarching3 (rayMarchingMethods.h:330-384, the commented call src/rt.cpp:791): the deterministic ray-marching solver, SURVEY.md 8(f)-2.
CPU: the FP64 restatement against the unmodified reference's vectors (tests/golden/march.npz; live against oracle/_ref where it exists).
GPU: the unit kernel and whole frames of VPT_METHOD_RAYMARCH against the same vectors / the oracle, through the C-ABI."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle_lib import DEFAULT_SCENE

CASES = ["src7_step01", "src8_step01", "src8_step05", "src9_step1"]


@pytest.fixture(scope="module")
def march():
    return dict(np.load(os.path.join(GOLDEN, "march.npz")))


def rel(got, want):
    return np.abs(got - want).max(axis=1) / np.maximum(np.abs(want).max(axis=1), 1e-300)


# ---- CPU: oracle pinned on the reference's own outputs ----------------------------------------------------------------------
@pytest.mark.parametrize("quirks", [3, 0])
def test_oracle_matches_reference_vectors(l1, march, quirks):
    for k, name in enumerate(CASES):
        sa, ss, step, src = march["cases"][k]
        want = march["q%d_%s" % (quirks, name)]
        got = l1.ray_march3(DEFAULT_SCENE, quirks, sa, ss, step, int(src), march["o"], march["d"])
        assert np.array_equal(got[:, :3] == 0, want == 0)
        assert rel(got[:, :3], want)[(want != 0).any(axis=1)].max(initial=0.0) < 1e-13
        if int(src) != 8:  # an area source: its own sphere blocks the shadow ray that starts at its centre -> black (as in the reference)
            assert not want.any()


def test_oracle_matches_reference_live(l0, l1, march):
    rng = np.random.default_rng(5)
    l0.set_quirks(3)
    for i in rng.integers(0, len(march["o"]), 12):
        want = l0.rayMarching3(march["o"][i], march["d"][i], 0.001, 0.009, 0.25, 8)
        got = l1.ray_march3(DEFAULT_SCENE, 3, 0.001, 0.009, 0.25, 8, march["o"][i], march["d"][i])[0, :3]
        assert np.allclose(got, want, rtol=1e-13, atol=0)


def test_march_parameters_are_validated(vpt):
    import ctypes as C
    lib = vpt.load_library()
    host = np.zeros((4, 4, 3), dtype=np.float32)

    def rc(**kw):
        p = vpt.default_params(width=4, height=4, spp=1, method=vpt.METHOD_RAYMARCH, **kw)
        return lib.vpt_render(C.byref(p), vpt.default_scene(), 10, host.ctypes.data_as(C.POINTER(C.c_float)), None)
    d = vpt.default_params()
    assert (d.march_source, d.march_step) == (7, 0.1)          # the literals of rt.cpp:791
    assert rc(march_step=0.0) == -1 and rc(march_step=float("nan")) == -1 and rc(march_source=-1) == -1 and rc(march_source=40) == -1
    assert rc() in (-4, -5)                                     # valid arguments: only the missing device stops it here (no CPU fallback)


# ---- GPU ---------------------------------------------------------------------------------------------------------------------------
def unit_rows(march, step, src):
    n = len(march["o"])
    return np.hstack([march["o"], march["d"], np.full((n, 1), step), np.full((n, 1), src)])


@pytest.mark.gpu
@pytest.mark.parametrize("quirks", [3, 0])
def test_gpu_fp64_unit_matches_reference_vectors(gpu, march, quirks):
    for k, name in enumerate(CASES):
        sa, ss, step, src = march["cases"][k]
        p = gpu.default_params(precision=gpu.PRECISION_FP64_REF, quirks=quirks, sigma_a=sa, sigma_s=ss)
        got = gpu.unit(gpu.UNIT.RAYMARCH, unit_rows(march, step, src), p)
        want = march["q%d_%s" % (quirks, name)]
        assert np.array_equal(got[:, :3] == 0, want == 0)
        assert rel(got[:, :3], want)[(want != 0).any(axis=1)].max(initial=0.0) < 1e-11


@pytest.mark.gpu
def test_gpu_fp32_unit_matches_reference_vectors(gpu, l1, march):
    """FP32: 1e-5 relative (north_star's unit tolerance).  The loop bound t / step decides one step more or less (3e-4 of the sum) when
    fp32 rounding of t moves it across an integer: allowed for at most 1 % of the rays, and then nothing worse than one step."""
    for k, name in enumerate(CASES):
        sa, ss, step, src = march["cases"][k]
        p = gpu.default_params(sigma_a=sa, sigma_s=ss)
        got = gpu.unit(gpu.UNIT.RAYMARCH, unit_rows(march, step, src), p)
        want = march["q0_%s" % name]
        steps = l1.ray_march3(DEFAULT_SCENE, 0, sa, ss, step, int(src), march["o"], march["d"])[:, 3]
        assert np.array_equal(got[:, :3] == 0, want == 0)
        lit = (want != 0).any(axis=1)
        if lit.any():
            e = rel(got[:, :3], want)[lit]
            assert np.mean(e < 1e-5) >= 0.99 and e.max() < 3.0 / steps[lit].min()
        assert np.mean(got[:, 3] == steps) >= 0.99 and np.abs(got[:, 3] - steps).max() <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [0, 1])
def test_gpu_march_frames_match_oracle(gpu, l1, precision):
    """whole frames (rt.cpp:791 as the active line): same Philox jitter on both sides; tile / sample shards add up"""
    w, h, spp = 48, 36, 2
    kw = dict(width=w, height=h, spp=spp, method=gpu.METHOD_RAYMARCH, march_source=8, march_step=0.5, sigma_a=0.001, sigma_s=0.0125, seed=9,
              output=gpu.OUTPUT_SUM, precision=precision, quirks=3 if precision else 0)
    img, st = gpu.render(gpu.default_params(**kw), stats=True)
    ref = l1.render_march(DEFAULT_SCENE, 3 if precision else 0, 0.001, 0.0125, 0.5, 8, w, h, 9, spp)
    assert st.paths == w * h * spp and st.nonfinite == 0 and img.any()
    e = np.abs(img - ref) / np.maximum(np.abs(ref), 1e-12)
    if precision:
        assert e.max() < 1e-6   # fp64 sums, float32 output
    else:
        assert np.mean(e[..., 0] < 1e-5) >= 0.98 and e.max() < 2e-2
    a = gpu.render(gpu.default_params(**kw, sample_begin=0, sample_end=1)) + gpu.render(gpu.default_params(**kw, sample_begin=1, sample_end=2))
    assert np.allclose(a, img, rtol=1e-6, atol=0)
    t = sum(gpu.render(gpu.default_params(**kw, tile_rank=r, tile_count=3)) for r in range(3))
    assert np.array_equal(t, img)
    black = gpu.render(gpu.default_params(**{**kw, "march_source": 7}))  # rt.cpp:791 passes 7: an area light hides behind its own sphere
    assert not black.any()
