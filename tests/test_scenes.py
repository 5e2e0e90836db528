"""The reference's five commented alternate scenes (include/Sphere.cpp:27-106) as scene files (scenes/*.txt): loaded through
vpt_load_scene instead of recompiling.  CPU: the loader; the FP64 restatement against the UNMODIFIED reference on these scenes
(tests/golden/scenes.npz, 120 seeded erand48 paths per scene / method / quirk setting).  GPU: the product kernel, one path per pixel,
against the oracle on every scene."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle_lib import DEFAULT_SCENE
from test_oracle_pinning import check

SA, SS = 0.001, 0.009
ALT = ["scene2_sigma", "scene3_near_camera", "scene4_area_light", "scene5_infinite", "scene6_two_points"]


def path_of(name):
    return os.path.join(ROOT, "scenes", name + ".txt")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLDEN, "scenes.npz")))


def test_scene_files_load_through_the_abi(vpt, gold, tmp_path):
    assert np.array_equal(vpt.scene_to_rows(vpt.load_scene(path_of("default"))), DEFAULT_SCENE)
    assert np.array_equal(vpt.scene_to_rows(vpt.load_scene(path_of("default"))), vpt.scene_to_rows(vpt.default_scene()))
    for name in ALT:
        assert np.array_equal(vpt.scene_to_rows(vpt.load_scene(path_of(name))), gold["rows_" + name])
    # commas, comments, blank lines; errors: missing file, short line, too many spheres, fractional material
    f = tmp_path / "s.txt"
    f.write_text("# c\n\n1, 0,0,0, .5,.5,.5, 0,0,0, 0, 0,0,0, 0,0,0, 0 # tail\n0 1 2 3 0 0 0 5 5 5 0 0 0 0 0 0 0 0\n")
    rows = vpt.scene_to_rows(vpt.load_scene(str(f)))
    assert rows.shape == (2, 18) and rows[1, 0] == 0 and rows[1, 7] == 5 and rows[0, 4] == 0.5
    for bad in ("1 2 3\n", "1 0 0 0 0 0 0 0 0 0 0.5 0 0 0 0 0 0 0\n", "x\n", "".join("1 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0\n" for _ in range(33)), "# nothing\n"):
        f.write_text(bad)
        with pytest.raises(vpt.VptError) as e:
            vpt.load_scene(str(f))
        assert e.value.status == -2
    with pytest.raises(vpt.VptError) as e:
        vpt.load_scene(str(tmp_path / "missing.txt"))
    assert e.value.status == -6


@pytest.mark.parametrize("name", ALT)
@pytest.mark.parametrize("quirks", [3, 0])
def test_oracle_matches_the_reference_on_the_alternate_scenes(l1, gold, name, quirks):
    for method in (0, 1, 2):
        check(l1, gold["rows_" + name], quirks, method, gold["o"], gold["d"], gold["seeds"], gold["%s_q%d_m%d" % (name, quirks, method)])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ALT)
def test_gpu_per_path_parity_on_the_alternate_scenes(gpu, l1, name):
    """spp = 1: every pixel is one path of the product kernel; FP64 oracle on the same Philox streams.  The glossy scenes (Beckmann alpha 0.02-0.03
    walls and spheres) amplify fp32 rounding of the half vector, hence 1e-4 there for the bulk and the median as the tight statement."""
    scene = gpu.load_scene(path_of(name))
    rows = gpu.scene_to_rows(scene)
    w, h = 160, 120
    for method in (0, 1):
        p = gpu.default_params(width=w, height=h, spp=1, method=method, seed=5, output=gpu.OUTPUT_SUM)
        img, st = gpu.render(p, scene, stats=True)
        ref, _, rst = l1.render(rows, 0, method, SA, SS, w, h, 5, 1, want_sumsq=False)
        assert st.paths == w * h and st.nonfinite == 0
        assert abs(int(st.events) - int(rst["events"])) <= 5e-4 * rst["events"] + 2
        err = np.abs(img - ref).max(axis=2) / np.maximum(np.abs(ref).max(axis=2), 1e-4)
        assert np.median(err) < 2e-6, float(np.median(err))
        assert np.mean(err < 1e-4) > 0.97, float(np.mean(err < 1e-4))
        np.testing.assert_allclose(img.sum(axis=(0, 1)), ref.sum(axis=(0, 1)), rtol=0.03)
