"""CPU: the C-ABI library loads, exports every symbol include/vpt.h declares, its structs have the layout the Python mirror
assumes, argument validation answers before any device work, and compute entry points FAIL LOUDLY without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "vpt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vpt_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(vpt):
    lib = vpt.load_library()
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "include/vpt.h declares %s but libvpt_b200.so does not export it" % n


def test_struct_layout_matches_header(vpt):
    assert C.sizeof(vpt.Sphere) == 8 * 17 + 8  # 17 doubles + material + pad
    assert C.sizeof(vpt.Params) == 9 * 4 + 4 + 10 * 8 + 8 + 6 * 4 or C.sizeof(vpt.Params) % 8 == 0
    p = vpt.default_params()
    # the reference's literals: rt.cpp:752,755-759,794 and vptShadeMethods.h:1275
    assert (p.width, p.height, p.method, p.max_depth) == (1024, 768, 0, 0)
    assert (p.sigma_a, p.sigma_s, p.continue_prob, p.fov) == (0.001, 0.009, 0.6, 0.5095)
    assert list(p.cam_o) == [0, 11.2, 214] and list(p.cam_dir) == [0, -0.042612, -1]
    assert (p.precision, p.quirks, p.output) == (vpt.PRECISION_FP32, 0, vpt.OUTPUT_MEAN)


def test_default_scene_is_the_references(vpt):
    from oracle_lib import DEFAULT_SCENE
    rows = vpt.scene_to_rows(vpt.default_scene())
    assert np.array_equal(rows, DEFAULT_SCENE)  # include/Sphere.cpp:11-22
    assert np.array_equal(vpt.scene_to_rows(vpt.scene_from_rows(DEFAULT_SCENE)), DEFAULT_SCENE)


def test_unit_strides_table(vpt):
    for fn in range(22):
        si, so = vpt.unit_strides(fn)
        assert si > 0 and so > 0
    with pytest.raises(vpt.VptError):
        vpt.unit_strides(99)


def _rc(vpt, params, scene=None, hdr=True):
    lib = vpt.load_library()
    scene = scene if scene is not None else vpt.default_scene()
    buf = np.zeros((max(params.height, 1), max(params.width, 1), 3), dtype=np.float32)
    return lib.vpt_render(C.byref(params), scene, len(scene), buf.ctypes.data_as(C.POINTER(C.c_float)) if hdr else None, None)


def test_argument_validation_precedes_device_work(vpt):
    from oracle_lib import DEFAULT_SCENE
    P = lambda **kw: vpt.default_params(**{**dict(width=32, height=16, spp=2), **kw})
    INVALID, SCENE, UNSUPPORTED = -1, -2, -3
    assert _rc(vpt, P(spp=0)) == INVALID
    assert _rc(vpt, P(width=0)) == INVALID
    assert _rc(vpt, P(method=6)) == INVALID
    assert _rc(vpt, P(method=5)) == UNSUPPORTED                              # VPT_METHOD_VOLUME_SPHERES: FP64_REF precision only
    assert _rc(vpt, P(sample_begin=1, sample_end=1)) == INVALID
    assert _rc(vpt, P(sample_begin=0, sample_end=3)) == INVALID
    assert _rc(vpt, P(tile_rank=2, tile_count=2)) == INVALID
    assert _rc(vpt, P(continue_prob=0.0)) == INVALID
    # a roulette that never fires + no depth limit = paths that never end in a scene without emitter geometry (scenes/scene2,3,5,6): a GPU hang.
    # Unlimited depth needs continue_prob <= 0.99; above that an explicit max_depth; every kernel caps paths at VPT_MAX_DEPTH = 4095 bounces.
    assert _rc(vpt, P(continue_prob=1.0)) == INVALID
    assert _rc(vpt, P(continue_prob=0.995, max_depth=0)) == INVALID
    assert _rc(vpt, P(continue_prob=1.0, max_depth=4096)) == INVALID
    assert _rc(vpt, P(max_depth=1 << 22)) == INVALID
    assert _rc(vpt, P(continue_prob=1.0, max_depth=64)) in (0, -4, -5)        # valid: refused only for lack of a device
    assert _rc(vpt, P(continue_prob=0.99)) in (0, -4, -5)
    assert _rc(vpt, P(spp=1 << 24)) == INVALID                               # samples per call: 32-bit per-item counters
    assert _rc(vpt, P(spp=1 << 25, sample_begin=5, sample_end=5 + (1 << 24))) == INVALID
    assert _rc(vpt, P(spp=1 << 25, sample_begin=5, sample_end=9)) in (0, -4, -5)
    assert _rc(vpt, P(sigma_a=0.0, sigma_s=0.0)) == INVALID
    assert _rc(vpt, P(quirks=8)) == INVALID
    assert _rc(vpt, P(), hdr=False) == INVALID
    assert _rc(vpt, P(quirks=vpt.QUIRKS_REFERENCE)) == UNSUPPORTED          # rounding-decided behaviours are FP64-only
    assert _rc(vpt, P(kernel=9)) == INVALID
    assert _rc(vpt, P(kernel=vpt.KERNEL_WAVEFRONT_HBM, precision=vpt.PRECISION_FP64_REF)) == UNSUPPORTED   # the multi-kernel wavefront is FP32 only
    assert _rc(vpt, P(kernel=vpt.KERNEL_WAVEFRONT_SM, precision=vpt.PRECISION_FP64_REF)) in (0, -4, -5)   # FP64 reference mode has its own SM-wide wavefront
    for gone in (vpt.KERNEL_WAVEFRONT, vpt.KERNEL_MEGA_SCAN):              # superseded variants: no longer built
        assert _rc(vpt, P(kernel=gone)) == UNSUPPORTED
    rows = DEFAULT_SCENE.copy(); rows[6, 10] = 3                             # volumetric sphere: undefined in the reference's active methods
    assert _rc(vpt, P(), vpt.scene_from_rows(rows)) == UNSUPPORTED
    rows[6, 10] = 4
    assert _rc(vpt, P(), vpt.scene_from_rows(rows)) == SCENE
    rows = DEFAULT_SCENE.copy(); rows[6, 0] = -1
    assert _rc(vpt, P(), vpt.scene_from_rows(rows)) == SCENE
    many = np.tile(DEFAULT_SCENE[7], (17, 1))                                # 17 emitters > VPT_MAX_EMITTERS (reference: UB beyond 4)
    assert _rc(vpt, P(), vpt.scene_from_rows(many)) == SCENE
    assert _rc(vpt, P(), vpt.scene_from_rows(np.tile(DEFAULT_SCENE[0], (33, 1)))) == SCENE
    assert vpt.load_library().vpt_strerror(-4).decode().startswith("no usable CUDA device")


def test_no_cpu_fallback(vpt):
    """Without a CUDA device every compute entry point must refuse; with one this test has nothing to say."""
    if vpt.device_count() > 0:
        pytest.skip("a CUDA device is present")
    p = vpt.default_params(width=16, height=16, spp=1)
    with pytest.raises(vpt.VptError) as e:
        vpt.render(p)
    assert e.value.status in (-4, -5)
    with pytest.raises(vpt.VptError):
        vpt.unit(vpt.UNIT.POWER_HEURISTIC, [[0.3, 0.7]])
    with pytest.raises(vpt.VptError):
        vpt.philox([[0, 0, 0, 0]], [[0, 0]])
    with pytest.raises(vpt.VptError):
        vpt.measure_fp32_peak()


def test_product_never_touches_the_oracle():
    """the package and the C sources must not import, include or link anything under oracle/"""
    pkg = os.path.join(ROOT, "minimal_volumetric_path_tracer_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_obj" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in text and "libvpt_oracle" not in text and "libvpt_l0" not in text, f
                assert not re.search(r'#include\s+"[^"]*oracle', text), f
