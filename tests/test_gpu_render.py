"""GPU: whole renders through the C-ABI.
  - common random numbers against the FP64 CPU oracle at sizes the oracle finishes in seconds,
  - z-score comparison of 16x16-block means against renders of the UNMODIFIED reference (tests/golden/image_*.npz,
    1024x768; SURVEY.md section 0 fact 10: a fixed 1 % tolerance would sit below the Monte Carlo noise floor),
  - size-independent properties at BASELINE.json's full sizes (determinism, shard additivity, tile union, seeds),
  - edge cases."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle_lib import DEFAULT_SCENE, scene_without

pytestmark = pytest.mark.gpu
SA, SS = 0.001, 0.009


# ---- common random numbers vs the oracle ------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", [0, 1, 2, 4])
def test_small_render_crn_fp64(gpu, l1, method):
    w, h, spp = 96, 72, 8
    p = gpu.default_params(width=w, height=h, spp=spp, method=method, precision=gpu.PRECISION_FP64_REF, quirks=0, seed=3, output=gpu.OUTPUT_SUM)
    hdr, st = gpu.render(p, stats=True)
    ref, _, rst = l1.render(DEFAULT_SCENE, 0, method, SA, SS, w, h, 3, spp, want_sumsq=False)
    assert st.events == rst["events"] and st.paths == w * h * spp and st.nonfinite == 0
    np.testing.assert_allclose(hdr, ref, rtol=2e-6, atol=1e-7)  # output buffer is fp32


@pytest.mark.parametrize("method", [0, 1, 2, 4])
def test_small_render_crn_fp32(gpu, l1, method):
    w, h, spp = 128, 96, 16
    p = gpu.default_params(width=w, height=h, spp=spp, method=method, seed=3, output=gpu.OUTPUT_SUM)
    hdr, st = gpu.render(p, stats=True)
    ref, _, rst = l1.render(DEFAULT_SCENE, 0, method, SA, SS, w, h, 3, spp, want_sumsq=False)
    assert abs(int(st.events) - rst["events"]) <= 2e-4 * rst["events"] and st.nonfinite == 0
    err = np.abs(hdr - ref) / np.maximum(np.abs(ref), 1e-3)
    assert np.median(err) < 2e-6 and np.mean(err > 1e-3) < 0.02        # a pixel differs only if one of its paths flipped a decision
    np.testing.assert_allclose(hdr.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=2e-3)


def test_dense_medium_long_paths_crn(gpu, l1):
    """BASELINE.json config 4: albedo 0.99, mean free path 20, continue_prob 0.95, max depth 64"""
    w, h, spp = 64, 48, 4
    kw = dict(sigma_a=0.0005, sigma_s=0.0495, continue_prob=0.95, max_depth=64)
    for method in (0, 2):
        p = gpu.default_params(width=w, height=h, spp=spp, method=method, precision=gpu.PRECISION_FP64_REF, seed=8, output=gpu.OUTPUT_SUM, **kw)
        hdr, st = gpu.render(p, stats=True)
        ref, _, rst = l1.render(DEFAULT_SCENE, 0, method, kw["sigma_a"], kw["sigma_s"], w, h, 8, spp, cp=0.95, max_depth=64, want_sumsq=False)
        assert st.events == rst["events"] and st.events / st.paths > 10
        np.testing.assert_allclose(hdr, ref, rtol=1e-5, atol=1e-7)
        p32 = p.copy(precision=gpu.PRECISION_FP32)
        hdr32 = gpu.render(p32)
        np.testing.assert_allclose(hdr32.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.03)


# ---- statistical comparison with the unmodified reference ---------------------------------------------------------------
def block_means(hdr, block=16):
    h, w, _ = hdr.shape
    return hdr[:h // block * block, :w // block * block].reshape(h // block, block, w // block, block, 3).mean(axis=(1, 3))


def z_scores(gpu, golden_name, precision, quirks, spp, scene_rows=None, seed=77, batches=16, stand_ins=6, main_method=None):
    """Is the reference render one more draw from the distribution the GPU samples?

    Block means at the reference's 256-512 spp are right-skewed (the estimator has fireflies: throughput grows 1.5x per
    medium bounce, 1/d^2 near the point light), so a normalised difference is NOT a unit normal even when both sides sample the
    same distribution.  The test therefore compares the reference with an ensemble of GPU stand-ins (other seeds, the
    reference's spp) through identical statistics:
      z        (main - X) / sigma per block and channel, main = a high-spp GPU render, sigma from a third independent
               render (batch-to-batch variance: correlated with neither side); summarised by clipped mean, median,
               median |z| and the 3-sigma tail fraction, for X = reference and X = each stand-in;
      sign     fraction of blocks where a stand-in exceeds the reference: 1/2 under H0 by symmetry, no variance model;
      glob_z   whole-image mean (a sum of ~1e8 samples: a clean normal test)."""
    g = np.load(os.path.join(GOLDEN, "image_%s.npz" % golden_name))
    w, h, method, spp_ref = int(g["width"]), int(g["height"]), int(g["method"]), int(g["spp"])
    p = gpu.default_params(width=w, height=h, spp=spp, method=method, precision=precision, quirks=quirks, seed=seed)
    scene = gpu.scene_from_rows(scene_rows) if scene_rows is not None else None
    # main_method: the high-spp render comes from ANOTHER estimator that claims the same expectation; noise model and stand-ins stay the
    # reference method's (the high-spp render's own variance enters sigma only through the small 1 / spp term)
    main, st = gpu.render(p if main_method is None else p.copy(method=main_method), scene, stats=True)
    assert st.nonfinite <= 1e-8 * st.paths  # dropped NaN/Inf paths (0 * inf in a BRDF at an exactly grazing fp32 direction): counted, < 1 in 1e8
    main = block_means(main.astype(np.float64))
    per = max(spp // (4 * batches), 4)
    q = p.copy(spp=per * batches, seed=seed + 2, output=gpu.OUTPUT_SUM)
    parts = np.stack([block_means(gpu.render(q.copy(sample_begin=b * per, sample_end=(b + 1) * per), scene).astype(np.float64)) / per for b in range(batches)])
    sigma = np.sqrt(np.maximum(parts.var(axis=0, ddof=1) * per * (1.0 / spp_ref + 1.0 / spp), 1e-300))
    ref = g["block_mean"].astype(np.float64)
    nulls, signs = [], []
    for k in range(stand_ins):
        s_in = block_means(gpu.render(p.copy(spp=spp_ref, seed=seed + 100 + k), scene).astype(np.float64))
        nulls.append(summary(((main - s_in) / sigma).ravel()))
        signs.append(np.mean(s_in > ref))
    glob = parts.mean(axis=(1, 2))
    glob_sigma = np.sqrt(glob.var(axis=0, ddof=1) * per * (1.0 / spp_ref + 1.0 / spp))
    return dict(ref=summary(((main - ref) / sigma).ravel()), nulls=np.array(nulls), sign=float(np.mean(signs)),
                glob_z=(main.mean(axis=(0, 1)) - ref.mean(axis=(0, 1))) / glob_sigma)


def summary(z):
    return np.array([np.mean(np.clip(z, -6, 6)), np.median(z), np.median(np.abs(z)), np.mean(np.abs(z) > 3)])


def check_statistically_equal(r):
    mu, sd = r["nulls"].mean(axis=0), r["nulls"].std(axis=0, ddof=1)
    tol = np.maximum(5 * sd, [0.06, 0.06, 0.04, 0.015])  # floors: sd from 6 stand-ins is itself coarse; a 1 % block bias is a 0.2 shift
    assert np.all(np.abs(r["ref"] - mu) < tol), "reference %s outside the stand-in ensemble %s +- %s" % (r["ref"], mu, tol)
    assert abs(r["sign"] - 0.5) < 0.02, "sign test: %.4f" % r["sign"]
    assert np.all(np.abs(r["glob_z"]) < 4.5), "whole-image mean off by z = %s" % r["glob_z"]


@pytest.mark.parametrize("method", [0, 1, 2])
def test_ref_mode_matches_the_as_shipped_reference(gpu, method):
    """FP64 REF mode with both quirks = the reference exactly as it ships, point light and rounding-decided branches included"""
    check_statistically_equal(z_scores(gpu, "strict_m%d" % method, gpu.PRECISION_FP64_REF, gpu.QUIRKS_REFERENCE, spp=512))


@pytest.mark.parametrize("method", [0, 1, 2])
def test_fp32_matches_reference_with_robust_hooks(gpu, method):
    """FP32 = the reference with its two rounding-decided behaviours replaced by the well-defined alternative"""
    check_statistically_equal(z_scores(gpu, "robust_m%d" % method, gpu.PRECISION_FP32, 0, spp=4096))


@pytest.mark.parametrize("method", [0, 1, 2])
def test_fp32_matches_unmodified_reference_without_point_light(gpu, method):
    """without the r = 0 sphere the UNMODIFIED reference has no rounding-decided branch: direct comparison, no hooks"""
    check_statistically_equal(z_scores(gpu, "no8_m%d" % method, gpu.PRECISION_FP32, 0, spp=4096, scene_rows=scene_without([8])))


@pytest.mark.parametrize("golden", ["robust_m0", "robust_m1", "no8_m2"])
def test_mis_distance_matches_the_reference_renders(gpu, golden):
    """VPT_METHOD_MIS_DISTANCE (not in the reference) has the expectation of the reference's methods: its renders must be
    statistically indistinguishable from the reference's free-flight and equi-angular renders (robust hooks) and from the unmodified
    reference on the scene without the point light.  The 4096-spp render is method 4's; the stand-in ensemble and the noise model are the
    reference method's own (GPU renders of that method at the reference's spp, shown equal to the reference by the tests above)."""
    check_statistically_equal(z_scores(gpu, golden, gpu.PRECISION_FP32, 0, spp=4096, main_method=gpu.METHOD_MIS_DISTANCE,
                                       scene_rows=scene_without([8]) if golden.startswith("no8") else None))


def test_mis_distance_lowers_the_variance(gpu):
    """the balance heuristic's point: per-pixel variance (from 32 independent 8-spp renders) summed over the image is below that of
    BOTH single-technique estimators, in every channel (FP64 oracle at 96x72: 2.90 / 0.82 / 0.39 against free flight 5.77 / 1.05 / 0.44 and
    equi-angular 3.39 / 0.90 / 0.50)"""
    def total_var(method):
        p = gpu.default_params(width=256, height=192, spp=256, method=method, seed=21, output=gpu.OUTPUT_SUM)
        parts = np.stack([gpu.render(p.copy(sample_begin=8 * b, sample_end=8 * b + 8)).astype(np.float64) / 8 for b in range(32)])
        return parts.var(axis=0, ddof=1).mean(axis=(0, 1)), parts.mean(axis=(0, 1, 2))
    (v0, m0), (v1, m1), (v4, m4) = total_var(0), total_var(1), total_var(gpu.METHOD_MIS_DISTANCE)
    assert np.all(v4 < 0.97 * np.minimum(v0, v1)), (v0, v1, v4)
    np.testing.assert_allclose(m4, m0, rtol=0.03); np.testing.assert_allclose(m4, m1, rtol=0.03)


def test_fp32_matches_the_reference_render_at_4096_spp(gpu):
    """north_star's correctness statement, literally: the reference's CPU image at 4096 spp ("MIS" method, tests/golden/image_robust_m2_4096.npz:
    11 minutes of the unmodified reference code with the robust hooks) against a GPU render.  "Per-16x16-block mean radiance within 1 %" sits AT
    the Monte Carlo noise floor of a 4096-spp render (1.0-1.3 % for the median block between two correct renders;
    measured here: 1.08 % against a 16384-spp GPU render), so the 1 % statement is made relative to what noise alone does: the reference is
    as close to the GPU render as ANOTHER GPU render with the reference's 4096 spp is -- same median block difference, same fraction of blocks
    within 1 %.  Plus: image RMSE within 3 sigma of the estimated variance, and the calibrated ensemble test."""
    r = z_scores(gpu, "robust_m2_4096", gpu.PRECISION_FP32, 0, spp=16384)
    check_statistically_equal(r)
    g = np.load(os.path.join(GOLDEN, "image_robust_m2_4096.npz"))
    ref = g["block_mean"].astype(np.float64)
    img = block_means(gpu.render(gpu.default_params(spp=16384, method=2, seed=4321)).astype(np.float64))
    twin = block_means(gpu.render(gpu.default_params(spp=4096, method=2, seed=999)).astype(np.float64))   # a correct 4096-spp render, by construction
    rel = np.abs(img - ref) / np.maximum(ref, 1e-9)
    rel_twin = np.abs(img - twin) / np.maximum(twin, 1e-9)
    assert np.median(rel) < 0.0125 and np.median(rel) < 1.12 * np.median(rel_twin), (float(np.median(rel)), float(np.median(rel_twin)))
    assert abs(np.mean(rel < 0.01) - np.mean(rel_twin < 0.01)) < 0.03, (float(np.mean(rel < 0.01)), float(np.mean(rel_twin < 0.01)))
    var = g["block_var"].astype(np.float64) * (1.0 + 4096.0 / 16384.0)          # variance of the difference of the two block means
    rmse, sigma = np.sqrt(np.mean((img - ref) ** 2, axis=(0, 1))), np.sqrt(np.mean(var, axis=(0, 1)))
    assert np.all(rmse < 3 * sigma), (rmse, sigma)
    assert np.all(rmse > 0.3 * sigma)                                           # ... and not suspiciously small either (the estimate means something)


def test_c3_matches_the_reference_render(gpu):
    """BASELINE.json config 3 as written -- 1920x1080, 4096 spp, the reference's "MIS" method -- rendered by the unmodified reference code (robust
    hooks; tests/golden/image_robust_m2_c3.npz, half an hour of 8 CPU threads) against the GPU: calibrated ensemble test on the 67 x 120 blocks"""
    if not os.path.exists(os.path.join(GOLDEN, "image_robust_m2_c3.npz")):
        pytest.skip("tests/golden/image_robust_m2_c3.npz not generated (python tools/gen_golden.py images robust_m2_c3: 30 minutes of CPU)")
    check_statistically_equal(z_scores(gpu, "robust_m2_c3", gpu.PRECISION_FP32, 0, spp=16384, stand_ins=4))


def test_fp64_robust_matches_reference_with_robust_hooks(gpu):
    check_statistically_equal(z_scores(gpu, "robust_m0", gpu.PRECISION_FP64_REF, 0, spp=512))


def test_the_statistical_test_has_power(gpu):
    """the same machinery must REJECT renders that are wrong by a few per cent: fog 5 % denser; FP32 semantics against the
    as-shipped reference (whose point light is attenuated by the rounding-decided branches)"""
    g = np.load(os.path.join(GOLDEN, "image_robust_m0.npz"))
    p = gpu.default_params(width=1024, height=768, spp=1024, method=0, seed=5, sigma_s=0.009 * 1.05)
    m = block_means(gpu.render(p).astype(np.float64))
    var = g["block_var"].astype(np.float64) * (1.0 + float(g["spp"]) / 1024)
    glob_sigma = np.sqrt(var.sum(axis=(0, 1))) / (var.shape[0] * var.shape[1])
    glob_z = (m.mean(axis=(0, 1)) - g["block_mean"].astype(np.float64).mean(axis=(0, 1))) / glob_sigma
    assert np.max(np.abs(glob_z)) > 6
    r = z_scores(gpu, "strict_m0", gpu.PRECISION_FP32, 0, spp=1024)
    assert abs(r["glob_z"][0]) > 20 and np.all(np.abs(r["glob_z"][1:]) < 4.5)   # red (point light) differs, green and blue do not
    with pytest.raises(AssertionError):
        check_statistically_equal(r)


def test_noise_floor_self_check(gpu):
    """two seeds of the same render differ by Monte Carlo noise and nothing else"""
    p1 = gpu.default_params(spp=256, method=1, seed=1); p2 = p1.copy(seed=2)
    a = block_means(gpu.render(p1).astype(np.float64)); b = block_means(gpu.render(p2).astype(np.float64))
    r = np.abs(a - b) / np.maximum(0.5 * (a + b), 1e-6)
    assert 0.01 < np.median(r) < 0.12 and abs(a.mean() - b.mean()) / a.mean() < 0.01


# ---- size-independent properties at full size ------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", [0, 1])
def test_determinism_and_sample_shards(gpu, precision):
    q = gpu.QUIRKS_REFERENCE if precision else 0
    p = gpu.default_params(width=1024, height=768, spp=24, method=2, precision=precision, quirks=q, seed=9, output=gpu.OUTPUT_SUM)
    whole, st = gpu.render(p, stats=True)
    again = gpu.render(p)
    assert np.array_equal(whole, again)                                      # Philox keyed (pixel, sample, bounce): reruns are bit-identical
    a, sa = gpu.render(p.copy(sample_begin=0, sample_end=10), stats=True)
    b, sb = gpu.render(p.copy(sample_begin=10, sample_end=24), stats=True)
    assert sa.events + sb.events == st.events and sa.paths + sb.paths == st.paths
    np.testing.assert_allclose(a.astype(np.float64) + b, whole, rtol=3e-7, atol=1e-7)  # same samples, only fp32 store rounding differs
    mean = gpu.render(p.copy(output=gpu.OUTPUT_MEAN))
    np.testing.assert_allclose(mean, whole / 24, rtol=3e-7)


def test_c3_full_size_properties(gpu):
    """BASELINE.json config 3 at its full size (1920x1080, 4096 spp, the reference's "MIS" method; 8.5e9 paths): the two halves of the sample
    range add up to the whole frame (same samples, fp32 store rounding only), a rerun is bit-identical, and the whole-image mean agrees with
    the unmodified reference's render of the same scene and method (tests/golden/image_robust_m2.npz is 1024x768: same camera, same field of
    view per unit height, so the image means are comparable to the noise of the reference render)"""
    p = gpu.default_params(width=1920, height=1080, spp=4096, method=2, seed=3, output=gpu.OUTPUT_SUM)
    whole, st = gpu.render(p, stats=True)
    assert st.paths == 1920 * 1080 * 4096 and st.nonfinite <= 1e-8 * st.paths and abs(st.events / st.paths - 1.5) < 1e-3
    a = gpu.render(p.copy(sample_begin=0, sample_end=2048)); b = gpu.render(p.copy(sample_begin=2048, sample_end=4096))
    np.testing.assert_allclose(a.astype(np.float64) + b, whole, rtol=3e-7, atol=1e-6)
    assert np.array_equal(whole, gpu.render(p))
    g = np.load(os.path.join(GOLDEN, "image_robust_m2.npz"))
    ref_mean = g["block_mean"].astype(np.float64).mean(axis=(0, 1))
    ref_sigma = np.sqrt(g["block_var"].astype(np.float64).sum(axis=(0, 1))) / (g["block_var"].shape[0] * g["block_var"].shape[1])
    # the 16:9 frame sees more of the side walls than the 4:3 one: compare the central 4:3 part (1440 of the 1920 columns)
    centre = (whole[:, 240:1680].astype(np.float64) / 4096).mean(axis=(0, 1))
    assert np.all(np.abs(centre - ref_mean) < 5 * ref_sigma + 0.002 * ref_mean), (centre, ref_mean, ref_sigma)


def test_item_slot_instantiations_agree(gpu):
    """The product kernel keeps six work items in flight per SM below 96 samples per pixel, four below 384 and two from there on (three
    instantiations of the scheduler, vpt_smsched.cuh).  Which one runs must not show in the image: a 512-spp frame in one launch (two slots)
    equals the sum of its [0,256) and [256,512) sample ranges (four slots each) and the sum of its eight 64-sample ranges (six slots each) up
    to the fp32 store, and every one of these launches equals the multi-kernel HBM wavefront -- which has no items at all -- bit for bit;
    ragged frame (not a multiple of the 128-pixel tile), all four FP32 shade methods"""
    for method in (0, 1, 2, 4):
        p = gpu.default_params(width=250, height=131, spp=512, method=method, seed=11, output=gpu.OUTPUT_SUM)
        whole, st = gpu.render(p, stats=True)
        assert st.paths == 250 * 131 * 512
        assert np.array_equal(whole, gpu.render(p.copy(kernel=gpu.KERNEL_WAVEFRONT_HBM)))
        for n_parts in (2, 8):
            per = 512 // n_parts
            total = np.zeros(whole.shape, dtype=np.float64)
            for k in range(n_parts):
                q = p.copy(sample_begin=k * per, sample_end=(k + 1) * per)
                part = gpu.render(q)
                assert np.array_equal(part, gpu.render(q.copy(kernel=gpu.KERNEL_WAVEFRONT_HBM))), (method, n_parts, k)
                total += part
            np.testing.assert_allclose(total, whole, rtol=6e-7, atol=1e-6)


def test_c5_frame_size_tiles_and_determinism(gpu):
    """BASELINE.json config 5's frame (3840x2160 = 64800 tiles of 128 pixels) at a reduced sample count: eight interleaved tile shards -- the
    8-GPU partition -- reassemble the one-GPU frame bit for bit, and sample shards add up"""
    p = gpu.default_params(width=3840, height=2160, spp=16, method=2, seed=5, output=gpu.OUTPUT_SUM)
    whole, st = gpu.render(p, stats=True)
    assert st.paths == 3840 * 2160 * 16 and whole.shape == (2160, 3840, 3)
    acc = np.zeros_like(whole)
    for r in range(8):
        part = gpu.render(p.copy(tile_rank=r, tile_count=8))
        assert not (acc != 0).any(axis=-1)[(part != 0).any(axis=-1)].any()   # disjoint
        acc += part
    assert np.array_equal(acc, whole)
    halves = gpu.render(p.copy(sample_begin=0, sample_end=8)).astype(np.float64) + gpu.render(p.copy(sample_begin=8, sample_end=16))
    np.testing.assert_allclose(halves, whole, rtol=3e-7, atol=1e-7)


def test_tile_shards_reassemble_bit_identically(gpu):
    p = gpu.default_params(width=1000, height=333, spp=8, method=0, seed=4)   # 333000 pixels: not a multiple of the 128-pixel tile
    whole = gpu.render(p)
    parts = [gpu.render(p.copy(tile_rank=r, tile_count=3)) for r in range(3)]
    cover = sum((q != 0).any(axis=-1).astype(int) for q in parts)
    assert cover.max() <= 1                                                  # tiles are disjoint
    assert np.array_equal(parts[0] + parts[1] + parts[2], whole)             # adding zeros: bit-identical to one GPU
    flat = parts[1].reshape(-1, 3); owner = (np.arange(flat.shape[0]) // 128) % 3
    assert not flat[owner != 1].any()


def test_render_multi_single_device_equals_render(gpu):
    p = gpu.default_params(width=320, height=200, spp=4, method=1, seed=6)
    a = gpu.render(p)
    b, st = gpu.render_multi(p, None, [0], stats=True)
    c = gpu.render_multi(p, None, [0, 0, 0])                                  # three tile shards on the same device
    assert np.array_equal(a, b) and np.array_equal(a, c) and st.paths == 320 * 200 * 4


def test_statistics_match_the_reference_workload(gpu):
    """events per path = 1/(1-0.6) * 0.6 = 1.5; scans per path about 4.3 (the reference's 5.4 minus
    the scans this implementation proves redundant)"""
    for method in (0, 1, 2, 4):
        _, st = gpu.render(gpu.default_params(spp=16, method=method), stats=True)
        assert st.paths == 1024 * 768 * 16 and abs(st.events / st.paths - 1.5) < 0.01
        assert 3.5 < st.scene_scans / st.paths < 5.6 and st.kernel_ms > 0 and st.launches == 1


# ---- edge cases -------------------------------------------------------------------------------------------------------------------
def test_edge_cases(gpu):
    one = gpu.render(gpu.default_params(width=1, height=1, spp=1))
    assert one.shape == (1, 1, 3) and np.isfinite(one).all()
    odd = gpu.render(gpu.default_params(width=37, height=5, spp=3, method=2))
    assert np.isfinite(odd).all() and odd.max() > 0
    # no emitter: every path returns black (vptShadeMethods.h:1301)
    rows = DEFAULT_SCENE.copy(); rows[:, 7:10] = 0
    black, st = gpu.render(gpu.default_params(width=64, height=48, spp=4), gpu.scene_from_rows(rows), stats=True)
    assert not black.any() and st.launches == 0
    # a single emitting sphere seen directly: 0.6 * Le inside its silhouette (roulette at depth 0 without 1/cp, :1308-1312), 0 outside
    solo = np.zeros((1, 18)); solo[0, 0] = 20; solo[0, 1:4] = [0, 11.2, 100]; solo[0, 7:10] = [3, 2, 1]
    img = gpu.render(gpu.default_params(width=64, height=48, spp=512, sigma_s=1e-9, sigma_a=1e-9), gpu.scene_from_rows(solo))
    np.testing.assert_allclose(img[24, 32], [1.8, 1.2, 0.6], rtol=0.15)
    assert img[0, 0].max() < 1e-9                                            # only in-scattered light from the (almost) clear medium
    # more than four emitters (the reference overflows arr[4]): a checked, working configuration here
    rows = np.vstack([DEFAULT_SCENE] + [DEFAULT_SCENE[9:10] + np.r_[0, 6.0 * k, 0, 0, np.zeros(14)] for k in range(1, 4)])
    many = gpu.render(gpu.default_params(width=64, height=48, spp=8), gpu.scene_from_rows(rows))
    assert np.isfinite(many).all() and many.mean() > 0


def test_precision_quirk_contract(gpu):
    with pytest.raises(gpu.VptError) as e:
        gpu.render(gpu.default_params(width=8, height=8, spp=1, quirks=gpu.QUIRKS_REFERENCE))
    assert e.value.status == -3
    with pytest.raises(gpu.VptError):
        gpu.render(gpu.default_params(width=8, height=8, spp=1, device=99))


# ---- the FP32 kernel variants compute the same thing -----------------------------------------------------------------------------
@pytest.mark.parametrize("method", [0, 1, 2, 4])
def test_kernel_variants_agree(gpu, l1, method):
    """MEGA (one thread per pixel drives each path through the stages), WAVEFRONT_SM (one pool per SM, AUTO) and WAVEFRONT_HBM (multi-kernel,
    queues in HBM) run the SAME stage code (csrc/vpt_stages.cuh) on the same Philox streams: identical decisions, hence identical event and
    scan counts; the two wavefronts add the same contributions into fixed-point sums (bit-identical images), the megakernel sums them in
    double (equal to the fixed-point rounding); all against the oracle.  The superseded MEGA_SCAN / WAVEFRONT variants are no longer built."""
    w, h, spp = 160, 120, 16
    p = gpu.default_params(width=w, height=h, spp=spp, method=method, seed=12, output=gpu.OUTPUT_SUM)
    a, sa = gpu.render(p.copy(kernel=gpu.KERNEL_MEGA), stats=True)
    d, sd = gpu.render(p.copy(kernel=gpu.KERNEL_WAVEFRONT_SM), stats=True)
    e_, se = gpu.render(p.copy(kernel=gpu.KERNEL_WAVEFRONT_HBM), stats=True)
    for so in (sd, se):
        assert so.events == sa.events and so.scene_scans == sa.scene_scans and so.paths == w * h * spp and so.nonfinite == 0
    for other in (d, e_):
        np.testing.assert_allclose(other, a, rtol=2e-6, atol=spp * 2e-9)   # fixed-point 2^-30 per contribution against a double sum
    for kern, img in ((gpu.KERNEL_MEGA, a), (gpu.KERNEL_WAVEFRONT_SM, d), (gpu.KERNEL_WAVEFRONT_HBM, e_)):  # reruns are bit-identical
        assert np.array_equal(img, gpu.render(p.copy(kernel=kern)))
    assert np.array_equal(d, e_)   # same paths, same fixed-point sums: the HBM wavefront reproduces the on-chip one bit for bit
    ref, _, rst = l1.render(DEFAULT_SCENE, 0, method, SA, SS, w, h, 12, spp, want_sumsq=False)
    for img in (a, d, e_):
        e = np.abs(img - ref) / np.maximum(np.abs(ref), 1e-3)
        assert np.median(e) < 2e-6 and np.mean(e > 1e-3) < 0.02
    for kern in (gpu.KERNEL_MEGA_SCAN, gpu.KERNEL_WAVEFRONT):
        with pytest.raises(gpu.VptError) as err:
            gpu.render(p.copy(kernel=kern))
        assert err.value.status == -3


@pytest.mark.parametrize("quirks", [0, 3])
def test_fp64_wavefront_equals_the_fp64_megakernel(gpu, quirks):
    """FP64 reference mode: the SM-wide wavefront (AUTO; csrc/vpt_smwave_f64.cuh) runs the same three vertex parts (csrc/vpt_f64.cuh
    vertex_primary / vertex_medium / vertex_surface) as the one-thread-per-pixel kernel (VPT_KERNEL_MEGA) composes into vertex(): identical
    decisions, identical scan and event counts, pixel sums equal to the 2^-34 fixed-point rounding; reruns bit-identical."""
    w, h, spp = 160, 120, 8
    for method in (0, 1, 2, 4):
        p = gpu.default_params(width=w, height=h, spp=spp, method=method, seed=15, output=gpu.OUTPUT_SUM, precision=gpu.PRECISION_FP64_REF, quirks=quirks)
        a, sa = gpu.render(p.copy(kernel=gpu.KERNEL_MEGA), stats=True)
        b, sb = gpu.render(p, stats=True)
        assert (sa.paths, sa.events, sa.scene_scans, sa.nonfinite) == (sb.paths, sb.events, sb.scene_scans, sb.nonfinite)
        np.testing.assert_allclose(b, a, rtol=3e-7, atol=spp * 3 * 2.0 ** -34)
        assert np.array_equal(b, gpu.render(p))
        c = gpu.render(p.copy(tile_rank=1, tile_count=3))
        owner = (np.arange(w * h) // 128) % 3
        assert np.array_equal(c.reshape(-1, 3)[owner == 1], b.reshape(-1, 3)[owner == 1]) and not c.reshape(-1, 3)[owner != 1].any()


@pytest.mark.parametrize("method", [0, 1, 2, 4])
def test_auto_kernel_per_path_parity(gpu, l1, method):
    """spp = 1: every pixel of the product kernel (AUTO = the SM-wide wavefront) is ONE path; compare each with the FP64 oracle on the same
    Philox stream.  Tolerance: 1e-5 relative (north_star's unit tolerance) for at least 99 % of the paths -- the remainder are paths where
    an fp32-rounded decision (hit / miss of a light cone edge, surface / medium at Tr ~ xi) legitimately differs -- and equal event counts."""
    w, h = 256, 192
    p = gpu.default_params(width=w, height=h, spp=1, method=method, seed=77, output=gpu.OUTPUT_SUM)
    img, st = gpu.render(p, stats=True)
    ref, _, rst = l1.render(DEFAULT_SCENE, 0, method, SA, SS, w, h, 77, 1, want_sumsq=False)
    assert st.paths == w * h and st.nonfinite == 0
    assert abs(int(st.events) - int(rst["events"])) <= 2e-4 * rst["events"]
    err = np.abs(img - ref).max(axis=2) / np.maximum(np.abs(ref).max(axis=2), 1e-4)
    assert np.median(err) < 1e-6
    assert np.mean(err < 1e-5) > 0.99, float(np.mean(err < 1e-5))
    assert np.array_equal((img == 0).all(axis=2) & (ref == 0).all(axis=2), (ref == 0).all(axis=2)) or np.mean((img == 0).all(axis=2) != (ref == 0).all(axis=2)) < 2e-3


def _stress_scene():
    """Not in the reference: exercises what the default scene does not -- an ODD number of area lights (the cone-sample uniforms come in pairs
    per Philox block), two point lights, a mid-size sphere (r = 90: general root form without re-anchoring), a second microfacet object, an
    emitter with a zero red channel (MISv2 only samples lights with radiance.x > 0, misSamplingFunctions.h:106), 15 spheres."""
    rows = [r.copy() for r in DEFAULT_SCENE]
    z = [0.0] * 7
    rows.append(np.array([90, 60, -125, -30, .6, .4, .3, 0, 0, 0, 0, *z]))                    # r = 90 Lambert boulder poking through the floor
    rows.append(np.array([1.5, -20, 10, 40, 0, 0, 0, 40, 60, 90, 0, *z]))                      # third area light
    rows.append(np.array([0, 30, -10, 60, 0, 0, 0, 0, 900, 900, 0, *z]))                       # second point light
    rows.append(np.array([6, -5, -34.8, 60, 0, 0, 0, 0, 0, 0, 1, 0.2, 0.92, 1.1, 3.9, 2.45, 2.14, 0.2]))  # rough gold-ish microfacet ball
    rows.append(np.array([1.0, 10, 0, 90, 0, 0, 0, 0, 20, 20, 0, *z]))                         # emitter with radiance.x == 0
    return np.array(rows)


@pytest.mark.parametrize("n_big", [1, 2])
def test_auto_kernel_per_path_parity_at_the_sphere_limit(gpu, l1, n_big):
    """VPT_MAX_SPHERES = 32 spheres (the reference's std::vector is unbounded): the default scene plus small Lambert / microfacet balls and
    n_big spheres of radius >= 64 (general root form), so that the paired scan records of BOTH classes are exercised with even and odd counts
    (6 or 7 general-form spheres, 25 or 24 direct-root ones incl. the lights; an odd class carries a padding record no ray may hit)"""
    rng = np.random.default_rng(40 + n_big)
    rows = [r.copy() for r in DEFAULT_SCENE]
    z = [0.0] * 7
    for k in range(n_big):
        rows.append(np.array([70.0 + 10 * k, -80 + 160 * k, -95 - 10 * k, 20, .4, .5, .6, 0, 0, 0, 0, *z]))
    while len(rows) < 32:
        c = [rng.uniform(-40, 40), rng.uniform(-35, 20), rng.uniform(-70, 120)]
        if rng.random() < 0.25:
            rows.append(np.array([rng.uniform(2, 5), *c, 0, 0, 0, 0, 0, 0, 1, 0.2, 0.92, 1.1, 3.9, 2.45, 2.14, rng.uniform(0.1, 0.3)]))
        else:
            rows.append(np.array([rng.uniform(2, 6), *c, *rng.uniform(0.2, 0.9, 3), 0, 0, 0, 0, *z]))
    sc = np.array(rows)
    assert len(sc) == 32
    w, h = 192, 144
    for method in (0, 1):
        p = gpu.default_params(width=w, height=h, spp=1, method=method, seed=17, output=gpu.OUTPUT_SUM)
        img, st = gpu.render(p, gpu.scene_from_rows(sc), stats=True)
        ref, _, rst = l1.render(sc, 0, method, SA, SS, w, h, 17, 1, want_sumsq=False)
        assert st.paths == w * h and st.nonfinite == 0
        assert abs(int(st.events) - int(rst["events"])) <= 5e-4 * rst["events"] + 2
        err = np.abs(img - ref).max(axis=2) / np.maximum(np.abs(ref).max(axis=2), 1e-4)
        assert np.median(err) < 1e-6 and np.mean(err < 1e-5) > 0.98, (float(np.median(err)), float(np.mean(err < 1e-5)))
        assert np.array_equal(img, gpu.render(p.copy(kernel=gpu.KERNEL_WAVEFRONT_HBM), gpu.scene_from_rows(sc)))
    with pytest.raises(gpu.VptError) as e:   # 33 spheres: a checked limit
        gpu.render(gpu.default_params(width=8, height=8, spp=1), gpu.scene_from_rows(np.vstack([sc, sc[-1:]])))
    assert e.value.status == -2


@pytest.mark.parametrize("method", [0, 1, 2, 4])
def test_auto_kernel_per_path_parity_on_a_scene_not_in_the_reference(gpu, l1, method):
    """the product kernel against the FP64 oracle, one path per pixel, on the stress scene"""
    sc = _stress_scene()
    w, h = 192, 144
    p = gpu.default_params(width=w, height=h, spp=1, method=method, seed=31, output=gpu.OUTPUT_SUM)
    img, st = gpu.render(p, gpu.scene_from_rows(sc), stats=True)
    ref, _, rst = l1.render(sc, 0, method, SA, SS, w, h, 31, 1, want_sumsq=False)
    assert st.paths == w * h and st.nonfinite == 0
    assert abs(int(st.events) - int(rst["events"])) <= 3e-4 * rst["events"]
    err = np.abs(img - ref).max(axis=2) / np.maximum(np.abs(ref).max(axis=2), 1e-4)
    assert np.median(err) < 1e-6 and np.mean(err < 1e-5) > 0.985, (float(np.median(err)), float(np.mean(err < 1e-5)))
    # and with more samples the two images agree pixel by pixel (sums of 8 paths)
    p8 = p.copy(spp=8)
    img8 = gpu.render(p8, gpu.scene_from_rows(sc))
    ref8, _, _ = l1.render(sc, 0, method, SA, SS, w, h, 31, 8, want_sumsq=False)
    e8 = np.abs(img8 - ref8) / np.maximum(np.abs(ref8), 1e-3)
    assert np.median(e8) < 2e-6 and np.mean(e8 > 1e-3) < 0.03


def test_concurrent_renders_with_different_seeds(gpu):
    """the product kernel reads its Philox round keys from constant memory, one schedule per device at a time (csrc/vpt_kernels_f32.cu
    philox_keys_begin): launches of different seeds from different host threads / streams must order themselves on the device.  Four threads
    with four seeds (and two frame sizes, so that launches overlap differently) render concurrently; every frame must equal the frame the
    same parameters give alone."""
    import threading
    jobs = [gpu.default_params(width=w, height=h, spp=spp, method=m, seed=s, output=gpu.OUTPUT_SUM)
            for (w, h, spp, m, s) in ((320, 240, 24, 1, 11), (256, 192, 40, 0, 12), (320, 240, 24, 2, 13), (192, 144, 64, 4, 14))]
    want = [gpu.render(p) for p in jobs]
    errors = []

    def work(i):
        try:
            for _ in range(12):
                if not np.array_equal(gpu.render(jobs[i]), want[i]):
                    errors.append("job %d: frame differs under concurrency" % i)
                    return
        except Exception as e:  # noqa: BLE001
            errors.append("job %d: %r" % (i, e))
    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    # same seed on two threads: both share the schedule, nothing waits
    same = [threading.Thread(target=work, args=(0,)) for _ in range(3)]
    for t in same:
        t.start()
    for t in same:
        t.join()
    assert not errors, errors
