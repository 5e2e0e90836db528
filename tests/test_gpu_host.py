"""GPU: the reference-facing hosts.  The C++ `rt` keeps the reference's command line (rt <spp> -> image.ppm, "elapsed time"
on stdout) and must produce the same bytes as the Python host for the same seed; the device-buffer entry point and the torch
plumbing used by bench.py are exercised too."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cpp_host_writes_the_same_ppm_as_the_python_host(gpu, tmp_path):
    from minimal_volumetric_path_tracer_b200 import build, cli
    out_cpp = tmp_path / "image.ppm"
    r = subprocess.run([build.RT, "4", "--size", "160x120", "--method", "mis", "--seed", "5"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.startswith("elapsed time: ") and r.stdout.strip().endswith("s")      # rt.cpp:827
    text = out_cpp.read_text()
    assert text.startswith("P3\n160 120\n255\n") and text.count("\n") == 3                # rt.cpp:814-820
    out_py = tmp_path / "py.ppm"
    assert cli.main(["4", "--size", "160x120", "--method", "mis", "--seed", "5", "-o", str(out_py)]) == 0
    assert out_py.read_text() == text
    vals = np.array(text.split()[4:], dtype=int)
    assert vals.size == 160 * 120 * 3 and vals.min() >= 0 and vals.max() <= 255 and vals.mean() > 5


def test_pfm_side_output_of_both_hosts(gpu, tmp_path):
    """--pfm: the frame before the tonemap.  Both hosts write the same floats for the same seed, they are the frame vpt_render returns,
    and tonemapping them gives the PPM written next to them"""
    from minimal_volumetric_path_tracer_b200 import build, cli
    args = ["8", "--size", "96x64", "--method", "equi", "--seed", "7"]
    r = subprocess.run([build.RT] + args + ["-o", "c.ppm", "--pfm", "c.pfm"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert cli.main(args + ["-o", str(tmp_path / "p.ppm"), "--pfm", str(tmp_path / "p.pfm")]) == 0
    a, b = gpu.read_pfm(str(tmp_path / "c.pfm")), gpu.read_pfm(str(tmp_path / "p.pfm"))
    assert a.shape == (64, 96, 3) and np.array_equal(a, b)
    assert np.array_equal(a, gpu.render(gpu.default_params(width=96, height=64, spp=8, method=1, seed=7)))
    vals = np.array((tmp_path / "c.ppm").read_text().split()[4:], dtype=int)
    assert np.array_equal(gpu.tonemap(a).reshape(-1), vals)


def test_ref_flag_runs_fp64(gpu, tmp_path):
    from minimal_volumetric_path_tracer_b200 import build
    r = subprocess.run([build.RT, "2", "--size", "64x48", "--ref", "-o", "ref.ppm"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0 and os.path.exists(tmp_path / "ref.ppm")


def test_device_buffer_entry_point_with_torch(gpu):
    import torch
    p = gpu.default_params(width=256, height=192, spp=8, method=1, seed=2)
    want = gpu.render(p)
    hdr = torch.zeros((192, 256, 3), dtype=torch.float32, device="cuda:0")
    st = gpu.Stats()
    gpu.render_device(p, gpu.default_scene(), hdr.data_ptr(), torch.cuda.current_stream().cuda_stream, st)
    assert np.array_equal(hdr.cpu().numpy(), want) and st.launches == 1 and st.kernel_ms > 0
    hdr.zero_()
    gpu.render_device(p, gpu.default_scene(), hdr.data_ptr(), torch.cuda.current_stream().cuda_stream)  # asynchronous form
    torch.cuda.synchronize()
    assert np.array_equal(hdr.cpu().numpy(), want)


def test_sharded_render_single_rank(gpu):
    from minimal_volumetric_path_tracer_b200 import distributed as d
    p = gpu.default_params(width=128, height=96, spp=6, method=0, seed=3)
    out = d.render_sharded(p, mode="samples")
    np.testing.assert_allclose(out.cpu().numpy(), gpu.render(p), rtol=1e-6, atol=1e-8)


def test_fp32_peak_probe(gpu):
    tflops, clk = gpu.measure_fp32_peak()
    assert 30 < tflops < 90 and clk > 1000    # nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4 TFLOP/s
