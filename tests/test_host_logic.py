"""CPU: host-side logic around the kernels -- the output stage (clamp, gamma, P3 text; rt.cpp:803-820, mathUtilities.h:34-45),
shard arithmetic, the CLI surface."""
import os

import numpy as np
import pytest


def test_tonemap_is_the_references(vpt, units):
    x = units["tonemap_in"].astype(np.float32)
    hdr = np.stack([x, x, x], axis=-1).reshape(1, -1, 3)
    got = vpt.tonemap(hdr)[0, :, 0]
    # the golden values were computed by the reference on the float64 inputs; float32 rounding of x may move a value that
    # sits exactly on a bucket edge by one step -- compare on the rounded inputs instead
    want = np.array([int(np.clip(float(v), 0, 1) ** (1 / 2.2) * 255 + .5) for v in x])
    assert np.array_equal(got, want)
    exact = units["tonemap_out"]
    assert np.mean(got == exact) > 0.99 and np.max(np.abs(got.astype(int) - exact)) <= 1


def test_ppm_bytes(vpt, tmp_path):
    rng = np.random.default_rng(0)
    hdr = rng.uniform(-0.1, 1.2, size=(3, 5, 3)).astype(np.float32)
    path = tmp_path / "image.ppm"
    vpt.write_ppm(hdr, str(path))
    text = path.read_text()
    rgb = vpt.tonemap(hdr).reshape(-1, 3)
    want = "P3\n5 3\n255\n" + "".join("%d %d %d " % tuple(px) for px in rgb)  # rt.cpp:814-820: no newlines after the header
    assert text == want


def test_ppm_is_byte_identical_to_a_file_written_by_the_reference_binary(vpt, tmp_path):
    """`oracle/_ref/ref_rt 1` (the unmodified src/rt.cpp, rt.cpp:808-820) writes image.ppm; the 8-bit values it holds are mapped back to an HDR
    frame that tonemaps to exactly those values, and vpt_write_ppm of that frame must reproduce the reference's file byte for byte:
    header, separators, trailing blank, no newline after the header."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_rt")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_rt not built (reference tree absent)")
    subprocess.run([exe, "1"], cwd=tmp_path, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300,
                   env=dict(os.environ, OMP_NUM_THREADS=str(min(os.cpu_count() or 1, 8))))
    ref = (tmp_path / "image.ppm").read_bytes()
    head, body = ref.split(b"\n", 3)[:3], ref.split(b"\n", 3)[3]
    assert head == [b"P3", b"1024 768", b"255"] and b"\n" not in body and body.endswith(b" ")
    v = np.array(body.split(), dtype=np.int64)
    assert v.size == 1024 * 768 * 3 and v.min() >= 0 and v.max() <= 255 and len(set(v.tolist())) > 50   # a real image, not a constant
    hdr = ((v / 255.0) ** 2.2).astype(np.float32).reshape(768, 1024, 3)      # toDisplayValue(hdr) == v (mathUtilities.h:34-45)
    assert np.array_equal(vpt.tonemap(hdr).reshape(-1), v)
    out = tmp_path / "ours.ppm"
    vpt.write_ppm(hdr, str(out))
    assert out.read_bytes() == ref


def test_pfm_side_output(vpt, tmp_path):
    """the frame before the tonemap as a binary PFM.  Header, byte order, bottom row first -- checked byte by byte against
    the format's definition -- and a round trip through read_pfm; negative values, values above one, NaN and infinity pass unclamped."""
    rng = np.random.default_rng(3)
    hdr = rng.uniform(-0.5, 40.0, size=(4, 7, 3)).astype(np.float32)
    hdr[1, 2] = (np.inf, -0.0, np.nan)
    path = tmp_path / "frame.pfm"
    vpt.write_pfm(hdr, str(path))
    raw = path.read_bytes()
    head = b"PF\n7 4\n-1.0\n"
    assert raw[:len(head)] == head and len(raw) == len(head) + 4 * 7 * 3 * 4
    assert raw[len(head):] == hdr[::-1].astype("<f4").tobytes()          # rows bottom to top, little-endian floats
    back = vpt.read_pfm(str(path))
    assert back.shape == hdr.shape and back.tobytes() == hdr.tobytes()
    with pytest.raises(vpt.VptError):
        vpt.write_pfm(hdr, str(tmp_path / "no_such_dir" / "x.pfm"))


def test_sample_shards_partition_the_range():
    from minimal_volumetric_path_tracer_b200 import distributed as d
    for spp in (1, 7, 64, 1024, 16384):
        for world in (1, 2, 3, 4, 8):
            edges = [d.sample_shard(spp, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == spp
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        d.sample_shard(8, 2, 2)


def test_tile_ownership_is_interleaved():
    from minimal_volumetric_path_tracer_b200 import distributed as d
    owners = [d.tile_owner(p, 4) for p in range(0, 128 * 9, 128)]
    assert owners == [0, 1, 2, 3, 0, 1, 2, 3, 0]


def test_shard_params(vpt):
    from minimal_volumetric_path_tracer_b200 import distributed as d
    p = vpt.default_params(spp=10)
    q, work = d.shard_params(p, "samples", 1, 4)
    assert (q.sample_begin, q.sample_end, q.output, work) == (3, 6, vpt.OUTPUT_SUM, True)
    q, work = d.shard_params(vpt.default_params(spp=2), "samples", 3, 4)
    assert not work
    q, work = d.shard_params(p, "tiles", 2, 4)
    assert (q.tile_rank, q.tile_count, q.sample_begin, q.sample_end) == (2, 4, 0, 0)
    assert p.tile_count == 0  # the caller's params are not modified


def test_cli_surface(vpt):
    from minimal_volumetric_path_tracer_b200 import cli
    a = cli.parse_args(["16"])
    p = cli.params_from_args(a)
    assert (p.spp, p.width, p.height, p.method, a.output) == (16, 1024, 768, 0, "image.ppm")  # the reference's `rt <spp>`
    a = cli.parse_args(["4", "--method", "equi", "--size", "64x48", "--ref", "--seed", "9", "-o", "x.ppm"])
    p = cli.params_from_args(a)
    assert (p.method, p.width, p.height, p.precision, p.quirks, p.seed) == (1, 64, 48, vpt.PRECISION_FP64_REF, 3, 9)
    with pytest.raises(SystemExit):
        cli.parse_args([])  # the reference segfaults on a missing argument; here it is a usage error
    with pytest.raises(SystemExit):
        cli.parse_args(["0"])


def test_cpp_host_binary_exists_and_reports_usage(vpt):
    import subprocess
    from minimal_volumetric_path_tracer_b200 import build
    assert os.path.exists(build.RT)
    r = subprocess.run([build.RT], capture_output=True, text=True)
    assert r.returncode == 2 and "usage: rt <spp>" in r.stderr
