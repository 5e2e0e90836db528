"""CPU, world_size 2, gloo: the multi-GPU combine path (shard -> render -> ONE reduce) with an injected stand-in renderer
(a deterministic function of (pixel, sample)), since this container has no GPU.  What is checked is the host logic that is
the same on NCCL: every (pixel, sample) is rendered by exactly one rank and the reduce reassembles the whole frame."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, SPP = 40, 13, 7  # 520 pixels: 4 full 128-pixel tiles + a partial one


def fake_render(params, scene):
    """value(pixel, sample) = pixel + 1000 * sample in channel 0, 1 in channel 1 (a coverage counter)"""
    from minimal_volumetric_path_tracer_b200 import distributed as d
    b, e = (params.sample_begin, params.sample_end) if params.sample_end else (0, params.spp)
    pix = torch.arange(params.width * params.height, dtype=torch.float64)
    owned = torch.ones_like(pix) if params.tile_count <= 1 else ((pix.long() // d.TILE) % params.tile_count == params.tile_rank).double()
    out = torch.zeros(params.height * params.width, 3, dtype=torch.float32)
    for s in range(b, e):
        out[:, 0] += (owned * (pix + 1000.0 * s)).float()
        out[:, 1] += owned.float()
    return out.reshape(params.height, params.width, 3)


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import minimal_volumetric_path_tracer_b200 as v
    from minimal_volumetric_path_tracer_b200 import distributed as d
    p = v.default_params(width=W, height=H, spp=SPP)
    out = d.render_sharded(p, v.default_scene(), mode=mode, render_fn=fake_render, mean=False)
    if rank == 0:
        q.put(out.numpy())
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close(); return port


@pytest.mark.parametrize("mode", ["samples", "tiles"])
def test_two_rank_combine(mode, vpt):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    whole = fake_render(vpt.default_params(width=W, height=H, spp=SPP), None).numpy()
    assert np.array_equal(got[..., 1], np.full((H, W), SPP, dtype=np.float32))  # every (pixel, sample) exactly once
    np.testing.assert_allclose(got[..., 0], whole[..., 0], rtol=1e-6)


def test_single_process_path(vpt):
    from minimal_volumetric_path_tracer_b200 import distributed as d
    p = vpt.default_params(width=W, height=H, spp=SPP)
    out = d.render_sharded(p, vpt.default_scene(), mode="samples", render_fn=fake_render, mean=True)
    np.testing.assert_allclose(out[..., 1].numpy(), 1.0, rtol=1e-6)
