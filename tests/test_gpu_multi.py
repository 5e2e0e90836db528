"""GPU, >= 2 devices: the multi-GPU paths against one device (split of the loop src/rt.cpp:767-798).
  - vpt_render_multi (one process, one host thread per device, interleaved tiles, no reduction): bit-identical to one device;
  - one process per GPU + ONE NCCL reduce of the HDR buffers (distributed.render_sharded): tile shards bit-identical (adding zeros),
    sample shards identical up to the fp32 summation order of N partial sums (<= 1e-6 relative).
Skipped on a single-GPU lease (the CPU twin of the combine logic is tests/test_distributed_cpu.py, gloo, world size 2)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, SPP = 640, 360, 32


def _devices(gpu):
    n = gpu.device_count()
    if n < 2:
        pytest.skip("needs at least two CUDA devices (this lease has %d)" % n)
    return min(n, 8)


@pytest.mark.parametrize("method", [0, 2, 4])
def test_render_multi_over_all_devices_is_bit_identical(gpu, method):
    n = _devices(gpu)
    p = gpu.default_params(width=W, height=H, spp=SPP, method=method, seed=21)
    one, st1 = gpu.render(p, stats=True)
    multi, stn = gpu.render_multi(p, None, list(range(n)), stats=True)
    assert np.array_equal(one, multi)
    assert (stn.paths, stn.events, stn.scene_scans) == (st1.paths, st1.events, st1.scene_scans)
    # page-locked destination frame and a device subset in another order (tile shard k runs on devices[k])
    frame = gpu.PinnedFrame(H, W)
    lib = gpu.load_library()
    import ctypes as C
    dev = (C.c_int32 * 2)(n - 1, 0)
    rc = lib.vpt_render_multi(C.byref(p), gpu.default_scene(), 10, dev, 2, frame.array.ctypes.data_as(C.POINTER(C.c_float)), None)
    assert rc == 0 and np.array_equal(frame.array, one)
    frame.close()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import minimal_volumetric_path_tracer_b200 as v
    from minimal_volumetric_path_tracer_b200 import distributed as d
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    out = {}
    for method in (1, 2):
        p = v.default_params(width=W, height=H, spp=SPP, method=method, seed=21, device=rank)
        tiles = d.render_sharded(p, mode="tiles", mean=False)
        samples = d.render_sharded(p, mode="samples", mean=False)
        torch.cuda.synchronize()
        assert torch.cuda.current_device() == rank          # the library leaves the caller's current device alone
        if rank == 0:
            out[method] = (tiles.cpu().numpy(), samples.cpu().numpy(), v.render(p.copy(output=v.OUTPUT_SUM)))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        q.put(out)


def test_one_rank_per_gpu_nccl_reduce_matches_one_gpu(gpu):
    import torch.multiprocessing as mp
    world = _devices(gpu)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=200)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for method, (tiles, samples, whole) in out.items():
        assert np.array_equal(tiles, whole), "method %d: tile shards + NCCL reduce differ from one GPU" % method
        err = np.abs(samples - whole) / np.maximum(np.abs(whole), 1e-6)
        assert err.max() < 1e-6, "method %d: sample shards differ by %.2e" % (method, err.max())
