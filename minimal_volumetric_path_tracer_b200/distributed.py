"""Multi-GPU sharding of one frame: one process per GPU (torch.distributed), no exchange while rendering, ONE reduce of the
fp32 HDR buffers at the end (NCCL over NVLink on GPUs; gloo in the CPU tests).  SURVEY.md section 8e.

Two partitions of the reference's pixel loop (src/rt.cpp:768-798), both leave every Philox stream (pixel, sample) untouched,
so the partition never changes a sample:
  "samples": rank r renders samples [b_r, e_r) of every pixel (best balance; the reduce changes the summation order of the
             low bits only),
  "tiles":   rank r renders the 128-pixel tiles with tile_id % world == r into a zeroed full-size buffer; adding zeros keeps
             the combined frame bit-identical to a single-GPU render.
"""
import ctypes as C

from . import api

TILE = 128  # csrc/vpt_internal.h kTile


def sample_shard(spp, rank, world):
    """contiguous near-equal split of [0, spp); the first spp % world ranks get one extra sample"""
    if world <= 0 or not (0 <= rank < world) or spp < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(spp, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def tile_owner(pixel_index, world):
    return (pixel_index // TILE) % world


def shard_params(params, mode, rank, world):
    """Params for this rank's share of the frame described by `params` (which must describe the WHOLE frame). Output is SUM."""
    if mode == "samples":
        b, e = sample_shard(params.spp, rank, world)
        return params.copy(sample_begin=b, sample_end=e, tile_rank=0, tile_count=0, output=api.OUTPUT_SUM), (e > b)
    if mode == "tiles":
        return params.copy(sample_begin=0, sample_end=0, tile_rank=rank, tile_count=world, output=api.OUTPUT_SUM), True
    raise ValueError("mode must be 'samples' or 'tiles'")


def _render_cuda(params, scene, torch):
    if torch.cuda.current_device() != params.device:
        # one rank per GPU: the frame is reduced by NCCL on torch's current device; a rank that leaves params.device at its default 0 would
        # render on GPU 0 next to rank 0 and hand NCCL a buffer of another device
        raise ValueError("params.device = %d but this process's current CUDA device is %d: pass device=local_rank (and torch.cuda.set_device(local_rank))"
                         % (params.device, torch.cuda.current_device()))
    dev = torch.device("cuda", params.device)
    hdr = torch.empty((params.height, params.width, 3), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    api.render_device(params, scene, hdr.data_ptr(), stream)
    return hdr


def render_sharded(params, scene=None, mode="samples", group=None, dst=0, render_fn=None, mean=True):
    """Render this process's shard and combine with ONE reduce.  Returns the combined (h, w, 3) tensor on rank `dst`
    (None elsewhere).  `render_fn(params, scene) -> tensor` is injectable so that the sharding / combine logic is testable on
    CPU with gloo; by default the CUDA kernels are used (there is no CPU renderer in this package)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    scene = scene if scene is not None else api.default_scene()
    mine, has_work = shard_params(params, mode, rank, world)
    fn = render_fn or (lambda p, s: _render_cuda(p, s, torch))
    if has_work:
        hdr = fn(mine, scene)
    else:  # more ranks than samples
        ref = fn(mine.copy(sample_begin=0, sample_end=1), scene)
        hdr = torch.zeros_like(ref)
    if world > 1:
        dist.reduce(hdr, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if rank != dst:
        return None
    if mean:
        hdr = hdr * (1.0 / params.spp)
    return hdr
