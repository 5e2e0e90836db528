"""Multi-GPU equivalence check (run under torchrun on N GPUs of one box): the frame combined by distributed.render_sharded -- one rank per
GPU, ONE NCCL reduce of the HDR buffers -- against the same frame rendered by rank 0 alone.
    tiles:   bit-identical (disjoint tiles, adding zeros);   samples: identical up to the fp32 summation order of N partial sums.
Also vpt_render_multi (one process, N devices, no reduction) against the same frame.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_multi_gpu.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import minimal_volumetric_path_tracer_b200 as v
from minimal_volumetric_path_tracer_b200 import distributed as d

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for method in (0, 1, 2, 4):
    p = v.default_params(width=1024, height=768, spp=64, method=method, seed=21, device=local)
    tiles = d.render_sharded(p, mode="tiles", mean=False)
    samples = d.render_sharded(p, mode="samples", mean=False)
    torch.cuda.synchronize()
    if rank == 0:
        whole = v.render(p.copy(output=v.OUTPUT_SUM))
        t, s = tiles.cpu().numpy(), samples.cpu().numpy()
        same_tiles = np.array_equal(t, whole)
        err = np.abs(s - whole) / np.maximum(np.abs(whole), 1e-6)
        multi = v.render_multi(p.copy(output=v.OUTPUT_SUM), None, list(range(world)))
        same_multi = np.array_equal(multi, whole)
        good = same_tiles and same_multi and err.max() < 2e-6
        ok &= bool(good)
        print("method %d, %d GPUs: tile shards + NCCL reduce bit-identical to one GPU: %s; sample shards max rel diff %.2e; vpt_render_multi bit-identical: %s -> %s"
              % (method, world, same_tiles, err.max(), same_multi, "ok" if good else "FAIL"), flush=True)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("ALL OK" if ok else "FAILED", flush=True)
    sys.exit(0 if ok else 1)
