#!/usr/bin/env python
"""bench.py -- Mpaths/s of the per-pixel radiance loop (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2|c1|c3|c4|c5]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] = the reference's default scene, equi-angular shade method, 1024x768 at
1024 spp -- one "step" is one full render of it (8.05e8 camera paths).  With N GPUs the frame is sharded by SAMPLES: rank r
renders 1024 samples per pixel of a 1024*N-spp frame (weak scaling), and ONE NCCL reduce of the fp32 HDR buffers inside the
timed region combines them.  `value` counts camera paths of all ranks over the max-over-ranks device time.

`e2e` is the same metric through the reference-facing C-ABI call vpt_render() with HOST buffers (scene + params in, HDR
frame out, every step).  `roofline` is FP32 CUDA-core throughput: algorithmic 1700 FLOP per camera path
against the FFMA peak measured live on this GPU.  `cpu_baseline` / `--impl reference` time the reference's own CPU code
(oracle/_ref, compiled from the unmodified sources) on this box's host cores."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PATH = 1700.0  # 18*S*T + 460*E at S=10 spheres, T=5.5 scans, E=1.5 events

CONFIGS = {  # BASELINE.json configs, SURVEY.md section 8d
    "c1": dict(name="C1 default scene, free-flight, 1024x768 @ 64 spp", method=0, width=1024, height=768, spp=64),
    "c2": dict(name="C2 default scene, equi-angular, 1024x768 @ 1024 spp", method=1, width=1024, height=768, spp=1024),
    "c3": dict(name="C3 default scene, MIS, 1920x1080 @ 4096 spp", method=2, width=1920, height=1080, spp=4096),
    "c4": dict(name="C4 dense high-albedo medium, MIS, 1024x768 @ 256 spp, continue_prob 0.95, max depth 64", method=2, width=1024, height=768, spp=256,
               sigma_a=0.0005, sigma_s=0.0495, continue_prob=0.95, max_depth=64),
    "c5": dict(name="C5 default scene, MIS, 3840x2160 @ 16384 spp", method=2, width=3840, height=2160, spp=16384),
}
METHOD_NAMES = {0: "free-flight", 1: "equi-angular", 2: "mis", 4: "mis-distance"}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md clocks line)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ---- reference arm / cpu baseline: the reference's own CPU code on the host cores ---------------------------------------------
def cpu_reference_rate(cfg, seconds_target, threads):
    """Mpaths/s of the reference's code for cfg's shade method on a bounded sample; returns (rate, kind, sample text)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    w, h, method = cfg["width"], cfg["height"], cfg["method"]
    sa, ss = cfg.get("sigma_a", 0.001), cfg.get("sigma_s", 0.009)
    custom = "continue_prob" in cfg or "max_depth" in cfg
    if oracle_lib.L0.available() and not custom:
        l0 = oracle_lib.L0(); l0.reset_scene(); l0.set_quirks(3)
        run = lambda spp: l0.render(w, h, spp, method, sa, ss, seed=1, nthreads=threads, want_sumsq=False)
        kind, what = "reference", "oracle/_ref/libvpt_l0.so (unmodified reference sources, thread-private erand48)"
    else:  # the reference cannot run this configuration (or was not compiled here): time the FP64 restatement instead
        l1 = oracle_lib.L1()
        run = lambda spp: l1.render(oracle_lib.DEFAULT_SCENE, 3, method, sa, ss, w, h, 1, spp, cp=cfg.get("continue_prob", 0.6), max_depth=cfg.get("max_depth", 0),
                                    nthreads=threads, want_sumsq=False)
        kind, what = "port", "oracle/vpt_oracle.hpp (FP64 CPU restatement)"
    t0 = time.time(); run(1); t1 = time.time() - t0
    spp = max(1, min(256, int(seconds_target / max(t1, 1e-3))))
    t0 = time.time(); run(spp); dt = time.time() - t0
    rate = w * h * spp / dt / 1e6
    return rate, kind, "%dx%d %s at %d spp (%.1f s) via %s" % (w, h, METHOD_NAMES[method], spp, dt, what)


def as_shipped_rate(threads, spp=4):
    """the as-shipped binary (global shared seed, free-flight only, 1024x768): ./ref_rt <spp>, elapsed minus the spp=0 overhead"""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_rt")
    if not os.path.exists(exe):
        return None
    import tempfile
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))

    def run(n):
        with tempfile.TemporaryDirectory() as d:
            t0 = time.time()
            subprocess.run([exe, str(n)], cwd=d, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
            return time.time() - t0
    try:
        base = run(0); t = run(spp)
    except Exception:
        return None
    return {"value": 1024 * 768 * spp / max(t - base, 1e-3) / 1e6, "unit": "Mpaths/s", "cores": threads,
            "sample": "ref_rt %d (free-flight, 1024x768, shared global seed as shipped), %.1f s minus %.1f s fixed overhead" % (spp, t, base)}


def run_reference_arm(args, cfg):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    rates, sample = [], ""
    for i in range(args.warmup + args.steps):
        rate, kind, sample = cpu_reference_rate(cfg, seconds_target=max(3.0, 60.0 / (args.warmup + args.steps)), threads=threads)
        if i >= args.warmup:
            rates.append(rate)
    value = sum(rates) / len(rates)
    paths_per_step = cfg["width"] * cfg["height"] * cfg["spp"]
    line = {"impl": "reference", "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": paths_per_step / (value * 1e6) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (the reference's fixed scene, include/Sphere.cpp:11-22)", "config": config_dict(cfg, args.gpus, "cpu"),
            "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": threads, "kind": kind, "sample": sample + "; ms_per_step extrapolated linearly in spp"},
            "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def config_dict(cfg, n_gpus, where):
    return {"workload": cfg["name"], "method": METHOD_NAMES[cfg["method"]], "width": cfg["width"], "height": cfg["height"], "spp_per_gpu": cfg["spp"],
            "spp_total": cfg["spp"] * n_gpus, "paths_per_step": cfg["width"] * cfg["height"] * cfg["spp"] * n_gpus, "sharding": "samples" if n_gpus > 1 else "none",
            "sigma_a": cfg.get("sigma_a", 0.001), "sigma_s": cfg.get("sigma_s", 0.009), "continue_prob": cfg.get("continue_prob", 0.6),
            "max_depth": cfg.get("max_depth", 0), "precision": "fp32" if where == "gpu" else "fp64", "seed": 1,
            "l2": "flushed between steps (256 MiB write); the kernel reads no HBM input: scene and parameters live in constant memory"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64ref"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference_arm(args, cfg)

    import numpy as np
    import torch
    import torch.distributed as dist
    import minimal_volumetric_path_tracer_b200 as v
    from minimal_volumetric_path_tracer_b200 import build as vbuild, distributed as vdist
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not os.path.exists(v.LIB_PATH):
        if local == 0:
            vbuild.build_all()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    v.load_library()
    dev = torch.device("cuda", local)
    scene = v.default_scene()
    W, H, SPP = cfg["width"], cfg["height"], cfg["spp"]
    extra = {k: cfg[k] for k in ("sigma_a", "sigma_s", "continue_prob", "max_depth") if k in cfg}
    prec = v.PRECISION_FP32 if args.precision == "fp32" else v.PRECISION_FP64_REF
    quirks = 0 if args.precision == "fp32" else v.QUIRKS_REFERENCE
    whole = v.default_params(width=W, height=H, spp=SPP * world, method=cfg["method"], seed=1, device=local, precision=prec, quirks=quirks, **extra)
    mine, _ = vdist.shard_params(whole, "samples", rank, world)       # this rank: samples [rank*SPP, (rank+1)*SPP), SUM output
    hdr = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    paths_per_step = W * H * SPP * world

    def step(timed_events=None):
        flush.zero_()
        if timed_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        v.render_device(mine, scene, hdr.data_ptr(), stream.cuda_stream)
        if timed_events is not None:
            e1.record(stream); timed_events.append((e0, e1))
        if world > 1:
            dist.reduce(hdr, dst=0, op=dist.ReduceOp.SUM)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    sync()
    sampler = ClockSampler(local); sampler.start(); time.sleep(0.15)
    kernel_events = []
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    t_begin.record(stream)
    for _ in range(args.steps):
        step(kernel_events)
    t_end.record(stream)
    sync()
    clocks = sampler.stop()
    ms = torch.tensor([t_begin.elapsed_time(t_end)], dtype=torch.float64, device=dev)
    kernel_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in kernel_events) / len(kernel_events)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(kernel_ms, op=dist.ReduceOp.MAX)
    ms, kernel_ms = float(ms.item()), float(kernel_ms.item())
    value = paths_per_step * args.steps / (ms * 1e-3) / 1e6
    frame_mean = (hdr.double().mean(dim=(0, 1)) / (SPP * world)).tolist() if rank == 0 else None

    # ---- e2e: the reference-facing call with host buffers (copies inside the timed region) ----------------------------------------
    host = np.empty((H, W, 3), dtype=np.float32)
    lib = v.load_library()
    def e2e_step():
        if world == 1:
            rc = lib.vpt_render(C.byref(mine), scene, len(scene), host.ctypes.data_as(C.POINTER(C.c_float)), None)
            assert rc == 0, rc
        else:
            out = vdist.render_sharded(whole, scene, mode="samples", mean=False)
            if out is not None:
                host[...] = out.cpu().numpy()
    e2e_step(); sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    sync()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = paths_per_step * args.steps / float(e2e_s.item()) / 1e6
    h2d = C.sizeof(v.Params) + len(scene) * C.sizeof(v.Sphere)
    d2h = W * H * 3 * 4

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return 0

    # ---- rank 0 only: roofline denominator, secondary figures, CPU baseline ---------------------------------------------------------
    peak_tflops, max_clk = v.measure_fp32_peak(local)
    per_gpu_paths = W * H * SPP
    achieved = per_gpu_paths / (kernel_ms * 1e-3) * FLOP_PER_PATH / 1e12
    roofline = {"bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                "traffic": d2h + 107264,  # bytes per launch: the HDR store + the 107 KB DRAM read of the ncu capture (profiles/r1_smwave_v8_ncu.txt)
               
                "kernel": "render_f32_smwave_kernel<%d>" % cfg["method"] if args.precision == "fp32" else "render_f64_kernel", "kernel_ms_per_launch": kernel_ms,
                "flop_per_path": FLOP_PER_PATH, "peak_source": "measured live: vpt_measure_fp32_peak FFMA chains (MEASURED_PEAKS.json has no FP32 entry; nominal 74.4)",
                "hbm_note": "the kernel is FP32-issue bound, not HBM bound: path state lives in shared memory; algorithmic HBM traffic is the %d-byte HDR store per launch (ncu: dram read 107 KB, DRAM throughput 0.00 %%)" % d2h}
    extras = {}
    if args.config == "c2" and args.precision == "fp32":  # short secondary measurements: the other two methods and the FP64 REF mode
        for name, kw, spp in (("free_flight_mpaths_s", dict(method=0), 256), ("mis_mpaths_s", dict(method=2), 256), ("mis_distance_mpaths_s", dict(method=4), 256),
                              ("fp64_ref_mode_equi_mpaths_s", dict(method=1, precision=v.PRECISION_FP64_REF, quirks=v.QUIRKS_REFERENCE), 64)):
            q = v.default_params(width=W, height=H, spp=spp, seed=1, device=local, **kw)
            st = v.Stats()
            v.render_device(q, scene, hdr.data_ptr(), stream.cuda_stream, st)
            v.render_device(q, scene, hdr.data_ptr(), stream.cuda_stream, st)
            extras[name] = st.paths / st.kernel_ms / 1e3
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        try:
            rate, kind, sample = cpu_reference_rate(cfg, seconds_target=15.0, threads=threads)
            cpu = {"value": rate, "unit": "Mpaths/s", "cores": threads, "kind": kind, "sample": sample, "as_shipped_binary": as_shipped_rate(threads)}
        except Exception as e:  # the bench line must still print
            cpu = {"value": None, "unit": "Mpaths/s", "cores": threads, "kind": "unavailable", "sample": repr(e)}
    line = {"metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "f64",
            "data": "synthetic (the reference's fixed scene, include/Sphere.cpp:11-22; Philox seed 1)", "config": config_dict(cfg, world, "gpu"),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "vpt_render() host buffers" if world == 1 else "distributed.render_sharded() + D2H on rank 0"},
            "clocks": clocks, "gpu_launches": args.steps, "frame_mean_rgb": frame_mean, "extras": extras}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
