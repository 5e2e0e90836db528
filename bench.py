#!/usr/bin/env python
"""bench.py -- Mpaths/s of the per-pixel radiance loop (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c3|c1|c2|c4|c5]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (config.workload): ONE fixed frame, by default BASELINE.json configs[2] = the reference's default scene, "MIS" shade
method, 1920x1080 at 4096 spp (8.49e9 camera paths; `--config c5` is configs[4] as written, 3840x2160 at 16384 spp).  One
"step" is one full render of that frame.  With N GPUs the SAME frame is sharded by samples (rank r renders samples
[r*spp/N, (r+1)*spp/N) of every pixel -- the split of the loop src/rt.cpp:767-798) and ONE NCCL reduce of the fp32 HDR
buffers inside the timed region combines them: strong scaling, `spp_total` is constant.  On rank 0 the reduced frame of the
last timed step is compared with the reference's own render of this configuration (tests/golden/image_robust_m2_c3.npz,
16x16-block means, whole-image z < 4.5), so every point of the scaling curve is also a correctness run.

`value` counts the frame's camera paths over the max-over-ranks device time.  `e2e` is the same metric through the
reference-facing call with HOST buffers (vpt_render() at N = 1; distributed.render_sharded() + the D2H of the reduced frame
at N > 1), copies inside the timed region.  `roofline` is FP32 CUDA-core throughput: algorithmic FLOP per camera path
(1700 canonical for C1/C2/C3/C5, W = 18*S*T + 460*E from vpt_stats for C4) against the FFMA peak measured live.
`extras` carries the other BASELINE.json configurations measured with the same build (N = 1; at N > 1 one frame of config 5 as written,
3840x2160 at 16384 spp sharded over the N GPUs).  `cpu_baseline` /
`--impl reference` time the reference's own CPU code (oracle/_ref, compiled from the unmodified sources) on the host cores."""
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

FLOP_PER_PATH = 1700.0  # 18*S*T + 460*E at S=10 spheres, T=5.5 scans, E=1.5 events (canonical for C1/C2/C3/C5)
N_SPHERES = 10

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


def flop_per_path(scans_per_path, events_per_path):
    """W = 18 * S * T + 460 * E"""
    return 18.0 * N_SPHERES * scans_per_path + 460.0 * events_per_path


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
            "ms_per_step": paths_per_step / (value * 1e6) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (the reference's fixed scene, include/Sphere.cpp:11-22)", "config": config_dict(cfg, args.gpus, "cpu"),
            "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": threads, "kind": kind,
                             "sample": sample + "; each step is such a sample, ms_per_step is the whole frame extrapolated linearly in spp (the cost is exactly linear)"},
            "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def config_dict(cfg, n_gpus, where):
    return {"workload": cfg["name"], "method": METHOD_NAMES[cfg["method"]], "width": cfg["width"], "height": cfg["height"],
            "spp_total": cfg["spp"], "spp_per_gpu": cfg["spp"] / n_gpus, "paths_per_step": cfg["width"] * cfg["height"] * cfg["spp"],
            "sharding": "samples (one fixed frame, one NCCL reduce of the HDR buffers per step)" if n_gpus > 1 else "none",
            "sigma_a": cfg.get("sigma_a", 0.001), "sigma_s": cfg.get("sigma_s", 0.009), "continue_prob": cfg.get("continue_prob", 0.6),
            "max_depth": cfg.get("max_depth", 0), "precision": "fp32" if where == "gpu" else "fp64", "seed": 1,
            "l2": "flushed between steps (256 MiB write); the kernel reads no HBM input: scene and parameters live in constant memory"}


def golden_check(cfg, hdr_sum_np, spp):
    """rank 0: the reduced frame against the reference's own render of C3 (tests/golden/image_robust_m2_c3.npz: 16x16-block means and
    their variances at 4096 spp).  The C5 frame is the same view at twice the resolution: 32x32 blocks cover the same image regions."""
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "image_robust_m2_c3.npz")
    if cfg["method"] != 2 or cfg["width"] * 9 != cfg["height"] * 16 or cfg["width"] % 1920 or "sigma_a" in cfg or not os.path.exists(path):
        return None
    g = np.load(path)
    block = 16 * (cfg["width"] // 1920)
    h, w, _ = hdr_sum_np.shape
    bm = (hdr_sum_np.astype(np.float64) / spp)[:h // block * block, :w // block * block].reshape(h // block, block, w // block, block, 3).mean(axis=(1, 3))
    ref, var = g["block_mean"].astype(np.float64), g["block_var"].astype(np.float64)
    if bm.shape != ref.shape:
        return None
    # variance of the difference of the two whole-image means: the reference's block variances at its 4096 spp, ours scaled to this render's spp
    sigma = np.sqrt((var * (1.0 + float(g["spp"]) / spp)).sum(axis=(0, 1))) / (var.shape[0] * var.shape[1])
    z = (bm.mean(axis=(0, 1)) - ref.mean(axis=(0, 1))) / sigma
    rel = np.abs(bm - ref) / np.maximum(ref, 1e-9)
    return {"golden": "tests/golden/image_robust_m2_c3.npz (the unmodified reference's 1920x1080 @ 4096 spp render, robust hooks)",
            "global_z_rgb": [float(x) for x in z], "ok": bool(np.all(np.abs(z) < 4.5)), "median_block_rel_diff": float(np.median(rel)),
            "blocks": list(ref.shape[:2]), "block_pixels": block}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64ref"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
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
    whole = v.default_params(width=W, height=H, spp=SPP, method=cfg["method"], seed=1, device=local, precision=prec, quirks=quirks, **extra)
    mine, has_work = vdist.shard_params(whole, "samples", rank, world)       # this rank: samples [rank*SPP/N, (rank+1)*SPP/N) of every pixel, SUM output
    if not has_work:
        raise SystemExit("more ranks than samples")
    hdr = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    paths_per_step = W * H * SPP                                       # the whole frame, whatever N is

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
    visible = [t.strip() for t in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if t.strip()]  # nvidia-smi does not honour the mask: translate
    sampler = ClockSampler(visible[local] if local < len(visible) else local); sampler.start(); time.sleep(0.15)
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
    frame = hdr.cpu().numpy() if rank == 0 else None                 # the reduced frame of the last timed step (SUM over all samples)

    # ---- e2e: the reference-facing call with host buffers (copies inside the timed region) ----------------------------------------
    pinned = v.PinnedFrame(H, W)                                     # page-locked host frame (the contract's "pinned host memory")
    host = pinned.array
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

    # ---- N > 1: BASELINE.json config 5 as written (3840x2160 @ 16384 spp sharded over the N GPUs, one NCCL reduce), one warm-up + one timed frame
    c5_sharded = None
    if world > 1 and args.config != "c5" and args.precision == "fp32" and not args.no_extras:
        c5 = CONFIGS["c5"]
        whole5 = v.default_params(width=c5["width"], height=c5["height"], spp=c5["spp"], method=c5["method"], seed=1, device=local)
        mine5, _ = vdist.shard_params(whole5, "samples", rank, world)
        hdr5 = torch.zeros((c5["height"], c5["width"], 3), dtype=torch.float32, device=dev)
        ev5 = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
        for timed in (False, True):
            sync()
            if timed:
                ev5[0].record(stream)
            v.render_device(mine5, scene, hdr5.data_ptr(), stream.cuda_stream)
            dist.reduce(hdr5, dst=0, op=dist.ReduceOp.SUM)
            if timed:
                ev5[1].record(stream)
            sync()
        ms5 = torch.tensor([ev5[0].elapsed_time(ev5[1])], dtype=torch.float64, device=dev)
        dist.all_reduce(ms5, op=dist.ReduceOp.MAX)
        if rank == 0:
            paths5 = c5["width"] * c5["height"] * c5["spp"]
            c5_sharded = {"workload": c5["name"] + ", sharded by samples over %d GPUs, one NCCL reduce" % world, "mpaths_s": paths5 / float(ms5.item()) / 1e3,
                          "ms_per_frame": float(ms5.item()), "frame_check": golden_check(c5, hdr5.cpu().numpy(), c5["spp"])}
        del hdr5

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return 0

    # ---- rank 0 only: correctness of the frame, roofline, the other configurations, CPU baseline ----------------------------------
    check = golden_check(cfg, frame, SPP) if args.precision == "fp32" else None
    frame_mean = (frame.astype(np.float64).mean(axis=(0, 1)) / SPP).tolist()
    peak_tflops, max_clk = v.measure_fp32_peak(local)
    st = v.Stats()
    probe = whole.copy(spp=min(SPP, 64), sample_begin=0, sample_end=0, output=v.OUTPUT_SUM)   # scans and events per path of this workload (vpt_stats)
    v.render_device(probe, scene, hdr.data_ptr(), stream.cuda_stream, st)
    scans_pp, events_pp = st.scene_scans / st.paths, st.events / st.paths
    fpp_formula = flop_per_path(scans_pp, events_pp)
    fpp = fpp_formula if args.config == "c4" else FLOP_PER_PATH
    per_gpu_paths = W * H * (mine.sample_end - mine.sample_begin)
    achieved = per_gpu_paths / (kernel_ms * 1e-3) * fpp / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r2_dram_traffic.json")           # dram__bytes_read.sum + dram__bytes_write.sum of ncu --set full captures
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("%dx%d" % (W, H), {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                "traffic": traffic, "algorithmic_hbm_bytes": d2h,
                "kernel": ("render_f32_smwave_kernel<%d, %d, 1>" % (cfg["method"], 6 if (mine.sample_end - mine.sample_begin) < 96 else 4 if (mine.sample_end - mine.sample_begin) < 384 else 2)) if args.precision == "fp32" else "render_f64_smwave_kernel", "kernel_ms_per_launch": kernel_ms,
                "flop_per_path": fpp, "flop_per_path_formula": fpp_formula, "scans_per_path": scans_pp, "events_per_path": events_pp,
                "frac_formula": per_gpu_paths / (kernel_ms * 1e-3) * fpp_formula / 1e12 / peak_tflops,
                "peak_source": "measured live: vpt_measure_fp32_peak FFMA chains (MEASURED_PEAKS.json has no FP32 entry; nominal 74.4)",
                "hbm_note": "FP32-issue bound, not HBM bound: path state lives in shared memory; the algorithmic HBM traffic is the %d-byte HDR store per launch "
                            "(n_pixels * 12); `traffic` = DRAM bytes of the ncu --set full capture of this frame size (profiles/r2_dram_traffic.json), null if not captured" % d2h}
    extras = {}
    if world == 1 and args.precision == "fp32" and not args.no_extras:
        # the other BASELINE.json configurations with the same build: device-timed value, e2e through vpt_render(), roofline fraction
        for name in ("c1", "c2", "c4", "c5"):
            if name == args.config:
                continue
            c = CONFIGS[name]
            kw = {k: c[k] for k in ("sigma_a", "sigma_s", "continue_prob", "max_depth") if k in c}
            q = v.default_params(width=c["width"], height=c["height"], spp=c["spp"], method=c["method"], seed=1, device=local, **kw)
            buf = torch.empty((c["height"], c["width"], 3), dtype=torch.float32, device=dev)
            s = v.Stats()
            reps = 1 if name == "c5" else 3
            best = None
            for _ in range(reps + (0 if name == "c5" else 1)):
                flush.zero_()
                v.render_device(q, scene, buf.data_ptr(), stream.cuda_stream, s)
                best = s.kernel_ms if best is None else min(best, s.kernel_ms)
            pf = v.PinnedFrame(c["height"], c["width"]); hb = pf.array
            if name != "c5":  # warm-up of the e2e path (an 18 s frame needs none)
                lib.vpt_render(C.byref(q), scene, len(scene), hb.ctypes.data_as(C.POINTER(C.c_float)), None)
            t0 = time.perf_counter()
            for _ in range(reps):
                rc = lib.vpt_render(C.byref(q), scene, len(scene), hb.ctypes.data_as(C.POINTER(C.c_float)), None)
                assert rc == 0, rc
            e2e_c = s.paths * reps / (time.perf_counter() - t0) / 1e6
            f = flop_per_path(s.scene_scans / s.paths, s.events / s.paths)
            rate = s.paths / best / 1e3
            use = f if name == "c4" else FLOP_PER_PATH
            extras[name] = {"workload": c["name"], "mpaths_s": rate, "e2e_mpaths_s": e2e_c, "kernel_ms": best, "scans_per_path": s.scene_scans / s.paths,
                            "events_per_path": s.events / s.paths, "flop_per_path": use, "roofline_frac": rate * 1e6 * use / 1e12 / peak_tflops,
                            "mvertices_s": s.events / best / 1e3}
            del buf, hb; pf.close()
        for name, kw, spp in (("free_flight_mpaths_s", dict(method=0), 256), ("equiangular_mpaths_s", dict(method=1), 256), ("mis_distance_mpaths_s", dict(method=4), 256),
                              ("fp64_ref_mode_equi_mpaths_s", dict(method=1, precision=v.PRECISION_FP64_REF, quirks=v.QUIRKS_REFERENCE), 64)):
            q = v.default_params(width=1024, height=768, spp=spp, seed=1, device=local, **kw)
            s = v.Stats()
            v.render_device(q, scene, hdr.data_ptr(), stream.cuda_stream, s)
            v.render_device(q, scene, hdr.data_ptr(), stream.cuda_stream, s)
            extras[name] = s.paths / s.kernel_ms / 1e3
    if c5_sharded is not None:
        extras["c5_sharded"] = c5_sharded
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        try:
            rate, kind, sample = cpu_reference_rate(cfg, seconds_target=15.0, threads=threads)
            cpu = {"value": rate, "unit": "Mpaths/s", "cores": threads, "kind": kind, "sample": sample, "as_shipped_binary": as_shipped_rate(threads)}
        except Exception as e:  # the bench line must still print
            cpu = {"value": None, "unit": "Mpaths/s", "cores": threads, "kind": "unavailable", "sample": repr(e)}
    line = {"metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "f64",
            "data": "synthetic (the reference's fixed scene, include/Sphere.cpp:11-22; Philox seed 1)", "config": config_dict(cfg, world, "gpu"),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "vpt_render() host buffers" if world == 1 else "distributed.render_sharded() + D2H of the reduced frame on rank 0"},
            "clocks": clocks, "gpu_launches": args.steps, "frame_mean_rgb": frame_mean, "frame_check": check, "extras": extras}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
