#!/usr/bin/env python
"""Benchmark of the TAG scoring hot path (BASELINE.json metric: videos/sec scored, encode + AC + TC).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path over one batch of synthetic videos: K1 feature fuse -> K2 encoder
-> K4 AC (distance to precomputed per-action centroids) + TC. Workload at every N = BASELINE config 2
per GPU (5000 generated videos x 64 frames, clip_len 32 / stride 8 -> 25000 windows): weak scaling,
videos are sharded across ranks with no data-path collective; the centroid build that
precedes the timed region is where the one NCCL all-reduce happens.

  value  : whole-job videos/s with inputs already resident in HBM (CUDA events, max over ranks)
  e2e    : same metric through the public streaming call (TagScorer.score_stream) with HOST (pinned)
           input buffers: every step's H2D of all input arrays (copy stream, prefetched across step
           boundaries) + scoring + D2H of the per-video results, all inside the timed region
  roofline: dominant kernel = the dilated-conv tensor-core GEMM; achieved = algorithmic FLOPs per launch
           / mean launch duration measured with CUDA events on the launching stream during the timed
           steps; peak = MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)
  cpu_baseline: the oracle port of the reference's CPU path (torch CPU ops, all host cores) on a bounded
           sample of the same workload (N=1, rank 0 only)
`--impl reference` times that same CPU path as the reference arm (the reference is pure Python/torch and
cannot be compiled into oracle/_ref).
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "videos/sec scored (encode+AC+TC)"
UNIT = "videos/s"
CLIP_LEN, STRIDE = 32, 8
GFLOP_PER_WINDOW = 2.0248      # SURVEY.md §8(d), T=32, M=5, reference layer shapes


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--videos", type=int, default=5000, help="generated videos per GPU (config 2: 5000)")
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--precision", default="fp16_tc", choices=["fp16_tc", "fp32"])
    ap.add_argument("--max-windows", type=int, default=0, help="windows per internal pass (0 = auto)")
    ap.add_argument("--cpu-sample-videos", type=int, default=768)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)),
                "source": "MEASURED_PEAKS.json (bf16_tflops_sustained)"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_cpus(index):
    """Restrict this process to the CPUs NVML reports as local to GPU `index` (NUMA locality of the pinned staging
    buffers of the end-to-end leg). Returns a short description, or None when NVML has no answer."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1} & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} CPUs local to GPU {phys}"
    except Exception:
        pass
    return None


def h2d_bandwidth(torch, dev, nbytes=1 << 30):
    """Measured pinned host -> device copy rate on this box (explains the e2e/value gap)."""
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(dev)
    return 3 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9


# ------------------------------------------------------------------------------------------------
def cpu_reference_pass(pkg, n_videos, frames, seed, sd, dims_raw, dims_diff, ostats, centroids):
    """The reference's CPU compute path (oracle port): window features -> encoder -> AC + TC for n_videos
    videos held in memory. Returns (seconds, n_videos)."""
    import torch
    O = importlib.import_module("oracle.tag_oracle")
    vb = pkg.make_videos(n_videos, frames, seed=seed)
    vids = [vb.video(v) for v in range(n_videos)]
    label_dict = {c: i for i, c in enumerate(pkg.ACTION_CLASSES)}
    t0 = time.perf_counter()
    with torch.no_grad():
        ac, tc, _ = O.score_videos(vids, vb.names, [vb.cls_name(v) for v in range(n_videos)], sd, dims_raw, dims_diff,
                                   ostats, centroids, label_dict, clip_len=CLIP_LEN, stride=STRIDE, batch=32)
    dt = time.perf_counter() - t0
    assert len(ac) == n_videos and len(tc) == n_videos
    return dt, n_videos


def cpu_setup(pkg, frames):
    import torch
    O = importlib.import_module("oracle.tag_oracle")
    torch.set_num_threads(os.cpu_count() or 1)
    dims_raw, dims_diff = pkg.dims_maps(False)
    sd = pkg.make_state_dict(dims_raw, dims_diff, seed=0)
    real = pkg.make_videos(20, frames, seed=1337 + 3)
    ostats = O.compute_stats([real.video(v) for v in range(real.n_videos)])
    g = torch.Generator().manual_seed(0)
    centroids = torch.nn.functional.normalize(torch.randn(10, 256, generator=g), dim=-1)
    return dims_raw, dims_diff, sd, ostats, centroids


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the path (oracle port; the Python
    reference cannot travel to the GPU box) on the host cores, bounded sample per step."""
    if rank != 0:
        return
    import torch
    pkg = importlib.import_module("video-gen-evals_b200")
    dims_raw, dims_diff, sd, ostats, centroids = cpu_setup(pkg, args.frames)
    n = max(32, min(256, 4800 // max(1, args.steps)))      # ~50 videos/s on these hosts: the whole run stays under ~2 minutes
    for w in range(min(args.warmup, 1)):
        cpu_reference_pass(pkg, 8, args.frames, 5, sd, dims_raw, dims_diff, ostats, centroids)
    total, vids = 0.0, 0
    for k in range(args.steps):
        dt, nv = cpu_reference_pass(pkg, n, args.frames, 100 + k, sd, dims_raw, dims_diff, ostats, centroids)
        total += dt; vids += nv
    value = vids / total
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * total / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config2 sample: {n} videos x {args.frames} frames per step, clip {CLIP_LEN}/stride {STRIDE}, "
                                   "reference CPU path (window features + encoder + AC + TC) on in-memory tensors"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n} videos ({n * 5} windows) per step x {args.steps} steps"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("video-gen-evals_b200")
    lib = pkg.load_library()          # raises if libtag_b200.so is missing: no fallback
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA (B200) device")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_cpus(local_rank)      # pinned host buffers are then first-touched on the GPU's own NUMA node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    dims_raw, dims_diff = pkg.dims_maps(False)
    sd = pkg.make_state_dict(dims_raw, dims_diff, seed=0)
    n_windows = args.videos * ((args.frames - CLIP_LEN) // STRIDE + 1 if args.frames >= CLIP_LEN else 1)
    if args.max_windows:
        max_windows = args.max_windows
    elif args.precision == "fp16_tc":
        # windows per pass: a whole number of waves of the persistent tensor-core GEMM (74 CTA pairs x 256 rows = 592
        # windows of 32 frames per wave), about 12.5k windows (8 GB of workspace) per pass
        n_pass = -(-n_windows // 12800)
        max_windows = min(n_windows, -(-(-(-n_windows // n_pass)) // 592) * 592)
    else:
        max_windows = min(n_windows, 2048)
    model = pkg.HumanActionScorer(dims_raw, dims_diff, precision=args.precision, max_windows=max_windows)
    model.load_state_dict(sd)
    model.to(dev).eval()

    # --- setup (untimed): stats + centroids from a synthetic "real" set, sharded, one all-reduce
    real_all = 200
    lo, hi = pkg.shard_range(real_all, rank, world)
    real = pkg.make_videos(real_all, args.frames, seed=1337 + 3, device=dev)
    stats = pkg.compute_stats_from_videos(real, dims_raw, dims_diff, dev)
    scorer = pkg.TagScorer(model, stats, CLIP_LEN, STRIDE, dev)
    centroids, counts = scorer.build_centroids(scorer.to_device(real.select(range(lo, hi))), 10)
    assert int(counts.sum().item()) == real_all * ((args.frames - CLIP_LEN) // STRIDE + 1)
    del real

    # --- the per-GPU batch of generated videos (resident in HBM)
    gen = pkg.make_videos(args.videos, args.frames, seed=1337 + 2 + 1000 * rank, device=dev)
    dv = scorer.to_device(gen)
    in_bytes = gen.input_bytes()
    util_h = pkg.scoring.util_handle(dev)

    def launches():
        return int(model.launch_count()) + int(lib.tag_launch_count(util_h))

    def step():
        return scorer.score(dv, centroids)

    for _ in range(max(args.warmup, 3)):
        ac, tc = step()
    torch.cuda.synchronize(dev)
    assert bool(torch.isfinite(ac).all()) and bool(torch.isfinite(tc).all())
    assert int(scorer.last_flags.item()) == 0

    # --- timed region: value (device-resident inputs)
    h = model.handle(dev, CLIP_LEN)
    pkg._lib.check(h, lib.tag_set_profiling(h, 1), "tag_set_profiling")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        ac, tc = step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    l1 = launches()
    prof = (__import__("ctypes").c_double * 12)()
    pkg._lib.check(h, lib.tag_get_profile(h, prof), "tag_get_profile")
    pkg._lib.check(h, lib.tag_set_profiling(h, 0), "tag_set_profiling")
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * args.videos * args.steps / (ms_max / 1000.0)

    # --- e2e: host (pinned) inputs -> H2D -> score -> D2H, every step, through the streaming public call
    gen_host = gen.to("cpu").pin()
    for _ in scorer.score_stream((gen_host for _ in range(2)), centroids):
        pass
    e2e_steps = args.steps
    barrier()
    ev0.record()
    for hac, htc in scorer.score_stream((gen_host for _ in range(e2e_steps)), centroids):
        pass
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * args.videos * e2e_steps / (float(t.item()) / 1000.0)
    h2d_gbs = h2d_bandwidth(torch, dev)
    for tid in os.listdir("/proc/self/task"):   # the CPU baseline below uses every host core, on every thread
        try:
            os.sched_setaffinity(int(tid), all_cpus)
        except OSError:
            pass
    assert float((hac - ac.cpu()).abs().max()) < 1e-5

    if rank == 0:
        pk = peaks()
        other_ms, _, other_n, conv_ms, conv_flops, conv_n, og_ms, og_flops, og_n, k1_ms, k1_bytes, k1_n = [float(x) for x in prof]
        achieved = (conv_flops / conv_n) / (conv_ms / conv_n * 1e-3) / 1e12 if conv_n > 0 and conv_ms > 0 else None
        tc_mode = args.precision == "fp16_tc"
        roof_peak = pk["tflops"] if tc_mode else 75.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "conv_gemm_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(args.precision)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if tc_mode else "f32", "data": "synthetic",
            "config": {"workload": f"config2 (TAG-Bench scale) per GPU: {args.videos} videos x {args.frames} frames, clip {CLIP_LEN} / "
                                   f"stride {STRIDE} -> {n_windows} windows, M=5 D=2596, precomputed centroids [10,256]",
                       "precision": "fp16 operands, fp32 accumulate/norms (tcgen05)" if tc_mode else "fp32 CUDA cores",
                       "weights": "random-init (seeded), reference state_dict layout", "windows_per_pass": max_windows,
                       "l2": f"inputs resident in HBM ({in_bytes / 1e9:.2f} GB per GPU) exceed the 126 MB L2; no flush needed",
                       "parallelism": f"videos sharded over {world} GPU(s), no data-path collective"},
            "roofline": {"bound": "tensor", "kernel": "k_gemm_tc (dilated conv, 5 taps)" if tc_mode else "k_gemm_f32 (dilated conv, 5 taps)",
                         "achieved": achieved, "peak": roof_peak, "unit": "TFLOP/s",
                         "frac": (achieved / roof_peak) if achieved else None, "traffic": traffic,
                         "peak_source": pk["source"] if tc_mode else "nominal fp32 CUDA-core peak (~75 TFLOP/s), fp32 mode only",
                         "launches_sampled": int(conv_n), "mean_launch_ms": conv_ms / conv_n if conv_n else None,
                         "share_of_step": {"conv_gemm_ms": conv_ms, "other_gemm_ms": og_ms, "feature_fuse_ms": k1_ms,
                                           "other_kernels_ms": other_ms},
                         "feature_fuse_hbm": {"bound": "hbm", "achieved": (k1_bytes / (k1_ms * 1e-3) / 1e9) if k1_ms > 0 else None,
                                              "peak": pk["hbm_gbs"], "unit": "GB/s",
                                              "frac": (k1_bytes / (k1_ms * 1e-3) / 1e9 / pk["hbm_gbs"]) if k1_ms > 0 else None},
                         "whole_encoder_tflops": value / world * (n_windows / args.videos) * GFLOP_PER_WINDOW / 1e3},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": int(2 * args.videos * 4),
                    "steps": e2e_steps, "h2d_gbs_measured": h2d_gbs, "host_affinity": numa, "call": "TagScorer.score_stream: every step's 5 input arrays copied from pinned host memory "
                    "(one block per encoder pass, short ramp-up blocks at the start of the stream, prefetched on a copy stream), per-video AC/TC read back to the host every step"},
            "gpu_launches": l1 - l0,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            dr, dd, sdc, ostats, cen_cpu = cpu_setup(pkg, args.frames)
            cpu_reference_pass(pkg, 8, args.frames, 5, sdc, dr, dd, ostats, cen_cpu)          # warm-up
            dt, nv = cpu_reference_pass(pkg, args.cpu_sample_videos, args.frames, 7, sdc, dr, dd, ostats, cen_cpu)
            line["cpu_baseline"] = {"value": nv / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{nv} videos x {args.frames} frames ({nv * 5} windows) of the same workload, "
                                              f"{dt:.1f} s, oracle port of the reference CPU path (torch CPU ops)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
