#!/usr/bin/env python
"""Benchmark of the TAG scoring hot path (BASELINE.json metric: videos/sec scored, encode + AC + TC).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path over one batch of synthetic videos: K1 feature fuse -> K2 encoder
-> K4 AC (distance to precomputed per-action centroids) + TC. Workload at every N = BASELINE config 2
per GPU (5000 generated videos x 64 frames, clip_len 32 / stride 8 -> 25000 windows): weak scaling,
videos are sharded across ranks with no data-path collective.

  value  : whole-job videos/s with inputs already resident in HBM (CUDA events, max over ranks)
  e2e    : same metric through the public streaming call (TagScorer.score_stream) with HOST (pinned)
           input buffers: every step's H2D of all input arrays (copy stream, prefetched across step
           boundaries) + scoring + D2H of the per-video results, all inside the timed region
  roofline: dominant kernel = the dilated-conv blocks on tcgen05 (k_tcn_block: a fused TemporalConvBlock, 2 convs per
           launch; k_gemm_tc where a shape takes the two-kernel path); achieved = algorithmic FLOPs of those launches
           / their total duration measured with CUDA events on the launching stream during the timed
           steps; peak = MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)
  hbm_kernels: the bandwidth-bound kernels (K1, merge-fusion, attention, build-tokens, finalize+TC, K3, K4):
           algorithmic bytes / CUDA-event time against MEASURED_PEAKS.json hbm_gbs
  configs: the other named shapes of BASELINE.json at this N (per GPU, weak scaling): config 1 (64 x 32
           frames), config 3 (centroid build, the 8-GPU share of 100k clips per rank, WITH the NCCL all-reduce
           of the [10,257] sums inside the timed region), config 4 (512 x 256 frames, M = 7), config 5
           (4096 clips: encoder + TCL similarity matrix)
  cpu_baseline / --impl reference: the UNMODIFIED reference (oracle/_ref: its own WindowDataset +
           DataLoader, HumanActionScorer on the CPU, extract_window_features, AC/TC scorers) on the box's
           host cores over a bounded sample of the same workload; the oracle port only if oracle/_ref is absent
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
    ap.add_argument("--cpu-sample-videos", type=int, default=0, help="reference arm: videos per step (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs 1/3/4/5 and standalone K3/K4 timings")
    ap.add_argument("--cpu-workers", type=int, default=4, help="DataLoader workers of the reference arm (eval.py:414 uses 4)")
    return ap.parse_args()


def workload_config(args, world):
    """`config` of the JSON line — a function of the workload only, identical in both arms."""
    wpv = (args.frames - CLIP_LEN) // STRIDE + 1 if args.frames >= CLIP_LEN else 1
    in_bytes = args.videos * args.frames * 5480
    return {"workload": f"config2 (TAG-Bench scale) per GPU: {args.videos} videos x {args.frames} frames, clip {CLIP_LEN} / "
                        f"stride {STRIDE} -> {args.videos * wpv} windows, M=5 D=2596, precomputed centroids [10,256]",
            "weights": "random-init (seeded), reference state_dict layout",
            "l2": f"inputs of a step ({in_bytes / 1e9:.2f} GB per GPU) exceed the 126 MB L2; no flush needed",
            "parallelism": f"videos sharded over {world} GPU(s), no data-path collective"}


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
class CpuReference:
    """The reference's CPU path on a bounded sample of the workload. With oracle/_ref present (placed there by
    __graft_entry__.build()): the UNMODIFIED reference — files in the reference's on-disk layout (written to /dev/shm,
    untimed), then per step eval.py:403-437 as is: WindowDataset + DataLoader (batch 32, `workers` workers),
    extract_window_features with the reference HumanActionScorer on the CPU, compute_action_consistency_scores,
    compute_temporal_coherence_scores. Otherwise the oracle port on in-memory tensors."""

    def __init__(self, pkg, frames, n_videos, workers):
        import torch
        self.pkg, self.frames, self.n, self.workers = pkg, frames, n_videos, workers
        self.O = importlib.import_module("oracle.tag_oracle")
        self.RR = importlib.import_module("oracle.ref_runner")
        self.ref = self.RR.load_ref()
        torch.set_num_threads(os.cpu_count() or 1)
        self.dims_raw, self.dims_diff = pkg.dims_maps(False)
        self.sd = pkg.make_state_dict(self.dims_raw, self.dims_diff, seed=0)
        real = pkg.make_videos(20, frames, seed=1337 + 3)
        self.ostats = self.O.compute_stats([real.video(v) for v in range(real.n_videos)])
        g = torch.Generator().manual_seed(0)
        self.centroids = torch.nn.functional.normalize(torch.randn(10, 256, generator=g), dim=-1)
        self.label_dict = {c: i for i, c in enumerate(pkg.ACTION_CLASSES)}
        self.vb = pkg.make_videos(n_videos, frames, seed=7)
        self.tmp = None
        if self.ref is not None:
            self.kind = "reference"
            self.tmp = self.RR.scratch_dir("tag_bench_ref_")
            self.gen_dir, self.kp_dir = os.path.join(self.tmp, "generated_meshes"), os.path.join(self.tmp, "generated_kps")
            self.RR.write_set(self.vb, self.gen_dir, self.kp_dir, generated=True)
            self.stats = self.RR.stats_object(self.ref, self.ostats)
            self.model = self.RR.reference_model(self.ref, self.sd, self.dims_raw, self.dims_diff)
        else:
            self.kind = "port"
        self.cores = torch.get_num_threads()

    def step(self):
        """-> seconds for one pass over the sample"""
        import torch
        if self.ref is not None:
            loader, _ = self.RR.generated_loader(self.ref, self.gen_dir, self.kp_dir, self.stats, CLIP_LEN, STRIDE, batch_size=32,
                                                 workers=self.workers)
            dt, ac, tc, _ = self.RR.reference_scoring_pass(self.ref, self.model, loader, self.centroids, self.label_dict, device="cpu")
        else:
            vids = [self.vb.video(v) for v in range(self.n)]
            t0 = time.perf_counter()
            with torch.no_grad():
                ac, tc, _ = self.O.score_videos(vids, self.vb.names, [self.vb.cls_name(v) for v in range(self.n)], self.sd,
                                                self.dims_raw, self.dims_diff, self.ostats, self.centroids, self.label_dict,
                                                clip_len=CLIP_LEN, stride=STRIDE, batch=32)
            dt = time.perf_counter() - t0
        assert len(ac) == self.n and len(tc) == self.n
        return dt

    def describe(self):
        wpv = (self.frames - CLIP_LEN) // STRIDE + 1 if self.frames >= CLIP_LEN else 1
        how = (f"unmodified reference (oracle/_ref: WindowDataset + DataLoader batch 32 / {self.workers} workers reading .npz/.npy from "
               f"/dev/shm, HumanActionScorer on CPU, extract_window_features, AC + TC scorers)") if self.ref is not None else \
              "oracle port of the reference CPU path (torch CPU ops on in-memory tensors; oracle/_ref absent)"
        return f"{self.n} videos x {self.frames} frames ({self.n * wpv} windows) of the same workload per step; {how}"

    def close(self):
        if self.tmp:
            import shutil
            shutil.rmtree(self.tmp, ignore_errors=True)


def run_reference(args, rank, world):
    """Reference arm: rank 0 alone times the reference's own CPU implementation on the host cores."""
    if rank != 0:
        return
    import torch
    pkg = importlib.import_module("video-gen-evals_b200")
    n = args.cpu_sample_videos or max(32, min(256, 4800 // max(1, args.steps)))   # ~50 videos/s on these hosts: under ~2 minutes
    cpu = CpuReference(pkg, args.frames, n, args.cpu_workers)
    try:
        for _ in range(min(args.warmup, 1)):
            cpu.step()
        total = sum(cpu.step() for _ in range(args.steps))
    finally:
        cpu.close()
    value = n * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * total / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind,
                             "sample": cpu.describe() + f"; {args.steps} timed steps"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(args):
    """cpu_baseline of the main arm: the reference arm in a fresh process (no CUDA context, DataLoader workers fork cleanly),
    one warm-up + one timed step over ~15-25 s of CPU work."""
    n = args.cpu_sample_videos or 768
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "1", "--frames", str(args.frames),
           "--videos", str(args.videos), "--cpu-sample-videos", str(n), "--cpu-workers", str(args.cpu_workers)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    for ln in reversed(r.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)["cpu_baseline"]
    return {"value": None, "unit": UNIT, "cores": None, "kind": "unavailable", "sample": (r.stderr or r.stdout)[-300:]}


# ------------------------------------------------------------------------------------------------
def timed(torch, dist, dev, world, fn, steps, warmup=2):
    """fn() `steps` times between barrier+sync, CUDA events, max over ranks -> ms per step"""
    for _ in range(warmup):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps


def other_configs(args, pkg, torch, dist, dev, rank, world, model, scorer, stats, centroids, pk):
    """BASELINE configs 1, 3, 4, 5 at this N (per-GPU work fixed: weak scaling) + K3 / K4 alone on HBM-sized inputs."""
    out = {}
    steps = max(2, min(args.steps, 5))
    tf_peak = pk["tflops"]
    dims_raw, dims_diff = pkg.dims_maps(False)

    def entry(ms, units, unit_name, gflop_per_unit, note):
        tfl = world * units * gflop_per_unit / ms          # GFLOP / ms = TFLOP/s, whole job
        return {"ms": ms, "value": world * units / (ms * 1e-3), "unit": unit_name + "/s", "tflops": tfl,
                "frac": tfl / world / tf_peak, "per_gpu": units, "note": note}

    # ---- config 1: 64 videos x 32 frames (one window each)
    v1 = pkg.make_videos(64, 32, seed=1337 + 1 + 1000 * rank, device=dev)
    dv1 = scorer.to_device(v1)
    ms = timed(torch, dist, dev, world, lambda: scorer.score(dv1, centroids), steps * 4)
    out["config1"] = entry(ms, 64, "videos", GFLOP_PER_WINDOW, "64 videos x 32 frames: one 64-window pass (launch-latency bound: ~240 launches)")
    del v1, dv1

    # ---- config 3: centroid build, 12,500 clips x 64 frames per rank (= 100k clips over 8 GPUs), all-reduce INSIDE the timed region
    n3 = 12500
    v3 = pkg.make_videos(n3, args.frames, seed=1337 + 3 + 1000 * rank, device=dev)
    dv3 = scorer.to_device(v3)
    ar_ev = []

    def build():
        sc = scorer.centroid_sums(dv3, 10)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        pkg.allreduce_centroid_sums(sc)
        b.record()
        ar_ev.append((a, b))
        return pkg.centroid_finalize(sc)

    ms = timed(torch, dist, dev, world, build, steps)
    cen3, cnt3 = build()
    torch.cuda.synchronize(dev)
    wpv = (args.frames - CLIP_LEN) // STRIDE + 1
    assert int(cnt3.sum().item()) == world * n3 * wpv, "centroid counts do not add up across ranks"
    identical = True
    if world > 1:
        allc = [torch.empty_like(cen3) for _ in range(world)]
        dist.all_gather(allc, cen3)
        identical = all(torch.equal(allc[0], c) for c in allc)
        assert identical, "ranks disagree on the all-reduced centroids"
    ar_us = ar_wait_us = None
    if world > 1:
        # events around the collective on the compute stream: a rank that arrives early also waits for the slowest rank in
        # there, so the MIN over ranks is the collective's own latency (seen by the last rank to arrive), the MAX the skew
        t = torch.tensor([1000.0 * statistics.median(a.elapsed_time(b) for a, b in ar_ev[-steps:])], device=dev, dtype=torch.float64)
        tmin, tmax = t.clone(), t.clone()
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ar_us, ar_wait_us = float(tmin.item()), float(tmax.item())
    out["config3"] = entry(ms, n3, "clips", wpv * GFLOP_PER_WINDOW,
                           f"centroid build: {n3} clips x {args.frames} frames per GPU ({n3 * wpv} windows; 8 GPUs = BASELINE's 100k clips), K1 + encoder + K3, "
                           "NCCL all-reduce of the packed [10,257] sums||counts + finalize inside the timed region")
    out["config3"].update({"allreduce_us": ar_us, "allreduce_incl_rank_skew_us_max": ar_wait_us, "allreduce_bytes": 10 * 257 * 4,
                           "centroids_identical_on_all_ranks": identical})
    del v3, dv3

    # ---- config 5: 4096 clips x 32 frames, encoder forward + TCL similarity matrix (per-rank batch, no all-gather: SURVEY.md §8e)
    v5 = pkg.make_videos(4096, 32, seed=1337 + 5 + 1000 * rank, device=dev)
    dv5 = scorer.to_device(v5)
    y5 = torch.tensor(v5.cls_idx, device=dev, dtype=torch.int32)
    tcl = pkg.TCL()

    def step5():
        return tcl(scorer.encode(dv5)["seq"], y5)

    ms = timed(torch, dist, dev, world, step5, steps * 2)
    loss5 = float(step5().item())
    assert loss5 == loss5, "TCL loss is NaN"
    out["config5"] = entry(ms, 4096, "clips", GFLOP_PER_WINDOW + 2 * 4096 * 256 / 1e9,
                           "4096 clips x 32 frames: encoder forward + TCL forward (Z Z^T as a tcgen05 GEMM, masked row sums in its epilogue)")
    out["config5"]["tcl_loss"] = loss5
    del v5, dv5

    # ---- config 4: 512 sequences x 256 frames as one window each (S = 257), M = 7 (vit + clip + dino), D = 5156
    r7, d7 = pkg.dims_maps(True)
    m7 = pkg.HumanActionScorer(r7, d7, precision=args.precision, max_windows=512)
    m7.load_state_dict(pkg.make_state_dict(r7, d7, seed=1))
    m7.to(dev).eval()
    real7 = pkg.make_videos(20, 256, seed=1337 + 4, appearance=True, device=dev)
    st7 = pkg.compute_stats_from_videos(real7, r7, d7, dev)
    sc7 = pkg.TagScorer(m7, st7, 256, 8, dev)
    v4 = pkg.make_videos(512, 256, seed=1337 + 40 + 1000 * rank, appearance=True, device=dev)
    dv4 = sc7.to_device(v4)
    ms = timed(torch, dist, dev, world, lambda: sc7.score(dv4, centroids), steps)
    out["config4"] = entry(ms, 512, "sequences", 22.362, "512 sequences x 256 frames, clip_len 256 (S = 257), M = 7, D = 5156: K1 + encoder + AC + TC")
    del v4, dv4, sc7, m7, real7

    # ---- K3 / K4 alone on inputs larger than L2 (500,000 windows = 512 MB of embeddings; 100,000 videos)
    hb = {}
    N, V = 500000, 100000
    g = torch.Generator(device=dev).manual_seed(1)
    z = torch.nn.functional.normalize(torch.randn(N, 256, device=dev, generator=g), dim=-1)
    vid_lab = (torch.arange(V, device=dev) % 10).to(torch.int32)
    win_lab = vid_lab.repeat_interleave(5).contiguous()
    tcw = torch.rand(N, device=dev, generator=g)
    seg = (torch.arange(V + 1, device=dev, dtype=torch.int64) * 5).contiguous()
    sc = torch.zeros(10, 257, device=dev)
    ms = timed(torch, None, dev, 1, lambda: pkg.centroid_accumulate(z, win_lab, sc), 10, 3)
    hb["k_centroid_partial+combine (K3)"] = {"ms": ms, "bytes": N * 1028 + 10 * 1028, "rows": N}
    lib = pkg.load_library()
    uh = pkg.scoring.util_handle(dev)
    ac = torch.empty(V, device=dev)
    tc = torch.empty(V, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def k4():
        pkg._lib.check(uh, lib.tag_score(uh, z.data_ptr(), tcw.data_ptr(), seg.data_ptr(), vid_lab.data_ptr(), centroids.data_ptr(), 10, V,
                                         ac.data_ptr(), tc.data_ptr(), stream), "tag_score")

    ms = timed(torch, None, dev, 1, k4, 10, 3)
    hb["k_score (K4)"] = {"ms": ms, "bytes": N * 1028 + V * 24, "rows": N}
    for v in hb.values():
        v["achieved"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9
        v["frac"] = v["achieved"] / pk["hbm_gbs"]
    return out, hb


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
    kinds = (__import__("ctypes").c_double * 24)()
    pkg._lib.check(h, lib.tag_get_profile_kinds(h, kinds, 8), "tag_get_profile_kinds")
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
    del gen_host

    pk = peaks()
    configs, hbm_extra = ({}, {})
    if not args.no_configs and args.precision == "fp16_tc":
        del dv, gen
        torch.cuda.empty_cache()
        configs, hbm_extra = other_configs(args, pkg, torch, dist, dev, rank, world, model, scorer, stats, centroids, pk)

    if rank == 0:
        other_ms, _, other_n, conv_ms, conv_flops, conv_n, og_ms, og_flops, og_n, k1_ms, k1_bytes, k1_n = [float(x) for x in prof]
        achieved = (conv_flops / conv_n) / (conv_ms / conv_n * 1e-3) / 1e12 if conv_n > 0 and conv_ms > 0 else None
        tc_mode = args.precision == "fp16_tc"
        roof_peak = pk["tflops"] if tc_mode else 75.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "conv_gemm_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(args.precision)
        kd = [float(x) for x in kinds]
        hbm = {}
        names = {3: "k_feature_fuse_staged (K1)", 4: "k_merge_fusion_h", 5: "k_finalize (+ per-window TC)", 6: "k_attention_mma", 7: "k_build_tokens"}
        for k, nm in names.items():
            ms_k, by_k, n_k = kd[3 * k], kd[3 * k + 1], kd[3 * k + 2]
            if n_k > 0 and ms_k > 0:
                ach = by_k / (ms_k * 1e-3) / 1e9
                hbm[nm] = {"ms": ms_k / n_k, "bytes": by_k / n_k, "launches": int(n_k), "achieved": ach, "frac": ach / pk["hbm_gbs"]}
        k1 = hbm.get(names[3])
        if k1:   # K1 writes the padded fp16 operand (5,480 B read + 5,888 B written per frame row); SURVEY.md §8d counts unpadded 5,192 B
            k1["bytes_per_frame_row"] = 11368
            k1["achieved_survey_bytes"] = k1["achieved"] * 10672.0 / 11368.0
            k1["frac_survey_bytes"] = k1["frac"] * 10672.0 / 11368.0
        hbm.update(hbm_extra)
        cfg = workload_config(args, world)
        assert cfg["workload"].split(" -> ")[1].startswith(str(n_windows))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if tc_mode else "f32", "data": "synthetic",
            "config": cfg,
            "precision": "fp16 operands, fp32 accumulate/norms (tcgen05)" if tc_mode else "fp32 CUDA cores",
            "windows_per_pass": max_windows,
            "roofline": {"bound": "tensor", "kernel": "k_tcn_block / k_gemm_tc (dilated-conv blocks, 5 taps: fused block kernel at dilation 1/2/4, two GEMM launches at dilation 8)" if tc_mode else "k_gemm_f32 (dilated conv, 5 taps)",
                         "achieved": achieved, "peak": roof_peak, "unit": "TFLOP/s",
                         "frac": (achieved / roof_peak) if achieved else None, "traffic": traffic,
                         "peak_source": pk["source"] if tc_mode else "nominal fp32 CUDA-core peak (~75 TFLOP/s), fp32 mode only",
                         "launches_sampled": int(conv_n), "mean_launch_ms": conv_ms / conv_n if conv_n else None,
                         "share_of_step": {"conv_gemm_ms": conv_ms, "other_gemm_ms": og_ms, "feature_fuse_ms": k1_ms,
                                           "other_kernels_ms": other_ms},
                         "other_gemm_tflops": (og_flops / (og_ms * 1e-3) / 1e12) if og_ms > 0 else None,
                         "feature_fuse_hbm": {"bound": "hbm", "achieved": (k1_bytes / (k1_ms * 1e-3) / 1e9) if k1_ms > 0 else None,
                                              "peak": pk["hbm_gbs"], "unit": "GB/s",
                                              "frac": (k1_bytes / (k1_ms * 1e-3) / 1e9 / pk["hbm_gbs"]) if k1_ms > 0 else None},
                         "whole_encoder_tflops": value / world * (n_windows / args.videos) * GFLOP_PER_WINDOW / 1e3,
                         "whole_encoder_frac": value / world * (n_windows / args.videos) * GFLOP_PER_WINDOW / 1e3 / roof_peak},
            "hbm_kernels": {"peak": pk["hbm_gbs"], "unit": "GB/s", "how": "algorithmic bytes / CUDA-event time; in-step kernels averaged over the "
                            "timed steps (tag_get_profile_kinds), K3 / K4 alone on 500,000 windows (512 MB, larger than L2)", "kernels": hbm},
            "configs": configs,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": int(2 * args.videos * 4),
                    "steps": e2e_steps, "h2d_gbs_measured": h2d_gbs, "host_affinity": numa, "call": "TagScorer.score_stream: every step's 5 input arrays copied from pinned host memory "
                    "(one block per encoder pass, short ramp-up blocks at the start of the stream, prefetched on a copy stream), per-video AC/TC read back to the host every step"},
            "gpu_launches": l1 - l0,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_subprocess(args)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
