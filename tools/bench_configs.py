"""Throughput of the OTHER BASELINE.json configurations (1, 3-per-GPU, 4, 5) on one B200 — parity cases in the test suite,
timed here only so DESIGN.md can state how the path behaves outside the config-2 bench line. Device-resident inputs,
CUDA events, 3 warm-ups + 5 timed repetitions each."""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("video-gen-evals_b200")
DEV = torch.device("cuda", 0)
GF = {(32, 5): 2.0248, (256, 7): 22.362}       # GFLOP per window, SURVEY.md §8(d)


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def setup(appearance, frames, clip, stride, n_real, max_windows):
    dr, dd = pkg.dims_maps(appearance)
    model = pkg.HumanActionScorer(dr, dd, precision="fp16_tc", max_windows=max_windows)
    model.load_state_dict(pkg.make_state_dict(dr, dd, seed=0))
    model.to(DEV).eval()
    real = pkg.make_videos(n_real, frames, seed=1340, appearance=appearance, device=DEV)
    stats = pkg.compute_stats_from_videos(real, dr, dd, DEV)
    scorer = pkg.TagScorer(model, stats, clip, stride, DEV)
    return scorer, real


out = []
# config 1: 64 videos x 32 frames, one window each
scorer, real = setup(False, 32, 32, 8, 100, 64)
cen, _ = scorer.build_centroids(scorer.to_device(real), 10)
dv = scorer.to_device(pkg.make_videos(64, 32, seed=1338, device=DEV))
ms = timed(lambda: scorer.score(dv, cen))
out.append({"config": 1, "what": "64 videos x 32 frames, score", "ms": ms, "videos_per_s": 64 / ms * 1e3})
# config 3 (per-GPU share at 8 GPUs): 12,500 clips x 64 frames -> 62,500 windows, centroid sums
scorer, real = setup(False, 64, 32, 8, 100, 12800)
dv = scorer.to_device(pkg.make_videos(12500, 64, seed=1339, device=DEV))
ms = timed(lambda: scorer.centroid_sums(dv, 10), reps=3, warm=2)
out.append({"config": 3, "what": "12,500 clips x 64 frames (1/8 of 100k), centroid sums", "ms": ms, "windows_per_s": 62500 / ms * 1e3,
            "tflops": 62500 * GF[(32, 5)] / ms})
del dv
# config 4: 512 sequences x 256 frames, one 256-frame window each, M = 7
scorer, real = setup(True, 256, 256, 256, 20, 512)
cen, _ = scorer.build_centroids(scorer.to_device(real), 10)
dv = scorer.to_device(pkg.make_videos(512, 256, seed=1341, appearance=True, device=DEV))
ms = timed(lambda: scorer.score(dv, cen), reps=3, warm=2)
out.append({"config": 4, "what": "512 x 256-frame windows, M=7, score", "ms": ms, "windows_per_s": 512 / ms * 1e3,
            "tflops": 512 * GF[(256, 7)] / ms})
del dv
# config 5: 4096 clips x 32 frames -> embeddings + TCL row losses
scorer, real = setup(False, 32, 32, 8, 100, 4096)
vb = pkg.make_videos(4096, 32, seed=1342, device=DEV)
dv = scorer.to_device(vb)
labels = torch.tensor(vb.cls_idx, device=DEV, dtype=torch.int32)
tcl = pkg.TCL()
def step5():
    enc = scorer.encode(dv)
    return tcl.loss_rows(enc["seq"], labels)
ms = timed(step5)
out.append({"config": 5, "what": "4096 clips x 32 frames, encoder + TCL forward", "ms": ms, "clips_per_s": 4096 / ms * 1e3,
            "tflops": 4096 * GF[(32, 5)] / ms})
for o in out:
    print(json.dumps(o), flush=True)
