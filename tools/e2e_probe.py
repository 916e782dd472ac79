"""End-to-end streaming probe: score_stream over K host batches with different block plans."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("video-gen-evals_b200")
dev = torch.device("cuda", 0)
dr, dd = pkg.dims_maps(False)
model = pkg.HumanActionScorer(dr, dd, precision="fp16_tc", max_windows=13024)
model.load_state_dict(pkg.make_state_dict(dr, dd, seed=0)); model.to(dev).eval()
real = pkg.make_videos(200, 64, seed=1340, device=dev)
stats = pkg.compute_stats_from_videos(real, dr, dd, dev)
scorer = pkg.TagScorer(model, stats, 32, 8, dev)
cen, _ = scorer.build_centroids(scorer.to_device(real), 10)
gen = pkg.make_videos(5000, 64, seed=1339, device=dev)
dv = scorer.to_device(gen)
for _ in range(3):
    scorer.score(dv, cen)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    scorer.score(dv, cen)
torch.cuda.synchronize()
print("resident: %.2f ms/step" % ((time.perf_counter() - t0) / 5 * 1e3))
host = gen.to("cpu").pin()
for pieces, prefetch in ((2, 2), (2, 3), (None, 3), (None, 2), (4, 3)):
    for _ in scorer.score_stream((host for _ in range(2)), cen, pieces=pieces, prefetch=prefetch):
        pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in scorer.score_stream((host for _ in range(5)), cen, pieces=pieces, prefetch=prefetch):
        pass
    torch.cuda.synchronize()
    print("pieces=%s prefetch=%d: %.2f ms/step" % (pieces, prefetch, (time.perf_counter() - t0) / 5 * 1e3), flush=True)

# how much does a concurrent (independent) pinned H2D stream slow the HBM-resident step? (DMA + HBM-write contention / power)
src = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
dst = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
cs = torch.cuda.Stream(device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(cs):
        for _ in range(2):                      # ~2 GB per step, like the end-to-end leg
            dst.copy_(src, non_blocking=True)
    scorer.score(dv, cen)
torch.cuda.synchronize()
print("resident + independent 2 GB/step H2D on a side stream: %.2f ms/step" % ((time.perf_counter() - t0) / 5 * 1e3))
