"""GPU parity at the sizes BASELINE.json names (pytest -m gpu), CUDA path vs the CPU oracle on the same seeded inputs:

  config 2   2,000 generated videos x 64 frames (10,000 windows, clip 32 / stride 8): per-video AC and TC     eval.py:229-257, :209-226
  config 3   the centroid build over the same 2,000 clips as "real" clips: centroids and counts               utils.py:1018-1045
  config 4   64 sequences of 256 frames, M = 7 (vit + clip + dino), one 256-frame window each (S = 257)       model.py:162-193
  config 5   TCL forward at B = 4096                                                                          losses.py:14-34

Every test prints max / p99 / median relative error and asserts the north-star bar: max < 1e-3 relative for the tensor-core
mode (fp16 operands, fp32 accumulate), ~1e-5 for the fp32 mode where it is run. The oracle is the fp32 restatement
pinned to the unmodified reference (tests/test_oracle_golden.py; it agrees with the reference itself to 2e-7 on these
scores, tests/test_ref_dropin.py). One oracle pass over the 2,000 videos (~40 s on the GPU box's host cores) is shared
by the config-2 and config-3 tests.
"""
import numpy as np
import pytest
import torch

import tag_b200 as tb
from helpers import oracle, max_abs

pytestmark = pytest.mark.gpu
O = oracle()
DEV = "cuda:0"
TOL = 1e-3            # north_star: per-video scores and per-action centroids within 1e-3 relative


def _summary(name, rel):
    rel = np.sort(np.asarray(rel, dtype=np.float64))
    p99 = rel[min(len(rel) - 1, int(0.99 * len(rel)))]
    print(f"{name}: n={len(rel)} max {rel[-1]:.3e}  p99 {p99:.3e}  median {rel[len(rel) // 2]:.3e}")
    return float(rel[-1])


@pytest.fixture(scope="module")
def big():
    """2,000 videos x 64 frames through the oracle once: window embeddings, AC/TC against fixed centroids."""
    torch.set_num_threads(max(1, torch.get_num_threads()))
    dims_raw, dims_diff = tb.dims_maps(False)
    sd = tb.make_state_dict(dims_raw, dims_diff, seed=0)
    real = tb.make_videos(60, 64, seed=1340)
    ostats = O.compute_stats([real.video(v) for v in range(real.n_videos)])
    V = 2000
    gen = tb.make_videos(V, 64, seed=4242)
    g = torch.Generator().manual_seed(5)
    cen0 = torch.nn.functional.normalize(torch.randn(10, 256, generator=g), dim=-1)
    # centroids near the data make AC small (the hard case: a difference of nearly equal unit vectors):
    # take them from the oracle embeddings of the first 200 videos instead of random directions
    label_dict = {c: i for i, c in enumerate(tb.ACTION_CLASSES)}
    with torch.no_grad():
        _, otc, feats = O.score_videos([gen.video(v) for v in range(V)], gen.names, [gen.cls_name(v) for v in range(V)], sd,
                                       dims_raw, dims_diff, ostats, cen0, label_dict, clip_len=32, stride=8, batch=64)
    y = torch.tensor([label_dict[c] for c in feats["cls_names"]])
    cen_near, _ = O.build_centroids(feats["seq_embeds"][:1000], y[:1000], 10)
    oac = O.action_consistency_scores(feats, cen_near, label_dict)
    return dict(dims_raw=dims_raw, dims_diff=dims_diff, sd=sd, ostats=ostats, gen=gen, feats=feats, y=y, cen=cen_near,
                oac=oac, otc=otc, label_dict=label_dict, V=V)


def _scorer(b, precision, max_windows):
    model = tb.HumanActionScorer(b["dims_raw"], b["dims_diff"], precision=precision, max_windows=max_windows)
    model.load_state_dict(b["sd"])
    model.to(DEV).eval()
    return tb.TagScorer(model, b["ostats"], 32, 8, DEV), model


def test_config2_scores_2000_videos_tc(big):
    b = big
    scorer, model = _scorer(b, "fp16_tc", 5920)
    dv = scorer.to_device(b["gen"].to(DEV))
    ac, tc = scorer.score(dv, b["cen"].to(DEV))
    d = scorer.scores_dict(b["gen"], ac, tc)
    assert len(d) == b["V"] and set(d) == set(b["oac"]) == set(b["otc"])
    e_ac = _summary("config 2 (2,000 videos, 10,000 windows) tensor-core AC rel", [abs(d[k]["ac"] - v) / v for k, v in b["oac"].items()])
    e_tc = _summary("config 2 (2,000 videos, 10,000 windows) tensor-core TC rel", [abs(d[k]["tc"] - v) / v for k, v in b["otc"].items()])
    print(f"   AC range {min(b['oac'].values()):.4f} .. {max(b['oac'].values()):.4f}; TC range {min(b['otc'].values()):.4f} .. {max(b['otc'].values()):.4f}")
    assert e_ac < TOL and e_tc < TOL
    # the ragged-window path (explicit window table, K1 per window) sees the same inputs: same bar
    enc = scorer.encode(dv, use_clips=False)
    z = enc["seq"].cpu()
    e_z = float((z - b["feats"]["seq_embeds"]).abs().max())
    print(f"   window-table path: seq-embed max abs {e_z:.3e}")
    assert e_z < 2e-3


def test_config2_scores_400_videos_fp32(big):
    """the reference-precision mode on a 400-video slice of the same set (fp32 CUDA-core GEMMs are ~15x slower)"""
    b = big
    scorer, model = _scorer(b, "fp32", 1000)
    sub = b["gen"].slice(0, 400)
    ac, tc = scorer.score(scorer.to_device(sub.to(DEV)), b["cen"].to(DEV))
    d = scorer.scores_dict(sub, ac, tc)
    e_ac = _summary("config 2 slice (400 videos) fp32 AC rel", [abs(d[k]["ac"] - b["oac"][k]) / b["oac"][k] for k in d])
    e_tc = _summary("config 2 slice (400 videos) fp32 TC rel", [abs(d[k]["tc"] - b["otc"][k]) / b["otc"][k] for k in d])
    assert e_ac < 1e-4 and e_tc < 1e-4


def test_config3_centroid_build_2000_clips(big):
    """build_real_centroids (eval.py:260-286 -> utils.py:1018-1045) over 2,000 clips x 64 frames = 10,000 windows: counts exact,
    centroids within the bar; the two-pass K3 is bit-reproducible; shard sums add up to the full sums (what the all-reduce does)."""
    b = big
    scorer, model = _scorer(b, "fp16_tc", 5920)
    dv = scorer.to_device(b["gen"].to(DEV))
    cen, cnt = scorer.build_centroids(dv, 10)
    ocen, ocnt = O.build_centroids(b["feats"]["seq_embeds"], b["y"], 10)
    assert max_abs(cnt.cpu(), ocnt) == 0 and float(cnt.sum()) == 10000
    rel = ((cen.cpu() - ocen).abs() / ocen.abs().clamp_min(1e-3)).flatten().numpy()     # elements are ~0.06: floor the denominator
    e_abs = max_abs(cen.cpu(), ocen)
    e_dir = float((cen.cpu() - ocen).norm(dim=1).max())                                  # centroids are unit vectors
    _summary("config 3 (2,000 clips) centroid element rel", rel)
    print(f"   centroid max abs {e_abs:.3e}, max |c - c_ref|_2 {e_dir:.3e}")
    assert e_dir < TOL and e_abs < 2e-4
    # determinism: the same build again is bit-identical (no floating-point atomics in K3)
    cen2, cnt2 = scorer.build_centroids(dv, 10)
    assert torch.equal(cen, cen2) and torch.equal(cnt, cnt2)
    # sharding: sums of 4 contiguous shards == sums of the whole set (up to fp32 reassociation), counts exactly
    full = scorer.centroid_sums(dv, 10)
    parts = sum(scorer.centroid_sums(scorer.to_device(b["gen"].slice(*tb.shard_range(b["V"], r, 4)).to(DEV)), 10) for r in range(4))
    assert torch.equal(parts[:, 256], full[:, 256])
    assert float((parts[:, :256] - full[:, :256]).abs().max()) < 1e-3 * float(full[:, :256].abs().max())
    # fp32 K3 alone against the oracle on the oracle's own embeddings: round-off only
    sc = torch.zeros(10, 257, device=DEV)
    tb.centroid_accumulate(b["feats"]["seq_embeds"].to(DEV), b["y"].to(DEV, torch.int32), sc)
    c3, n3 = tb.centroid_finalize(sc)
    assert max_abs(n3.cpu(), ocnt) == 0 and max_abs(c3.cpu(), ocen) < 1e-6


def test_config4_long_clips_64_windows_tc():
    """BASELINE config 4's shape: 256-frame sequences as ONE window each (clip_len 256, S = 257), M = 7 with appearance
    features, D = 5156 — 64 sequences (the bench runs 512) against the oracle."""
    dims_raw, dims_diff = tb.dims_maps(True)
    sd = tb.make_state_dict(dims_raw, dims_diff, seed=1)
    real = tb.make_videos(12, 256, seed=1341, appearance=True)
    ostats = O.compute_stats([real.video(v) for v in range(real.n_videos)])
    V = 64
    gen = tb.make_videos(V, 256, seed=4343, appearance=True)
    label_dict = {c: i for i, c in enumerate(tb.ACTION_CLASSES)}
    g = torch.Generator().manual_seed(6)
    cen0 = torch.nn.functional.normalize(torch.randn(10, 256, generator=g), dim=-1)
    with torch.no_grad():
        _, otc, feats = O.score_videos([gen.video(v) for v in range(V)], gen.names, [gen.cls_name(v) for v in range(V)], sd,
                                       dims_raw, dims_diff, ostats, cen0, label_dict, clip_len=256, stride=8, batch=8)
    y = torch.tensor([label_dict[c] for c in feats["cls_names"]])
    cen, _ = O.build_centroids(feats["seq_embeds"], y, 10)
    oac = O.action_consistency_scores(feats, cen, label_dict)
    model = tb.HumanActionScorer(dims_raw, dims_diff, precision="fp16_tc", max_windows=64)
    model.load_state_dict(sd)
    model.to(DEV).eval()
    scorer = tb.TagScorer(model, ostats, 256, 8, DEV)
    dv = scorer.to_device(gen.to(DEV))
    ac, tc = scorer.score(dv, cen.to(DEV))
    d = scorer.scores_dict(gen, ac, tc)
    e_ac = _summary("config 4 (64 x 256 frames, M=7) tensor-core AC rel", [abs(d[k]["ac"] - v) / v for k, v in oac.items()])
    e_tc = _summary("config 4 (64 x 256 frames, M=7) tensor-core TC rel", [abs(d[k]["tc"] - v) / v for k, v in otc.items()])
    assert e_ac < TOL and e_tc < TOL
    gcen, gcnt = scorer.build_centroids(dv, 10)
    assert max_abs(gcnt.cpu(), torch.bincount(y, minlength=10).float()) == 0
    assert float((gcen.cpu() - cen).norm(dim=1).max()) < TOL


def test_config5_tcl_b4096():
    """TCL forward (losses.py:14-34) at the training batch of BASELINE config 5: 4096 unit-norm embeddings, P x K labels
    (10 classes), per-row loss terms and the scalar against the oracle (fp64 and fp32)."""
    g = torch.Generator().manual_seed(9)
    B = 4096
    y = torch.arange(B) % 10
    proto = torch.randn(10, 256, generator=g)
    z = torch.nn.functional.normalize(proto[y] + 1.5 * torch.randn(B, 256, generator=g), dim=-1)     # clustered like trained embeddings
    ref64 = O.tcl_loss_rows(z.double(), y)
    ref32 = float(O.tcl_loss(z, y))
    tcl = tb.TCL()
    rows = tcl.loss_rows(z.to(DEV), y.to(DEV)).cpu().double()
    rel = ((rows - ref64).abs() / ref64.abs()).numpy()
    e = _summary("config 5 TCL rows (B=4096) rel", rel)
    got = float(tcl(z.to(DEV), y.to(DEV)))
    print(f"   TCL scalar: cuda {got:.7f}  oracle fp64 {float(ref64.mean()):.7f}  oracle fp32 {ref32:.7f}")
    assert e < TOL
    assert abs(got - float(ref64.mean())) < 1e-4 * abs(float(ref64.mean()))
