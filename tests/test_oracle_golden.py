"""CPU suite: pin the oracle (oracle/tag_oracle.py) against vectors produced by the unmodified
reference (tests/golden/make_golden.py). The reference has no tests of its own."""
import numpy as np
import pytest
import torch

from helpers import golden_case, oracle, rel_err, max_abs, GOLDEN
import os

O = oracle()


@pytest.fixture(scope="module")
def deltas():
    return np.load(os.path.join(GOLDEN, "deltas.npz"))


def test_delta_functions_match_reference(deltas):
    from helpers import synth
    vb = synth.make_videos(3, 24, seed=77)
    for v in range(3):
        d = vb.video(v)
        assert max_abs(O.vit_delta(d["vit"]), deltas[f"v{v}.vit_delta"]) == 0.0
        assert max_abs(O.rotmat_delta(d["pose"]), deltas[f"v{v}.pose_delta"]) < 1e-6
        assert max_abs(O.rotmat_delta(d["global_orient"]), deltas[f"v{v}.gori_delta"]) < 1e-6
        assert max_abs(O.betas_delta(d["betas"]), deltas[f"v{v}.betas_delta"]) == 0.0
        kd, nref = O.procrustes_kp_delta(d["keypoints"])
        assert max_abs(kd, deltas[f"v{v}.kp_delta"]) < 1e-6
        assert nref == 0          # the synthetic generator stays in the closed-form regime
        # closed form (what the CUDA kernel evaluates) agrees with the SVD form when det(H) > 0
        kc, det = O.procrustes_kp_delta_closed_form(d["keypoints"])
        assert bool((det > 0).all())
        assert max_abs(kc, deltas[f"v{v}.kp_delta"]) < 5e-7


def test_delta_edge_cases(deltas):
    R = torch.from_numpy(deltas["edge.R"])
    got = O.rotmat_delta(R)
    assert max_abs(got, deltas["edge.R_delta"]) < 2e-5      # theta near pi is ill-conditioned in fp32
    assert float(got[5].abs().max()) == 0.0                  # identical consecutive frames -> exact zero
    assert float(got[0].abs().max()) == 0.0                  # first frame pairs with itself
    x = torch.from_numpy(deltas["edge.x"])
    assert max_abs(O.vit_delta(x), deltas["edge.x_delta"]) == 0.0   # all-zero row: eps 1e-12 clamp
    kp = torch.from_numpy(deltas["edge.kp"])
    kd, _ = O.procrustes_kp_delta(kp)
    assert max_abs(kd, deltas["edge.kp_delta"]) < 1e-6
    kc, det = O.procrustes_kp_delta_closed_form(kp)
    assert max_abs(kc, deltas["edge.kp_delta"]) < 5e-7


def test_procrustes_mirror_regime_closed_form():
    """det(H) < 0 pairs: the closed form (polar-reflection angle) reproduces the reference's SVD path
    (tests/golden/deltas_mirror.npz, made by the unmodified reference), as does the oracle's own SVD restatement."""
    gold = np.load(os.path.join(GOLDEN, "deltas_mirror.npz"))
    for tag in ("iid", "flip"):
        kp = torch.from_numpy(gold[f"{tag}.kp"])
        kd, nref = O.procrustes_kp_delta(kp)
        assert max_abs(kd, gold[f"{tag}.kp_delta"]) < 1e-6
        kc, det = O.procrustes_kp_delta_closed_form(kp)
        assert int((det < 0).sum()) == nref and nref >= 10
        assert max_abs(kc, gold[f"{tag}.kp_delta"]) < 2e-6, tag
    # 20 k random 2x2 cross-covariances, both regimes, against torch.linalg.svd directly
    g = torch.Generator().manual_seed(3)
    worst = 0.0
    for _ in range(200):
        kp = torch.rand(101, 120, generator=g)
        kd, _ = O.procrustes_kp_delta(kp)
        kc, _ = O.procrustes_kp_delta_closed_form(kp)
        worst = max(worst, max_abs(kc, kd))
    assert worst < 5e-6, worst


def test_slice_or_pad_index():
    assert O.slice_or_pad_index(10, 2, 4).tolist() == [2, 3, 4, 5]
    assert O.slice_or_pad_index(10, 8, 4).tolist() == [8, 9, 9, 9]          # tail repeats last frame
    assert O.slice_or_pad_index(3, 0, 5).tolist() == [0, 1, 2, 2, 2]
    assert O.slice_or_pad_index(10, -1, 3).tolist() == [0, 0, 0]             # utils.py:371-374
    assert O.slice_or_pad_index(10, 10, 3).tolist() == [9, 9, 9]


@pytest.mark.parametrize("tag", ["m5_t32", "m7_t256"])
def test_stats_match_reference(tag):
    g = golden_case(tag)
    st = O.compute_stats(g.train_videos())
    gold = g.stats()
    assert set(gold.keys()) <= set(st.keys())
    for k, v in gold.items():
        assert max_abs(st[k], v) <= 1e-6 * max(1.0, float(v.abs().max())), k


@pytest.mark.parametrize("tag", ["m5_t32", "m7_t256"])
def test_window_features_match_reference(tag):
    g = golden_case(tag)
    stats = g.stats()
    wins = g.gen_windows()
    rowsum = g.npz["feats_rowsum"]
    abssum = g.npz["feats_abssum"]
    sel = {w: i for i, w in enumerate(g.meta["full_feat_windows"])}
    for i, (v, s) in enumerate(wins):
        f, nref = O.window_features(g.gen.video(v), s, g.clip_len, stats, g.mods)
        assert nref == 0
        assert f.shape == (g.clip_len, sum(g.dims_raw.values()) + sum(g.dims_diff.values()))
        assert abs(float(f.double().abs().sum()) - abssum[i]) <= 1e-5 * abssum[i]
        assert max_abs(f.double().sum(0).float(), rowsum[i]) <= 1e-3
        if i in sel:
            assert max_abs(f, g.npz["feats_sel"][sel[i]]) < 2e-5


def _oracle_features(g, windows, videos, dtype=torch.float32):
    stats = g.stats()
    feats = torch.stack([O.window_features(videos.video(v), s, g.clip_len, stats, g.mods)[0] for v, s in windows], 0)
    sd = {k: t.to(dtype) for k, t in g.sd.items()}
    with torch.no_grad():
        outs = [O.encoder_forward(sd, feats[i:i + 16].to(dtype), g.dims_raw, g.dims_diff) for i in range(0, len(feats), 16)]
    return feats, torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs]), torch.cat([o[2] for o in outs])


@pytest.mark.parametrize("tag", ["m5_t32", "m7_t256"])
def test_encoder_and_scores_match_reference(tag):
    g = golden_case(tag)
    wins = g.gen_windows()
    feats, seq, frm, tok = _oracle_features(g, wins, g.gen)
    assert max_abs(seq, g.npz["seq_embeds"]) < 5e-6
    assert max_abs(frm[g.meta["frame_windows"]], g.npz["frame_embeds_sel"]) < 5e-6
    # centroids from the real train windows (utils.py:1018-1045)
    rw = g.real_windows()
    _, rseq, _, _ = _oracle_features(g, rw, g.real)
    y = torch.tensor([g.label_dict[g.real.cls_name(v)] for v, _ in rw])
    cen, counts = O.build_centroids(rseq, y, len(g.label_dict))
    assert max_abs(counts, g.npz["counts"]) == 0
    assert max_abs(cen, g.npz["centroids"]) < 5e-6
    features = {"seq_embeds": seq, "frame_embeds": frm, "vid_names": g.meta["vid_names"],
                "cls_names": g.meta["cls_names"]}
    ac = O.action_consistency_scores(features, cen, g.label_dict)
    tc = O.temporal_coherence_scores(features)
    assert set(ac) == set(g.meta["ac"]) and set(tc) == set(g.meta["tc"])
    for k in ac:
        assert abs(ac[k] - g.meta["ac"][k]) <= 1e-4 * g.meta["ac"][k], (k, ac[k], g.meta["ac"][k])
        assert abs(tc[k] - g.meta["tc"][k]) <= 1e-5 * g.meta["tc"][k]
    if g.meta.get("tcl") is not None:
        yy = torch.arange(seq.shape[0]) % 4
        assert abs(float(O.tcl_loss(seq, yy)) - g.meta["tcl"]) < 1e-4


def test_encoder_intermediate_taps_match_reference():
    g = golden_case("m5_t32")
    wins = [g.gen_windows()[i] for i in g.meta["tap_windows"]]
    stats = g.stats()
    feats = torch.stack([O.window_features(g.gen.video(v), s, g.clip_len, stats, g.mods)[0] for v, s in wins], 0)
    taps = {}
    with torch.no_grad():
        O.encoder_forward(g.sd, feats, g.dims_raw, g.dims_diff, taps=taps)
    checked = 0
    for k in g.npz.files:
        if not k.startswith("tap."):
            continue
        name = k[4:]
        if name.endswith(".out"):
            continue
        key = name
        if key in taps:
            assert max_abs(taps[key], g.npz[k]) < 2e-5, key
            checked += 1
    assert checked >= 30


def test_oracle_fp64_noise_floor():
    """fp32 oracle vs the reference's own fp64 copy: the reference's fp32 noise."""
    g = golden_case("m5_t32")
    wins = g.gen_windows()[:8]
    _, seq32, _, _ = _oracle_features(g, wins, g.gen)
    assert max_abs(seq32, g.npz["seq_embeds_fp64"][:8]) < 5e-6


def test_window_rows_are_whole_clip_rows_except_the_first_frame():
    """The invariant the frame-table mode of the CUDA path (tag_encode_clips) rests on, checked on the pinned oracle: the
    z-scored feature row of frame t >= 1 of a window starting at s equals row s + t of the features computed over the
    whole clip (utils.py:142-217 deltas depend only on the previous frame), while the window's first row has zero
    differences whatever precedes it."""
    g = golden_case("m5_t32")
    stats = g.stats()
    v = next(i for i in range(g.gen.n_videos) if g.gen.length(i) >= 64)
    vid = g.gen.video(v)
    L = g.gen.length(v)
    whole, _ = O.window_features(vid, 0, L, stats, g.mods)
    D_raw = sum(int(g.dims_raw[m]) for m in g.mods)
    for s in (0, 8, 24, L - 32):
        win, _ = O.window_features(vid, s, 32, stats, g.mods)
        assert torch.equal(win[1:], whole[s + 1:s + 32])                       # bit-identical rows
        assert torch.equal(win[0, :D_raw], whole[s, :D_raw])                   # raw part of the first row too
        assert torch.equal(win[0, D_raw:], whole[0, D_raw:])                   # differences: the same z-scored zero everywhere
        if s > 0:
            assert not torch.equal(win[0, D_raw:], whole[s, D_raw:])           # ... not the clip's own difference at that frame
