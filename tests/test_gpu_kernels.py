"""GPU parity tests (run on the B200 box: pytest -m gpu). Everything goes through the C ABI of
libtag_b200.so via the Python shim; the oracle (oracle/tag_oracle.py) and the committed reference
goldens are only the checkers. Tolerances: north-star bar is 1e-3 relative on per-video scores and
centroids; the fp32 mode is held to ~1e-5, K1/K3/K4 to fp32 round-off."""
import ctypes as C
import math
import os

import numpy as np
import pytest
import torch

import tag_b200 as tb
from tag_b200 import _lib
from helpers import golden_case, oracle, rel_err, max_abs, synth

pytestmark = pytest.mark.gpu
O = oracle()
DEV = "cuda:0"


def _dv_and_fuser(g, vb):
    fuser = tb.FeatureFuser(g.dims_raw, g.dims_diff, DEV)
    return tb.DeviceVideos(vb, g.mods, DEV), fuser


# ------------------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("tag", ["m5_t32", "m7_t256"])
def test_feature_fuse_matches_reference_golden_and_oracle(tag):
    g = golden_case(tag)
    dv, fuser = _dv_and_fuser(g, g.gen)
    wins = g.gen_windows()
    ds = tb.WindowDataset(wins, g.clip_len, stats=g.stats(), videos=dv, dims_map_raw=g.dims_raw,
                          dims_map_diff=g.dims_diff, fuser=fuser)
    feats = torch.cat([b[0] for b in ds.batches(16)], 0).cpu()
    assert feats.shape == (len(wins), g.clip_len, fuser.D)
    assert int(ds._last_flags.item()) == 0
    # golden (reference WindowDataset): per-window column sums + |.| sums for every window, full tensors for a few
    assert max_abs(feats.double().sum(1).float(), g.npz["feats_rowsum"]) < 2e-3
    ab = feats.double().abs().sum(dim=(1, 2)).numpy()
    assert np.max(np.abs(ab - g.npz["feats_abssum"]) / g.npz["feats_abssum"]) < 1e-5
    for i, w in enumerate(g.meta["full_feat_windows"]):
        # z-scored deltas amplify fp32 round-off (differences of nearly equal numbers divided by a small std)
        assert max_abs(feats[w], g.npz["feats_sel"][i]) < 2e-4
        assert float((feats[w] - torch.from_numpy(g.npz["feats_sel"][i])).abs().median()) < 1e-6
    # oracle on every window (z-scored values are O(1); diffs are divided by small stds)
    stats = g.stats()
    for i in range(0, len(wins), max(1, len(wins) // 8)):
        v, s = wins[i]
        f, _ = O.window_features(g.gen.video(v), s, g.clip_len, stats, g.mods)
        # z-scored deltas are fp32 differences of nearly equal numbers divided by a small std: the reference's own
        # fp32 noise is ~1e-5 here; hold the max to 2e-4 and the typical element to 1e-6
        assert max_abs(feats[i], f) < 2e-4, i
        assert float((feats[i] - f).abs().median()) < 1e-6, i


def test_feature_fuse_no_stats_and_getitem():
    g = golden_case("m5_t32")
    dv, fuser = _dv_and_fuser(g, g.gen)
    ds = tb.WindowDataset([(4, 0), (0, 8), (8, 0)], 32, stats=None, videos=dv, dims_map_raw=g.dims_raw,
                          dims_map_diff=g.dims_diff, fuser=fuser)
    for i, (v, s) in enumerate(ds.samples):
        f, cls, name = ds[i]
        ref, _ = O.window_features(g.gen.video(v), s, 32, None, g.mods)
        assert cls == g.gen.cls_name(v) and name == g.gen.names[v]
        assert max_abs(f.cpu(), ref) < 2e-6


def test_feature_fuse_slice_or_pad_edges():
    """start < 0, start >= L, ragged tail (utils.py:366-381)."""
    g = golden_case("m5_t32")
    dv, fuser = _dv_and_fuser(g, g.gen)
    cases = [(0, -3), (0, 64), (0, 60), (8, 0), (4, 15)]       # video 8 has 7 frames, video 4 has 20
    wv = torch.tensor([c[0] for c in cases], dtype=torch.int32, device=DEV)
    ws = torch.tensor([c[1] for c in cases], dtype=torch.int32, device=DEV)
    feats, flags = fuser.fuse(dv, wv, ws, 32, None, None)
    for i, (v, s) in enumerate(cases):
        ref, _ = O.window_features(g.gen.video(v), s, 32, None, g.mods)
        assert max_abs(feats[i].cpu(), ref) < 2e-6, (v, s)


@pytest.mark.parametrize("tag,with_stats", [("m5_t32", True), ("m5_t32", False), ("m7_t256", True)])
def test_feature_fuse_fp16_operand_matches_fp32_features(tag, with_stats):
    """The tensor-core path's K1 (shared-memory staged kernel, padded fp16 rows) against the fp32 K1 output on the same
    windows — regular windows plus every slice-or-pad edge (negative / out-of-range start, ragged tail, short video):
    same values up to fp16 rounding, pad columns exactly zero."""
    g = golden_case(tag)
    dv, fuser = _dv_and_fuser(g, g.gen)
    T = g.clip_len
    lens = [g.gen.length(v) for v in range(g.gen.n_videos)]
    vs = int(np.argmin(lens))                                   # the shortest video: ragged tails / short-video padding
    wins = list(g.gen_windows()) + [(0, -3), (0, lens[0]), (0, lens[0] - 4), (vs, 0), (vs, 3), (vs, lens[vs] // 2)]
    wv = torch.tensor([c[0] for c in wins], dtype=torch.int32, device=DEV)
    ws = torch.tensor([c[1] for c in wins], dtype=torch.int32, device=DEV)
    mean = std = None
    if with_stats:
        mean, std = tb.features.stats_vectors(g.stats(), g.mods, DEV)
    feats, flags = fuser.fuse(dv, wv, ws, T, mean, std)
    lib = _lib.load()
    d16 = C.c_int32(0)
    _lib.check(fuser.handle, lib.tag_debug_feature_fuse16(fuser.handle, C.byref(dv.c), None, None, None, None, 0, T, None,
                                                          C.byref(d16), None, None), "d16")
    D16 = d16.value
    N = len(wins)
    f16 = torch.full((N * T, D16), float("nan"), device=DEV, dtype=torch.float16)
    flags16 = torch.zeros(1, device=DEV, dtype=torch.int32)
    _lib.check(fuser.handle, lib.tag_debug_feature_fuse16(fuser.handle, C.byref(dv.c), _lib.ptr(mean), _lib.ptr(std), wv.data_ptr(),
                                                          ws.data_ptr(), N, T, f16.data_ptr(), C.byref(d16), flags16.data_ptr(),
                                                          torch.cuda.current_stream().cuda_stream), "tag_debug_feature_fuse16")
    torch.cuda.synchronize()
    assert int(flags16.item()) == int(flags.item())
    # expected operand: every block padded to 64 columns, raw blocks then diff blocks
    exp = torch.zeros(N * T, D16, device=DEV, dtype=torch.float32)
    f2 = feats.reshape(N * T, -1)
    src = dst = 0
    for dims in (g.dims_raw, g.dims_diff):
        for m in g.mods:
            d = int(dims[m])
            exp[:, dst:dst + d] = f2[:, src:src + d]
            src += d
            dst += (d + 63) // 64 * 64
    assert dst == D16 and src == f2.shape[1]
    got = f16.float()
    assert torch.isfinite(got).all()
    pad = torch.ones(D16, dtype=torch.bool, device=DEV)
    dst = 0
    for dims in (g.dims_raw, g.dims_diff):
        for m in g.mods:
            d = int(dims[m])
            pad[dst:dst + d] = False
            dst += (d + 63) // 64 * 64
    assert float(got[:, pad].abs().max()) == 0.0
    # fp16 rounding of the same fp32 value: half an ulp (2^-11 relative) plus the two kernels' own fp32 differences
    err = (got - exp).abs()
    tol = exp.abs() * 2.0 ** -10 + 3e-4
    assert bool((err <= tol).all()), f"max err {float(err.max()):.3e} at {int(err.argmax())}"
    assert float(err.median()) < 2e-4


def _kp_videos(kp: torch.Tensor):
    """a one-video batch whose keypoints are `kp` [L,120] (the other modalities are seeded filler)"""
    vb = synth.make_videos(1, kp.shape[0], seed=5)
    vb.kp = kp.clone()
    return vb


def test_feature_fuse_mirror_regime_matches_reference():
    """det(H) < 0 (mirror-like consecutive keypoint frames, utils.py:202-215): K1 evaluates the polar-reflection closed
    form and must reproduce the REFERENCE's own `_procrustes_kp_delta` output (tests/golden/deltas_mirror.npz: i.i.d.
    random frames, and a smooth sequence whose odd frames are left/right flipped); such frames are still counted."""
    g = golden_case("m5_t32")
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deltas_mirror.npz"))
    off = sum(g.dims_raw.values()) + g.dims_diff["vit"] + g.dims_diff["global"] + g.dims_diff["pose"] + g.dims_diff["beta"]
    for tag in ("iid", "flip"):
        kp = torch.from_numpy(gold[f"{tag}.kp"])
        L = kp.shape[0]
        vb = _kp_videos(kp)
        dv, fuser = _dv_and_fuser(g, vb)
        wv = torch.zeros(1, dtype=torch.int32, device=DEV)
        ws = torch.zeros(1, dtype=torch.int32, device=DEV)
        feats, flags = fuser.fuse(dv, wv, ws, L, None, None)
        _, det = O.procrustes_kp_delta_closed_form(kp)
        n_mirror = int((det < 0).sum())
        assert n_mirror >= (L - 1 if tag == "flip" else 10)
        assert int(flags.item()) == n_mirror
        assert max_abs(feats[0].cpu()[:, off:off + 120], gold[f"{tag}.kp_delta"]) < 2e-6, tag
        # the staged (tensor-core path) K1 takes the same branch
        lib = _lib.load()
        d16 = C.c_int32(0)
        _lib.check(fuser.handle, lib.tag_debug_feature_fuse16(fuser.handle, C.byref(dv.c), None, None, None, None, 0, L, None,
                                                              C.byref(d16), None, None), "d16")
        f16 = torch.full((L, d16.value), float("nan"), device=DEV, dtype=torch.float16)
        fl16 = torch.zeros(1, device=DEV, dtype=torch.int32)
        _lib.check(fuser.handle, lib.tag_debug_feature_fuse16(fuser.handle, C.byref(dv.c), None, None, wv.data_ptr(), ws.data_ptr(),
                                                              1, L, f16.data_ptr(), C.byref(d16), fl16.data_ptr(),
                                                              torch.cuda.current_stream().cuda_stream), "tag_debug_feature_fuse16")
        torch.cuda.synchronize()
        off16 = sum((int(g.dims_raw[m]) + 63) // 64 * 64 for m in g.mods) + \
            sum((int(g.dims_diff[m]) + 63) // 64 * 64 for m in g.mods if m != "kp2d")
        got = f16[:, off16:off16 + 120].float().cpu()
        ref = torch.from_numpy(gold[f"{tag}.kp_delta"])
        assert int(fl16.item()) == n_mirror
        assert bool(((got - ref).abs() <= ref.abs() * 2.0 ** -10 + 1e-5).all()), tag


def test_feature_fuse_delta_edge_vectors():
    """The reference-made edge vectors of tests/golden/deltas.npz THROUGH K1 (not only the CPU oracle): rotations with
    theta near pi and a duplicated frame (utils.py:130-140, :165-174), an all-zero appearance row (1e-12 clamp,
    :142-147), invisible (-1) keypoints with a duplicated frame (:177-217)."""
    g = golden_case("m5_t32")
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deltas.npz"))
    raw_total = sum(g.dims_raw.values())
    doff = {}
    o = raw_total
    for m in g.mods:
        doff[m] = o
        o += g.dims_diff[m]
    # rotations: 16 frames x 4 joints of edge.R fill joints 0..3 of `pose` (identity elsewhere) and joint 0 feeds `global`
    R = torch.from_numpy(gold["edge.R"])                       # [16,4,3,3]
    L = R.shape[0]
    vb = synth.make_videos(1, L, seed=6)
    vb.pose = torch.eye(3).expand(L, 23, 3, 3).clone()
    vb.pose[:, :4] = R
    vb.gori = R[:, :1].clone()
    x = torch.from_numpy(gold["edge.x"])                       # [8,64] -> first 64 vit columns of the first 8 frames
    vb.vit = torch.zeros(L, 1024)
    vb.vit[:8, :64] = x
    vb.vit[8:, 0] = 1.0
    kp = torch.from_numpy(gold["edge.kp"])                     # [8,120]
    vb.kp = torch.cat([kp, kp[-1:].expand(L - 8, 120)], 0).clone()
    dv, fuser = _dv_and_fuser(g, vb)
    wv = torch.zeros(1, dtype=torch.int32, device=DEV)
    ws = torch.zeros(1, dtype=torch.int32, device=DEV)
    feats, flags = fuser.fuse(dv, wv, ws, L, None, None)
    f = feats[0].cpu()
    Rd = torch.from_numpy(gold["edge.R_delta"])                # [16,4,3]
    got_pose = f[:, doff["pose"]:doff["pose"] + 69].reshape(L, 23, 3)
    assert max_abs(got_pose[:, :4], Rd) < 5e-5                 # theta near pi is ill-conditioned in fp32 (values up to pi: 1.6e-5 relative)
    assert float(got_pose[:, 4:].abs().max()) == 0.0           # identity joints -> exactly zero
    assert float(got_pose[5, :4].abs().max()) < 1e-6           # duplicated frame -> zero rotation
    assert max_abs(f[:, doff["global"]:doff["global"] + 3], Rd[:, 0]) < 5e-5
    xd = torch.from_numpy(gold["edge.x_delta"])                # zero row: normalises to zero, no NaN
    assert max_abs(f[:8, doff["vit"]:doff["vit"] + 64], xd) < 1e-6
    assert bool(torch.isfinite(f).all())
    assert max_abs(f[:8, doff["kp2d"]:doff["kp2d"] + 120], gold["edge.kp_delta"]) < 2e-6
    assert float(f[8:, doff["kp2d"]:doff["kp2d"] + 120].abs().max()) < 1e-6    # repeated last frame -> zero motion


# ------------------------------------------------------------------------------------------- N2 stats
@pytest.mark.parametrize("tag", ["m5_t32", "m7_t256"])
def test_stats_match_reference(tag):
    g = golden_case(tag)
    idx = [g.real_index[n] for n in g.meta["train_names"]]
    st = tb.compute_stats_from_videos(g.real.select(idx), g.dims_raw, g.dims_diff, DEV)
    gold = g.stats()
    for k, v in gold.items():
        got = getattr(st, k)
        assert got is not None, k
        assert max_abs(got, v) <= 2e-6 * max(1.0, float(v.abs().max())), k
    assert tb.infer_dims_from_stats(st) == (g.dims_raw, g.dims_diff)


# ------------------------------------------------------------------------------------------- fp32 GEMM
def _gemm_ref(A, W, taps, dil, T, bias, res, act):
    M, K = A.shape
    N = W.shape[0]
    A64, W64 = A.double(), W.double()
    out = torch.zeros(M, N, dtype=torch.float64, device=A.device)
    t = torch.arange(M, device=A.device) % T
    for j in range(taps):
        sh = (j - taps // 2) * dil if taps > 1 else 0
        src = torch.arange(M, device=A.device) + sh
        ok = (t + sh >= 0) & (t + sh < T) if taps > 1 else torch.ones(M, dtype=torch.bool, device=A.device)
        Aj = torch.where(ok[:, None], A64[src.clamp(0, M - 1)], torch.zeros((), dtype=torch.float64, device=A.device))
        out += Aj @ W64[:, j * K:(j + 1) * K].T
    if bias is not None:
        out += bias.double()
    if res is not None:
        out += res.double()
    if act == 1:
        out = torch.nn.functional.gelu(out)
    elif act == 2:
        out = torch.relu(out)
    return out


@pytest.mark.parametrize("M,N,K,taps,dil,T,act,use_bias,use_res", [
    (200, 256, 256, 1, 1, 1, 0, False, False),
    (96, 256, 9, 1, 1, 32, 0, False, False),        # unaligned stem K
    (128, 256, 207, 1, 1, 32, 0, False, False),
    (330, 768, 256, 1, 1, 1, 0, True, False),
    (330, 256, 1024, 1, 1, 1, 0, True, True),
    (330, 1024, 256, 1, 1, 1, 2, True, False),
    (192, 256, 256, 5, 1, 32, 1, False, False),
    (192, 256, 256, 5, 8, 32, 1, False, True),
    (160, 256, 256, 5, 4, 20, 1, False, True),      # T not a power of two
])
def test_gemm_f32(M, N, K, taps, dil, T, act, use_bias, use_res):
    lib = _lib.load()
    h = tb.scoring.util_handle(DEV)
    gen = torch.Generator(device=DEV).manual_seed(M + N + K)
    lda = K + 3 if K % 4 else K                      # exercise the unaligned path through a column slice
    Abig = torch.randn(M, lda, device=DEV, generator=gen)
    A = Abig[:, lda - K:]
    ldw = (taps * K + 3) // 4 * 4
    W = torch.zeros(N, ldw, device=DEV)
    W[:, :taps * K] = torch.randn(N, taps * K, device=DEV, generator=gen) / math.sqrt(K * taps)
    bias = torch.randn(N, device=DEV, generator=gen) if use_bias else None
    res = torch.randn(M, N, device=DEV, generator=gen) if use_res else None
    Cout = torch.empty(M, N, device=DEV)
    rc = lib.tag_debug_gemm_f32(h, A.data_ptr(), lda, W.data_ptr(), ldw, M, N, K, taps, dil, T, _lib.ptr(bias), _lib.ptr(res),
                                Cout.data_ptr(), act, torch.cuda.current_stream().cuda_stream)
    _lib.check(h, rc, "tag_debug_gemm_f32")
    torch.cuda.synchronize()
    ref = _gemm_ref(A, W[:, :taps * K], taps, dil, T, bias, res, act)
    err = (Cout.double() - ref).abs().max().item()
    assert err < 2e-5 * max(1.0, ref.abs().max().item()), err


# ------------------------------------------------------------------------------------------- encoder, fp32 mode
def _model(g, precision, max_windows=64):
    m = tb.HumanActionScorer(g.dims_raw, g.dims_diff, precision=precision, max_windows=max_windows)
    m.load_state_dict(g.sd, strict=True)
    return m.to(DEV).eval()


def _gpu_features(g, model, windows, videos):
    dv = tb.DeviceVideos(videos, g.mods, DEV)
    ds = tb.WindowDataset(windows, g.clip_len, stats=g.stats(), videos=dv, dims_map_raw=g.dims_raw,
                          dims_map_diff=g.dims_diff)
    return tb.extract_window_features(model, ds.batches(16), DEV)


def _check_scores(g, model, tol_embed, tol_score):
    feats = _gpu_features(g, model, g.gen_windows(), g.gen)
    assert feats["vid_names"] == g.meta["vid_names"] and feats["cls_names"] == g.meta["cls_names"]
    e_seq = max_abs(feats["seq_embeds"], g.npz["seq_embeds"])
    e_frm = max_abs(feats["frame_embeds"][g.meta["frame_windows"]], g.npz["frame_embeds_sel"])
    # centroids through the drop-in build_train_centroids_subset over the real train windows
    dvr = tb.DeviceVideos(g.real, g.mods, DEV)
    dsr = tb.WindowDataset(g.real_windows(), g.clip_len, stats=g.stats(), videos=dvr, dims_map_raw=g.dims_raw,
                           dims_map_diff=g.dims_diff)
    cen, counts = tb.build_train_centroids_subset(model, dsr.batches(64), g.label_dict, DEV)
    model.eval()
    assert max_abs(counts.cpu(), g.npz["counts"]) == 0
    e_cen = max_abs(cen.cpu(), g.npz["centroids"])
    ac = tb.compute_action_consistency_scores(feats, cen, g.label_dict)
    tc = tb.compute_temporal_coherence_scores(feats)
    assert set(ac) == set(g.meta["ac"]) and set(tc) == set(g.meta["tc"])
    e_ac = max(abs(ac[k] - g.meta["ac"][k]) / g.meta["ac"][k] for k in ac)
    e_tc = max(abs(tc[k] - g.meta["tc"][k]) / g.meta["tc"][k] for k in tc)
    print(f"[{g.tag} {model.precision}] seq {e_seq:.2e} frame {e_frm:.2e} centroid {e_cen:.2e} AC rel {e_ac:.2e} TC rel {e_tc:.2e}")
    assert e_seq < tol_embed and e_frm < tol_embed and e_cen < tol_embed
    assert e_ac < tol_score and e_tc < tol_score
    return feats, cen


@pytest.mark.parametrize("tag", ["m5_t32", "m7_t256"])
def test_encoder_fp32_matches_reference(tag):
    g = golden_case(tag)
    model = _model(g, "fp32", max_windows=16)
    _check_scores(g, model, tol_embed=2e-5, tol_score=1e-4)
    assert model.launch_count() > 0


def test_encoder_fp32_tokens_and_chunking():
    """tokens output (raw, un-normalised) and internal chunking (max_windows smaller than the batch)."""
    g = golden_case("m5_t32")
    wins = g.gen_windows()[:11]
    stats = g.stats()
    x = torch.stack([O.window_features(g.gen.video(v), s, 32, stats, g.mods)[0] for v, s in wins], 0)
    with torch.no_grad():
        rs, rf, rt = O.encoder_forward(g.sd, x, g.dims_raw, g.dims_diff)
    model = _model(g, "fp32", max_windows=4)       # 11 windows -> 3 internal passes
    s, f, t = model(x.to(DEV))
    assert max_abs(s.cpu(), rs) < 2e-5 and max_abs(f.cpu(), rf) < 2e-5
    assert max_abs(t.cpu(), rt) < 1e-4 * float(rt.abs().max())
    # reference error behaviour (model.py:113-117) and shape checks
    with pytest.raises(ValueError):
        tb.HumanActionScorer({"vit": 4}, {"pose": 4})
    with pytest.raises(ValueError):
        model(x[:, :, :100].to(DEV))
    model.train()
    with pytest.raises(tb.TagError):
        model(x.to(DEV))


# ------------------------------------------------------------------------------------------- K3 / K4
def test_centroid_and_score_kernels_against_oracle():
    gen = torch.Generator().manual_seed(11)
    N, Cn = 5000, 10
    z = torch.nn.functional.normalize(torch.randn(N, 256, generator=gen), dim=-1)
    y = torch.randint(0, Cn, (N,), generator=gen)
    y[:700] = 3                                      # long run + random labels
    sc = torch.zeros(Cn, 257, device=DEV)
    tb.centroid_accumulate(z.to(DEV), y.to(DEV, torch.int32), sc)
    cen, cnt = tb.centroid_finalize(sc)
    rc, rn = O.build_centroids(z, y, Cn)
    assert max_abs(cnt.cpu(), rn) == 0
    assert max_abs(cen.cpu(), rc) < 1e-6
    # labels outside [0, C) are ignored; empty classes normalise to zero (counts.clamp_min(1))
    y2 = y.clone(); y2[::3] = -1; y2[1::3] = 99; y2[y2 == 5] = 6
    sc2 = torch.zeros(Cn, 257, device=DEV)
    tb.centroid_accumulate(z.to(DEV), y2.to(DEV, torch.int32), sc2)
    cen2, cnt2 = tb.centroid_finalize(sc2)
    keep = (y2 >= 0) & (y2 < Cn)
    rc2, rn2 = O.build_centroids(z[keep], y2[keep], Cn)
    assert max_abs(cnt2.cpu(), rn2) == 0 and float(cnt2[5]) == 0 and float(cen2[5].abs().max()) == 0
    assert max_abs(cen2.cpu(), rc2) < 1e-6
    # AC / TC dict functions: ragged, interleaved windows, unknown class, class index >= len(centroids)
    names = [f"vid{(i * 7) % 23}.npz" for i in range(300)]
    classes = [tb.ACTION_CLASSES[((i * 7) % 23) % 10] if (i * 7) % 23 != 5 else "Unknown" for i in range(300)]
    fe = torch.nn.functional.normalize(torch.randn(300, 9, 256, generator=gen), dim=-1)
    features = {"seq_embeds": z[:300], "frame_embeds": fe, "vid_names": names, "cls_names": classes}
    label_dict = {c: i for i, c in enumerate(tb.ACTION_CLASSES)}
    ac = tb.compute_action_consistency_scores(features, cen[:8], label_dict)
    tc = tb.compute_temporal_coherence_scores(features)
    rac = O.action_consistency_scores(features, rc[:8], label_dict)
    rtc = O.temporal_coherence_scores(features)
    assert list(ac) == list(rac) and list(tc) == list(rtc)
    assert max(abs(ac[k] - rac[k]) for k in ac) < 2e-6
    assert max(abs(tc[k] - rtc[k]) / rtc[k] for k in tc) < 2e-6
    # a window with a single frame has no TC (eval.py:220)
    f1 = {"frame_embeds": fe[:4, :2], "vid_names": names[:4]}
    assert tb.compute_temporal_coherence_scores(f1) == {}


def test_tcl_forward_matches_reference():
    g = golden_case("m5_t32")
    z = torch.from_numpy(g.npz["seq_embeds"])
    y = torch.arange(z.shape[0]) % 4
    got = float(tb.TCL()(z.to(DEV), y.to(DEV)))
    assert abs(got - g.meta["tcl"]) < 2e-4 * abs(g.meta["tcl"])
    assert abs(got - float(O.tcl_loss(z, y))) < 2e-4 * abs(g.meta["tcl"])


def test_empty_inputs_are_noops():
    lib = _lib.load()
    h = tb.scoring.util_handle(DEV)
    s = torch.cuda.current_stream().cuda_stream
    assert lib.tag_centroid_accumulate(h, None, None, 0, 10, None, s) == 0
    assert lib.tag_score(h, None, None, None, None, None, 10, 0, None, None, s) == 0
    assert lib.tag_stats_accumulate(h, None, 0, 8, None, None, s) == 0
    assert lib.tag_score(h, None, None, None, None, None, 10, 5, None, None, s) != 0     # NULL with work -> error code
    assert b"seg_offsets" in lib.tag_last_error(h)


# ------------------------------------------------------------------------------------------- fused pipeline
def test_fused_pipeline_matches_reference_scores_fp32():
    g = golden_case("m5_t32")
    model = _model(g, "fp32", max_windows=16)
    scorer = tb.TagScorer(model, g.stats(), clip_len=g.clip_len, stride=g.stride, device=DEV)
    cen = torch.from_numpy(g.npz["centroids"]).to(DEV)
    ac, tc = scorer.score_host(g.gen.pin(), cen)
    d = scorer.scores_dict(g.gen, ac, tc)
    for k, v in g.meta["ac"].items():
        assert abs(d[k]["ac"] - v) < 1e-4 * v
        assert abs(d[k]["tc"] - g.meta["tc"][k]) < 1e-4 * g.meta["tc"][k]
    # centroid build over the real train split (window order differs from the reference: sums commute)
    idx = [g.real_index[n] for n in g.meta["train_names"]]
    dvr = scorer.to_device(g.real.select(idx))
    cen2, cnt2 = scorer.build_centroids(dvr, len(g.label_dict))
    assert max_abs(cnt2.cpu(), g.npz["counts"]) == 0
    assert max_abs(cen2.cpu(), g.npz["centroids"]) < 2e-5


def test_full_size_properties_fp32():
    """BASELINE config-1 scale (64 videos x 32 frames) through size-independent properties: scores do not
    depend on batch composition / order, and the allreduce algebra (sum of shard sums == full sums)."""
    dims_raw, dims_diff = tb.dims_maps(False)
    sd = tb.make_state_dict(dims_raw, dims_diff, seed=3)
    model = tb.HumanActionScorer(dims_raw, dims_diff, precision="fp32", max_windows=24)
    model.load_state_dict(sd)
    model.to(DEV).eval()
    real = tb.make_videos(40, 48, seed=1338)
    stats = tb.compute_stats_from_videos(real, dims_raw, dims_diff, DEV)
    scorer = tb.TagScorer(model, stats, 32, 8, DEV)
    dvr = scorer.to_device(real)
    full = scorer.centroid_sums(dvr, 10)
    parts = [scorer.centroid_sums(scorer.to_device(real.select(range(*tb.shard_range(40, r, 4)))), 10) for r in range(4)]
    assert max_abs(sum(parts).cpu(), full.cpu()) < 1e-4
    cen, cnt = tb.centroid_finalize(full)
    assert float(cnt.sum()) == 40 * 3
    gen = tb.make_videos(64, 32, seed=1337)
    ac, tc = scorer.score(scorer.to_device(gen), cen)
    perm = list(reversed(range(64)))
    ac2, tc2 = scorer.score(scorer.to_device(gen.select(perm)), cen)
    assert max_abs(ac.cpu()[perm], ac2.cpu()) < 1e-5 and max_abs(tc.cpu()[perm], tc2.cpu()) < 1e-5
    assert bool(torch.isfinite(ac).all()) and bool(torch.isfinite(tc).all())
    assert float(ac.min()) > 0 and float(ac.max()) < 2.0 and float(tc.min()) > 0
