"""GPU tests of the tcgen05/TMA tensor-core path (pytest -m gpu): the GEMM kernel in isolation against a
float64 reference of the same fp16-rounded operands, then the whole encoder in 'fp16_tc' mode against
the reference goldens at the north-star tolerance (1e-3 relative on AC / TC / centroids)."""
import math

import numpy as np
import pytest
import torch

import tag_b200 as tb
from tag_b200 import _lib
from helpers import golden_case, oracle, max_abs
from test_gpu_kernels import _gemm_ref, _model, _check_scores

pytestmark = pytest.mark.gpu
O = oracle()
DEV = "cuda:0"


def _diag(got, ref, name):
    """Human-readable mismatch report (layout bugs show up as structured error patterns)."""
    err = (got.double() - ref).abs()
    scale = max(1.0, ref.abs().max().item())
    bad = err > 5e-3 * scale
    lines = [f"{name}: max err {err.max().item():.3e} (scale {scale:.3e}), mismatched {bad.float().mean().item() * 100:.2f}%"]
    if bad.any():
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        lines.append(f"  bad rows: n={rows.numel()} first={rows[:12].tolist()} last={rows[-4:].tolist()}")
        lines.append(f"  bad cols: n={cols.numel()} first={cols[:12].tolist()} last={cols[-4:].tolist()}")
        r0, c0 = rows[0].item(), cols[0].item()
        lines.append(f"  got[{r0},{c0}:{c0 + 6}] = {got[r0, c0:c0 + 6].tolist()}")
        lines.append(f"  ref[{r0},{c0}:{c0 + 6}] = {ref[r0, c0:c0 + 6].tolist()}")
        # is the output a permutation of the reference rows? (swizzle / row-mapping bugs)
        g0 = got[r0].double()
        d = (ref - g0[None, :]).abs().max(1).values
        lines.append(f"  got row {r0} is closest to ref row {d.argmin().item()} (err {d.min().item():.3e})")
    return "\n".join(lines)


def _run_tc(M, N, K, taps, dil, T, act, use_bias, res_kind, out_kind, seed=0):
    lib = _lib.load()
    h = tb.scoring.util_handle(DEV)
    gen = torch.Generator(device=DEV).manual_seed(seed + M + N + K)
    A = torch.randn(M, K, device=DEV, generator=gen).half()
    W = (torch.randn(N, taps * K, device=DEV, generator=gen) / math.sqrt(K * taps)).half()
    bias = torch.randn(N, device=DEV, generator=gen) if use_bias else None
    res16 = torch.randn(M, N, device=DEV, generator=gen).half() if res_kind == 16 else None
    res32 = torch.randn(M, N, device=DEV, generator=gen) if res_kind == 32 else None
    C16 = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float16) if out_kind in (16, 48) else None
    C32 = torch.full((M, N), float("nan"), device=DEV) if out_kind in (32, 48) else None
    rc = lib.tag_debug_gemm_tc(h, A.data_ptr(), K, W.data_ptr(), M, N, K, taps, dil, T, _lib.ptr(bias), _lib.ptr(res16),
                               _lib.ptr(res32), _lib.ptr(C16), _lib.ptr(C32), act, None, None, None, None,
                               torch.cuda.current_stream().cuda_stream)
    _lib.check(h, rc, "tag_debug_gemm_tc")
    torch.cuda.synchronize()
    res = res16.float() if res16 is not None else res32
    ref = _gemm_ref(A.float(), W.float(), taps, dil, T, bias, res, act)
    return C16, C32, ref


CASES = [
    # M, N, K, taps, dil, T, act, bias, res, out
    (128, 256, 64, 1, 1, 1, 0, False, 0, 32),          # one tile, one k-block
    (128, 256, 256, 1, 1, 1, 0, False, 0, 32),         # 4 k-blocks (ring wrap at 4 stages)
    (300, 256, 1024, 1, 1, 1, 0, True, 32, 32),        # ragged M, 16 k-blocks, bias + fp32 residual (FFN2 / out-proj form)
    (330, 768, 256, 1, 1, 1, 0, True, 0, 16),          # 3 n-tiles (QKV form)
    (330, 1024, 256, 1, 1, 1, 2, True, 0, 16),         # FFN1 form, ReLU
    (128 * 151, 256, 256, 1, 1, 1, 0, False, 0, 16),   # more tiles than SMs: persistent loop + both TMEM buffers
    (256, 256, 256, 5, 1, 32, 1, False, 0, 16),        # conv1 form: 5 taps, GELU
    (256, 256, 256, 5, 2, 32, 1, False, 16, 16),       # conv2 form: residual + GELU
    (256, 256, 256, 5, 8, 32, 1, False, 16, 16),       # dilation 8: taps +-16 frames fall outside half the window
    (384, 256, 256, 5, 4, 64, 1, False, 16, 16),       # T = 64 (2 windows / tile)
    (96, 256, 256, 5, 4, 16, 1, False, 0, 16),         # T = 16, window count not filling the tile
    (512, 256, 256, 5, 8, 256, 1, False, 16, 16),      # T = 256 (2 tiles / window)
    (128 * 37, 256, 256, 5, 2, 32, 0, False, 0, 16),   # many conv tiles
    (128 * 37 + 40, 256, 256, 1, 1, 1, 0, True, 32, 32),   # ragged last tile, fp32 residual + fp32 output
    (77, 512, 128, 1, 1, 1, 2, True, 16, 16),          # single short tile (1-CTA kernel), fp16 residual, 2 n-tiles
]


@pytest.mark.parametrize("case", CASES, ids=[f"M{c[0]}_N{c[1]}_K{c[2]}_t{c[3]}d{c[4]}T{c[5]}" for c in CASES])
def test_gemm_tc(case):
    M, N, K, taps, dil, T, act, use_bias, res_kind, out_kind = case
    C16, C32, ref = _run_tc(*case)
    scale = max(1.0, ref.abs().max().item())
    if C32 is not None:
        err = (C32.double() - ref).abs().max().item()
        assert err < 2e-4 * scale, _diag(C32, ref, "C32")
    if C16 is not None:
        err = (C16.double() - ref).abs().max().item()
        assert err < 2e-3 * scale, _diag(C16, ref, "C16")


def test_gemm_tc_rejects_unsupported_shapes():
    lib = _lib.load()
    h = tb.scoring.util_handle(DEV)
    A = torch.zeros(96, 256, device=DEV, dtype=torch.float16)
    W = torch.zeros(256, 1280, device=DEV, dtype=torch.float16)
    Cc = torch.zeros(96, 256, device=DEV, dtype=torch.float16)
    s = torch.cuda.current_stream().cuda_stream
    # T = 48 neither divides 128 nor is a multiple of it
    assert lib.tag_debug_gemm_tc(h, A.data_ptr(), 256, W.data_ptr(), 96, 256, 256, 5, 1, 48, None, None, None, Cc.data_ptr(), None, 0, None, None, None, None, s) != 0
    assert b"T dividing 128" in lib.tag_last_error(h)
    assert lib.tag_debug_gemm_tc(h, A.data_ptr(), 256, W.data_ptr(), 96, 100, 256, 1, 1, 1, None, None, None, Cc.data_ptr(), None, 0, None, None, None, None, s) != 0
    # exactly one output
    C32 = torch.zeros(96, 256, device=DEV)
    assert lib.tag_debug_gemm_tc(h, A.data_ptr(), 256, W.data_ptr(), 96, 256, 256, 1, 1, 1, None, None, None, Cc.data_ptr(), C32.data_ptr(), 0, None, None, None, None, s) != 0
    # fused GroupNorm only for convs whose tile owns whole windows
    g = torch.ones(256, device=DEV)
    assert lib.tag_debug_gemm_tc(h, A.data_ptr(), 256, W.data_ptr(), 96, 256, 256, 1, 1, 1, None, None, None, Cc.data_ptr(), None, 1, g.data_ptr(), g.data_ptr(), None, None, s) != 0


@pytest.mark.parametrize("W_,T,dil", [(8, 32, 1), (37, 32, 8), (12, 16, 2), (5, 64, 4), (3, 128, 8), (150 * 4, 32, 2), (9, 8, 1),
                                      (1, 256, 8), (7, 256, 2), (160, 256, 4)])   # T = 256: statistics cross the CTA pair
def test_gemm_tc_fused_groupnorm(W_, T, dil):
    """conv2 form of reference TemporalConvBlock (model.py:38-40): GroupNorm(1,256)(GELU(conv(y) + res)) in one kernel,
    written in place over the residual buffer as the encoder does."""
    lib = _lib.load()
    h = tb.scoring.util_handle(DEV)
    M, N, K, taps = W_ * T, 256, 256, 5
    gen = torch.Generator(device=DEV).manual_seed(W_ * 1000 + T)
    A = torch.randn(M, K, device=DEV, generator=gen).half()
    Wt = (torch.randn(N, taps * K, device=DEV, generator=gen) / math.sqrt(K * taps)).half()
    res = torch.randn(M, N, device=DEV, generator=gen).half()
    gamma = 1.0 + 0.1 * torch.randn(N, device=DEV, generator=gen)
    beta = 0.05 * torch.randn(N, device=DEV, generator=gen)
    z = _gemm_ref(A.float(), Wt.float(), taps, dil, T, None, res.float(), 1)              # GELU(conv + res), float64
    zz = z.reshape(W_, T * N)
    ref = ((zz - zz.mean(1, keepdim=True)) / torch.sqrt(zz.var(1, unbiased=False, keepdim=True) + 1e-5)).reshape(M, N)
    ref = ref * gamma.double() + beta.double()
    buf = res.clone()                                                                       # in place: C16 == res16
    rc = lib.tag_debug_gemm_tc(h, A.data_ptr(), K, Wt.data_ptr(), M, N, K, taps, dil, T, None, buf.data_ptr(), None,
                               buf.data_ptr(), None, 1, gamma.data_ptr(), beta.data_ptr(), None, None, torch.cuda.current_stream().cuda_stream)
    _lib.check(h, rc, "tag_debug_gemm_tc(gn)")
    torch.cuda.synchronize()
    err = (buf.double() - ref).abs().max().item()
    assert err < 4e-3 * max(1.0, ref.abs().max().item()), _diag(buf, ref, f"GN W={W_} T={T}")


@pytest.mark.parametrize("tag", ["m5_t32", "m7_t256"])
def test_encoder_tc_matches_reference_within_north_star_tolerance(tag):
    g = golden_case(tag)
    model = _model(g, "fp16_tc", max_windows=16)
    # north-star: per-video scores and centroids within 1e-3 relative of the reference's fp32 path
    _check_scores(g, model, tol_embed=2e-3, tol_score=1e-3)


def test_encoder_tc_vs_fp32_mode_full_batch():
    """Larger batch than the goldens (BASELINE config-1 size: 64 windows x 32 frames): tensor-core mode vs
    the fp32 mode of the same library on the same inputs."""
    g = golden_case("m5_t32")
    vb = tb.make_videos(64, 32, seed=1337 + 1)
    stats = g.stats()
    res = {}
    for prec in ("fp32", "fp16_tc"):
        model = _model(g, prec, max_windows=64)
        scorer = tb.TagScorer(model, stats, 32, 8, DEV)
        enc = scorer.encode(scorer.to_device(vb), want_frames=True)
        res[prec] = (enc["seq"].cpu(), enc["tc_window"].cpu(), enc["frames"].cpu())
    e_seq = max_abs(res["fp32"][0], res["fp16_tc"][0])
    e_tc = float(((res["fp32"][1] - res["fp16_tc"][1]).abs() / res["fp32"][1]).max())
    print(f"tc-vs-fp32: seq {e_seq:.2e} tc rel {e_tc:.2e}")
    assert e_seq < 2e-3 and e_tc < 1e-3


@pytest.mark.parametrize("T", [8, 16, 64, 128])
def test_encoder_tc_vs_fp32_mode_other_clip_lengths(T):
    """Clip lengths other than the goldens' 32 / 256: S = T + 1 = 9, 17 (one-warp attention), 65, 129 (flash-style attention
    over key chunks), fused GroupNorm with 16 / 8 / 2 / 1 windows per tile. Tensor-core mode vs the fp32 mode of the same
    library on the same windows (the fp32 mode is pinned to the reference by the golden tests)."""
    g = golden_case("m5_t32")
    vb = tb.make_videos(24, T + 5, seed=4242 + T)
    stats = g.stats()
    res = {}
    for prec in ("fp32", "fp16_tc"):
        model = _model(g, prec, max_windows=32)
        scorer = tb.TagScorer(model, stats, T, max(1, T // 4), DEV)
        enc = scorer.encode(scorer.to_device(vb), want_frames=True)
        res[prec] = (enc["seq"].cpu(), enc["tc_window"].cpu(), enc["frames"].cpu())
    assert res["fp32"][0].shape[0] >= 24 and res["fp32"][2].shape[1] == T + 1
    e_seq = max_abs(res["fp32"][0], res["fp16_tc"][0])
    e_fr = max_abs(res["fp32"][2], res["fp16_tc"][2])
    e_tc = float(((res["fp32"][1] - res["fp16_tc"][1]).abs() / res["fp32"][1]).max())
    print(f"T={T}: seq {e_seq:.2e} frames {e_fr:.2e} tc rel {e_tc:.2e}")
    assert e_seq < 2e-3 and e_fr < 4e-3 and e_tc < 1e-3


@pytest.mark.parametrize("L,T,stride", [(64, 32, 8), (37, 32, 8), (32, 32, 8), (64, 16, 4), (150, 128, 16), (70, 64, 3)])
def test_encode_clips_frame_table_matches_window_path(L, T, stride):
    """Clips of equal length: tag_encode_clips (tensor-core mode: features built once per source frame, windows gathered
    by the stem GEMMs, first frame of every window forced to the zero-motion row) against tag_encode_windows on the
    explicit window table, and against the fp32 mode."""
    g = golden_case("m5_t32")
    vb = tb.make_videos(21, L, seed=777 + L + T)
    stats = g.stats()
    out = {}
    for prec in ("fp16_tc", "fp32"):
        model = _model(g, prec, max_windows=40)           # several passes, the last one ragged
        scorer = tb.TagScorer(model, stats, T, stride, DEV)
        dv = scorer.to_device(vb)
        for clips in (None, False):
            enc = scorer.encode(dv, want_frames=True, use_clips=clips)
            out[(prec, clips)] = (enc["seq"].cpu(), enc["tc_window"].cpu(), enc["frames"].cpu(), int(enc["flags"].item()))
    n_win = 21 * ((L - T) // stride + 1)
    for k, v in out.items():
        assert v[0].shape[0] == n_win and v[3] == 0, k
    # fp32 mode: both entry points run the same kernels on the same windows
    assert torch.equal(out[("fp32", None)][0], out[("fp32", False)][0]) and torch.equal(out[("fp32", None)][2], out[("fp32", False)][2])
    a, b = out[("fp16_tc", None)], out[("fp16_tc", False)]
    e_seq, e_fr = max_abs(a[0], b[0]), max_abs(a[2], b[2])
    e_tc = float(((a[1] - b[1]).abs() / b[1]).max())
    print(f"L={L} T={T} stride={stride}: frame table vs window path: seq {e_seq:.2e} frames {e_fr:.2e} tc rel {e_tc:.2e}")
    # same fp16 operands everywhere; only the zero-motion rows are summed in a different order (fp32)
    assert e_seq < 2e-4 and e_fr < 5e-4 and e_tc < 2e-4
    r = out[("fp32", False)]
    assert max_abs(a[0], r[0]) < 2e-3 and max_abs(a[2], r[2]) < 4e-3 and float(((a[1] - r[1]).abs() / r[1]).max()) < 1e-3


def test_full_size_properties_tc():
    """One full encoder pass of BASELINE config 2 and a bit (2,700 videos x 64 frames -> 13,500 windows: the 13,024-window
    pass boundary falls inside the batch) through size-independent properties: per-video scores do not depend on the
    order of the videos, on the pass size, or on what else is in the batch; duplicated videos score identically; and a
    sample of the videos agrees with the CPU oracle to the north-star tolerance."""
    dims_raw, dims_diff = tb.dims_maps(False)
    sd = tb.make_state_dict(dims_raw, dims_diff, seed=0)
    real = tb.make_videos(100, 64, seed=1340, device=DEV)
    stats = tb.compute_stats_from_videos(real, dims_raw, dims_diff, DEV)
    V = 2700
    gen = tb.make_videos(V, 64, seed=1339, device=DEV)
    res = {}
    for mw in (13024, 4096):
        model = tb.HumanActionScorer(dims_raw, dims_diff, precision="fp16_tc", max_windows=mw)
        model.load_state_dict(sd)
        model.to(DEV).eval()
        scorer = tb.TagScorer(model, stats, 32, 8, DEV)
        if mw == 13024:
            cen, cnt = scorer.build_centroids(scorer.to_device(real), 10)
            assert float(cnt.sum()) == 100 * 5
        ac, tc = scorer.score(scorer.to_device(gen), cen)
        res[mw] = (ac.cpu(), tc.cpu())
        assert int(scorer.last_flags.item()) == 0
    ac, tc = res[13024]
    assert bool(torch.isfinite(ac).all()) and bool(torch.isfinite(tc).all())
    assert float(ac.min()) > 0 and float(ac.max()) < 2.0 and float(tc.min()) > 0
    # pass size: rows are processed independently, so the split into passes must not change a single bit
    assert torch.equal(res[4096][0], ac) and torch.equal(res[4096][1], tc)
    # order + batch composition + duplicates: reversed order, every 9th video, and video 7 three times
    g = torch.Generator().manual_seed(0)
    perm = torch.randperm(V, generator=g).tolist()
    sub = perm[::9] + [7, 7, 7]
    ac2, tc2 = scorer.score(scorer.to_device(gen.select(sub)), cen)
    assert torch.equal(ac2.cpu(), ac[sub]) and torch.equal(tc2.cpu(), tc[sub])
    # a sample against the oracle (fp32 CPU restatement of the reference)
    pick = [0, 1234, V - 1]
    host = gen.select(pick).to("cpu")
    ostats = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in (stats if isinstance(stats, dict) else vars(stats)).items()}
    label_dict = {c: i for i, c in enumerate(tb.ACTION_CLASSES)}
    oac, otc, _ = O.score_videos([host.video(i) for i in range(len(pick))], host.names, [host.cls_name(i) for i in range(len(pick))],
                                 sd, dims_raw, dims_diff, ostats, cen.cpu(), label_dict, clip_len=32, stride=8)
    for i, v in enumerate(pick):
        key = host.names[i].rsplit(".", 1)[0]
        assert abs(oac[key] - float(ac[v])) / oac[key] < 1e-3, (v, oac[key], float(ac[v]))
        assert abs(otc[key] - float(tc[v])) / otc[key] < 1e-3, (v, otc[key], float(tc[v]))


def test_fused_pipeline_tc_scores():
    g = golden_case("m5_t32")
    model = _model(g, "fp16_tc", max_windows=16)
    scorer = tb.TagScorer(model, g.stats(), clip_len=g.clip_len, stride=g.stride, device=DEV)
    cen = torch.from_numpy(g.npz["centroids"]).to(DEV)
    ac, tc = scorer.score_host(g.gen.pin(), cen)
    d = scorer.scores_dict(g.gen, ac, tc)
    e_ac = max(abs(d[k]["ac"] - v) / v for k, v in g.meta["ac"].items())
    e_tc = max(abs(d[k]["tc"] - v) / v for k, v in g.meta["tc"].items())
    print(f"fused tc pipeline: AC rel {e_ac:.2e} TC rel {e_tc:.2e}")
    assert e_ac < 1e-3 and e_tc < 1e-3
    assert int(scorer.last_flags.item()) == 0


@pytest.mark.parametrize("W_,T,dil", [(8, 32, 1), (9, 32, 2), (37, 32, 4), (150 * 4 + 3, 32, 1), (150 * 4, 32, 4), (12, 16, 2), (40, 16, 1),
                                      (5, 64, 2), (21, 64, 4), (3, 128, 1), (11, 128, 8), (33, 8, 1), (2500, 32, 2), (9, 64, 8), (40, 128, 4), (37, 32, 8), (610, 32, 8), (21, 128, 2)])
def test_tcn_block_fused_is_bit_identical_to_the_two_kernel_path(W_, T, dil):
    """One fused TemporalConvBlock kernel (tcn_block_tc.cu; model.py:22-41) against the two GEMM launches it replaces — conv1 + GELU into a
    buffer, conv2 + residual + GELU + GroupNorm in place — on the same inputs. GELU(conv1) is rounded to fp16 in both and every
    accumulation runs in the same order, so the outputs must be bit-identical (and the two-kernel path is held to the float64
    reference by test_gemm_tc / test_gemm_tc_fused_groupnorm). Covers partial last tiles, 1 .. 16 windows per tile, more tiles than SMs."""
    lib = _lib.load()
    h = tb.scoring.util_handle(DEV)
    M, N, K, taps = W_ * T, 256, 256, 5
    gen = torch.Generator(device=DEV).manual_seed(W_ * 100 + T + dil)
    x = torch.randn(M, K, device=DEV, generator=gen).half()
    W1 = (torch.randn(N, taps * K, device=DEV, generator=gen) / math.sqrt(K * taps)).half()
    W2 = (torch.randn(N, taps * K, device=DEV, generator=gen) / math.sqrt(K * taps)).half()
    gamma = 1.0 + 0.1 * torch.randn(N, device=DEV, generator=gen)
    beta = 0.05 * torch.randn(N, device=DEV, generator=gen)
    s = torch.cuda.current_stream().cuda_stream
    y1 = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float16)
    two = x.clone()
    _lib.check(h, lib.tag_debug_gemm_tc(h, two.data_ptr(), K, W1.data_ptr(), M, N, K, taps, dil, T, None, None, None, y1.data_ptr(), None, 1,
                                        None, None, None, None, s), "conv1")
    _lib.check(h, lib.tag_debug_gemm_tc(h, y1.data_ptr(), K, W2.data_ptr(), M, N, K, taps, dil, T, None, two.data_ptr(), None, two.data_ptr(),
                                        None, 1, gamma.data_ptr(), beta.data_ptr(), None, None, s), "conv2+gn")
    one = x.clone()
    rc = lib.tag_debug_tcn_block(h, one.data_ptr(), M, T, dil, W1.data_ptr(), W2.data_ptr(), gamma.data_ptr(), beta.data_ptr(), s)
    if M <= 128:
        assert rc != 0                                      # a single tile has no CTA pair: the schedule keeps the two-kernel path
        return
    _lib.check(h, rc, "tag_debug_tcn_block")
    torch.cuda.synchronize()
    assert torch.isfinite(one.float()).all()
    assert torch.equal(one, two), _diag(one.float(), two.double(), f"tcn block W={W_} T={T} dil={dil}")


def test_tcn_block_unsupported_shapes_fall_back():
    lib = _lib.load()
    h = tb.scoring.util_handle(DEV)
    x = torch.zeros(64 * 32, 256, device=DEV, dtype=torch.float16)
    W = torch.zeros(256, 1280, device=DEV, dtype=torch.float16)
    g = torch.ones(256, device=DEV)
    s = torch.cuda.current_stream().cuda_stream
    for T, dil in ((32, 16), (16, 8), (256, 1), (24, 1)):            # halo tiles too large for shared memory; window larger than a tile; not a power of two
        assert lib.tag_debug_tcn_block(h, x.data_ptr(), x.shape[0] // T * T, T, dil, W.data_ptr(), W.data_ptr(), g.data_ptr(), g.data_ptr(), s) != 0


def test_gemm_tc_shifted_load_mode_in_subprocess():
    """The conv GEMM's default activation path is the halo tile (one load per 64-channel chunk, taps as shifted descriptor
    views, (t, window)-ordered accumulator rows). The plain path — five shifted TMA loads per chunk, row-major tiles — lives on
    in the EXPERIMENTS build of the library (libtag_b200_exp.so, TAG_TC_HALO=0; the product library reads no environment
    variable) and must pass the same conv / fused-GroupNorm cases of this file against the float64 reference."""
    import os, subprocess, sys
    if os.environ.get("TAG_TC_HALO"):
        pytest.skip("already inside the child")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "video-gen-evals_b200", "libtag_b200_exp.so")):
        pytest.skip("experiments build not present")
    env = dict(os.environ, TAG_TC_HALO="0")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "run_exp.py"), "-m", "pytest", os.path.abspath(__file__), "-m", "gpu",
                        "-q", "-x", "-k", "(test_gemm_tc and _t5d) or fused_groupnorm"], env=env, capture_output=True, text=True,
                       timeout=600, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


def test_product_library_ignores_experiment_switches():
    """TAG_TC_DEBUG (which skips loads / stores in the experiments build) must do nothing to the product library."""
    import os, subprocess, sys
    if os.environ.get("TAG_TC_DEBUG"):
        pytest.skip("inside the child")
    env = dict(os.environ, TAG_TC_DEBUG="7", TAG_TC_HALO="3", TAG_K1_DEBUG="7", TAG_FRAME_TABLE="0", TAG_TC_PAIR="0", TAG_TC_TMA_STORE="0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "-x", "-k",
                        "fused_pipeline_tc_scores"], env=env, capture_output=True, text=True, timeout=600,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_score_stream_matches_resident_scores():
    """score_stream over several host batches of different sizes (prefetch across batch boundaries) returns, batch by
    batch and in order, what TagScorer.score returns for the same videos resident on the device."""
    g = golden_case("m5_t32")
    model = _model(g, "fp16_tc", max_windows=16)
    scorer = tb.TagScorer(model, g.stats(), clip_len=g.clip_len, stride=g.stride, device=DEV)
    cen = torch.from_numpy(g.npz["centroids"]).to(DEV)
    V = g.gen.n_videos
    cuts = [(0, V), (0, max(1, V // 3)), (max(1, V // 3), V), (0, 1), (0, V)]
    host = [g.gen.slice(a, b).pin() for a, b in cuts]
    want = []
    for hb in host:
        a, t = scorer.score(scorer.to_device(hb.to(DEV)), cen)
        want.append((a.cpu(), t.cpu()))
    for pieces, prefetch in ((None, 3), (1, 0), (2, 2), (3, 1), (7, 4)):
        got = list(scorer.score_stream(iter(host), cen, pieces=pieces, prefetch=prefetch))
        assert len(got) == len(host)
        for (ga, gt), (wa, wt) in zip(got, want):
            assert ga.shape == wa.shape
            # blocks whose videos happen to share one length take the frame-table path, ragged blocks the window path:
            # same fp16 operands, but the zero-motion rows are summed in a different order (see test_encode_clips_*)
            assert torch.allclose(ga, wa, rtol=5e-4, atol=0, equal_nan=True) and torch.allclose(gt, wt, rtol=5e-4, atol=0)
    assert list(scorer.score_stream(iter([]), cen)) == []


@pytest.mark.parametrize("M,K", [(330, 256), (128 * 9 + 5, 1024), (64, 256)])
def test_gemm_tc_fused_layernorm(M, K):
    """out-proj / FFN2 form of the post-norm transformer layer (model.py:145): X <- LayerNorm(A W^T + b + X), fp32 stream
    updated in place plus its fp16 copy."""
    lib = _lib.load()
    h = tb.scoring.util_handle(DEV)
    N = 256
    gen = torch.Generator(device=DEV).manual_seed(M + K)
    A = torch.randn(M, K, device=DEV, generator=gen).half()
    Wt = (torch.randn(N, K, device=DEV, generator=gen) / math.sqrt(K)).half()
    bias = 0.1 * torch.randn(N, device=DEV, generator=gen)
    X = torch.randn(M, N, device=DEV, generator=gen)
    gamma = 1.0 + 0.1 * torch.randn(N, device=DEV, generator=gen)
    beta = 0.05 * torch.randn(N, device=DEV, generator=gen)
    v = A.double() @ Wt.double().T + bias.double() + X.double()
    ref = (v - v.mean(1, keepdim=True)) / torch.sqrt(v.var(1, unbiased=False, keepdim=True) + 1e-5) * gamma.double() + beta.double()
    X16 = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float16)
    rc = lib.tag_debug_gemm_tc(h, A.data_ptr(), K, Wt.data_ptr(), M, N, K, 1, 1, 1, bias.data_ptr(), None, X.data_ptr(),
                               X16.data_ptr(), X.data_ptr(), 0, None, None, gamma.data_ptr(), beta.data_ptr(),
                               torch.cuda.current_stream().cuda_stream)
    _lib.check(h, rc, "tag_debug_gemm_tc(ln)")
    torch.cuda.synchronize()
    scale = max(1.0, ref.abs().max().item())
    assert (X.double() - ref).abs().max().item() < 2e-4 * scale, _diag(X, ref, "LN fp32")
    assert (X16.double() - ref).abs().max().item() < 2e-3 * scale, _diag(X16, ref, "LN fp16")


@pytest.mark.parametrize("M,F", [(257, 1024), (330, 1024), (128 * 9 + 5, 1024), (4096, 1024), (20000, 1024), (700, 512)])
def test_tlayer_tail_fused(M, F):
    """The fused transformer-layer tail (tlayer_tc.cu; model.py:145 post-norm layer after attention):
    x1 = LN1(x + att Wo^T + bo); x = LN2(x1 + relu(x1 W1^T + b1) W2^T + b2), against float64 with the kernel's two operand
    roundings (x1 and relu(h) enter the next GEMM as fp16) — odd tile counts, a partial last tile, many tiles per CTA pair."""
    lib = _lib.load()
    h = tb.scoring.util_handle(DEV)
    D = 256
    g = torch.Generator(device=DEV).manual_seed(M + F)
    att = torch.randn(M, D, device=DEV, generator=g).half()
    X = torch.randn(M, D, device=DEV, generator=g)
    Wo = (torch.randn(D, D, device=DEV, generator=g) / math.sqrt(D)).half()
    W1 = (torch.randn(F, D, device=DEV, generator=g) / math.sqrt(D)).half()
    W2 = (torch.randn(D, F, device=DEV, generator=g) / math.sqrt(F)).half()
    bo, b1, b2 = (0.1 * torch.randn(n, device=DEV, generator=g) for n in (D, F, D))
    g1, be1, g2, be2 = (1.0 + 0.1 * torch.randn(D, device=DEV, generator=g), 0.05 * torch.randn(D, device=DEV, generator=g),
                        1.0 + 0.1 * torch.randn(D, device=DEV, generator=g), 0.05 * torch.randn(D, device=DEV, generator=g))

    def ln(v, gam, bet):
        return (v - v.mean(1, keepdim=True)) / torch.sqrt(v.var(1, unbiased=False, keepdim=True) + 1e-5) * gam.double() + bet.double()

    x1 = ln(X.double() + att.double() @ Wo.double().T + bo.double(), g1, be1)
    hid = torch.relu(x1.float().half().double() @ W1.double().T + b1.double())
    ref = ln(x1 + hid.float().half().double() @ W2.double().T + b2.double(), g2, be2)
    Xio = X.clone()
    X16 = torch.full((M, D), float("nan"), device=DEV, dtype=torch.float16)
    rc = lib.tag_debug_tlayer_tail(h, att.data_ptr(), Xio.data_ptr(), X16.data_ptr(), M, F, Wo.data_ptr(), W1.data_ptr(), W2.data_ptr(),
                                   bo.data_ptr(), b1.data_ptr(), b2.data_ptr(), g1.data_ptr(), be1.data_ptr(), g2.data_ptr(),
                                   be2.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(h, rc, "tag_debug_tlayer_tail")
    torch.cuda.synchronize()
    scale = max(1.0, ref.abs().max().item())
    e32 = (Xio.double() - ref).abs().max().item()
    print(f"tlayer_tail M={M} F={F}: fp32 out max err {e32:.2e} (scale {scale:.2f})")
    assert e32 < 1e-3 * scale, _diag(Xio, ref, "tail fp32")
    assert (X16.double() - ref).abs().max().item() < 3e-3 * scale, _diag(X16, ref, "tail fp16")
    # unsupported shapes are refused up front, not mid-kernel
    assert lib.tag_debug_tlayer_tail(h, att.data_ptr(), Xio.data_ptr(), X16.data_ptr(), 100, F, Wo.data_ptr(), W1.data_ptr(), W2.data_ptr(),
                                     bo.data_ptr(), b1.data_ptr(), b2.data_ptr(), g1.data_ptr(), be1.data_ptr(), g2.data_ptr(),
                                     be2.data_ptr(), torch.cuda.current_stream().cuda_stream) != 0


@pytest.mark.parametrize("precision", ["fp16_tc", "fp32"])
def test_workspace_poisoning_changes_nothing(precision):
    """Every workspace buffer (activations, token stream, fp16 operand table incl. its pad columns, zero-motion rows) filled
    with NaN bytes before the call: scores must come out bit-identical to a run on the freshly allocated (zeroed) workspace —
    no kernel reads what an earlier kernel of the same call did not write. Both the window path and the frame-table path,
    ragged and uniform batches, and the drop-in forward."""
    g = golden_case("m5_t32")
    lib = _lib.load()
    model = _model(g, precision, max_windows=24)
    scorer = tb.TagScorer(model, g.stats(), clip_len=g.clip_len, stride=g.stride, device=DEV)
    cen = torch.from_numpy(g.npz["centroids"]).to(DEV)
    uni = tb.make_videos(9, 64, seed=31)
    batches = [scorer.to_device(g.gen.to(DEV)), scorer.to_device(uni.to(DEV))]
    x = torch.randn(5, 32, model.feat_dim, device=DEV)
    clean = [tuple(t.clone() for t in scorer.score(dv, cen)) for dv in batches]
    fwd_clean = tuple(t.clone() for t in model(x))
    h = model.handle(torch.device(DEV), g.clip_len)
    s = torch.cuda.current_stream().cuda_stream
    for rnd in range(2):
        for dv, want in zip(batches, clean):
            _lib.check(h, lib.tag_debug_poison_workspace(h, s), "tag_debug_poison_workspace")
            ac, tc = scorer.score(dv, cen)
            assert torch.equal(ac, want[0]) and torch.equal(tc, want[1])
        _lib.check(h, lib.tag_debug_poison_workspace(h, s), "tag_debug_poison_workspace")
        got = model(x)
        for a, b in zip(got, fwd_clean):
            assert torch.equal(a, b)


def test_score_stream_empty_batches_and_oversized_videos():
    """a batch with no videos yields an empty result in its place; a video with more windows than max_windows still
    scores (one oversized block)."""
    g = golden_case("m5_t32")
    model = _model(g, "fp16_tc", max_windows=4)
    scorer = tb.TagScorer(model, g.stats(), clip_len=g.clip_len, stride=g.stride, device=DEV)
    cen = torch.from_numpy(g.npz["centroids"]).to(DEV)
    full = g.gen.pin()                                  # 64-frame videos have 5 windows > max_windows = 4
    empty = g.gen.slice(0, 0)
    out = list(scorer.score_stream(iter([empty, full, empty, g.gen.slice(2, 5).pin(), empty]), cen))
    assert [o[0].numel() for o in out] == [0, g.gen.n_videos, 0, 3, 0]
    want = scorer.score(scorer.to_device(g.gen.to(DEV)), cen)
    assert torch.allclose(out[1][0], want[0].cpu(), rtol=5e-4, atol=0, equal_nan=True)
    assert torch.allclose(out[3][1], want[1].cpu()[2:5], rtol=5e-4, atol=0)


@pytest.mark.parametrize("precision", ["fp16_tc", "fp32"])
def test_weight_reload_keeps_the_handle(precision):
    """New parameter values (an optimiser step in the reference's training loop, train.py:502-505) re-pack the weights into the
    existing native handle — same handle pointer, same workspace — and give exactly what a freshly built model gives."""
    g = golden_case("m5_t32")
    x = torch.randn(6, 32, sum(g.dims_raw.values()) + sum(g.dims_diff.values()), device=DEV)
    model = _model(g, precision, max_windows=8)
    out0 = tuple(t.clone() for t in model(x))
    h0 = model._h.value
    sd2 = tb.make_state_dict(g.dims_raw, g.dims_diff, seed=5)
    model.load_state_dict(sd2)                          # in-place copy_: parameter versions change, storage does not
    out1 = model(x)
    assert model._h.value == h0                         # re-packed, not re-created
    fresh = tb.HumanActionScorer(g.dims_raw, g.dims_diff, precision=precision, max_windows=8)
    fresh.load_state_dict(sd2)
    fresh.to(DEV).eval()
    want = fresh(x)
    for a, b in zip(out1, want):
        assert torch.equal(a, b)
    assert not torch.equal(out1[0], out0[0])
    model.load_state_dict(g.sd)
    for a, b in zip(model(x), out0):
        assert torch.equal(a, b)
