"""GPU tests of SURVEY.md §8f row N1 (BASELINE config 5, the training-step embedding pass): TCL and
SupConWithHardNegatives forwards (losses.py:14-56) and the three hard-negative augmentations (utils.py:65-95) with the
forwards on them (train.py:511-524), against the oracle — and against the unmodified reference when oracle/_ref is there."""
import numpy as np
import pytest
import torch

import tag_b200 as tb
from helpers import golden_case, oracle, max_abs, reference
from test_gpu_kernels import _model

pytestmark = pytest.mark.gpu
O = oracle()
DEV = "cuda:0"


@pytest.mark.parametrize("B", [34, 500, 512, 1000, 4096])
def test_tcl_rows_cuda_core_and_tensor_core_paths(B):
    """B < 512 runs the warp-per-anchor kernel, B >= 512 the tcgen05 GEMM with the masked row sums in its epilogue (B = 1000:
    padded to 1024 columns); both against the float64 oracle, rows with no positive are NaN in both."""
    g = torch.Generator().manual_seed(B)
    y = torch.randint(0, 10, (B,), generator=g)
    if B == 34:
        y[5] = 77                                       # a class of one: no positives -> NaN row, as the reference
    z = torch.nn.functional.normalize(torch.randn(10, 256, generator=g)[y % 10] + 1.2 * torch.randn(B, 256, generator=g), dim=-1)
    ref = O.tcl_loss_rows(z.double(), y)
    rows = tb.TCL().loss_rows(z.to(DEV), y.to(DEV)).cpu().double()
    assert torch.equal(torch.isnan(rows), torch.isnan(ref))
    ok = ~torch.isnan(ref)
    rel = ((rows[ok] - ref[ok]).abs() / ref[ok].abs()).max().item()
    print(f"TCL B={B}: rows max rel {rel:.2e}")
    assert rel < 2e-5
    # other temperature / weights
    t2 = tb.TCL(temperature=0.07, k1=100.0, k2=2.0)
    r2 = O.tcl_loss_rows(z.double(), y, 0.07, 100.0, 2.0)
    g2 = t2.loss_rows(z.to(DEV), y.to(DEV)).cpu().double()
    assert ((g2[ok] - r2[ok]).abs() / r2[ok].abs()).max().item() < 2e-5


def test_supcon_hard_negatives_b4096():
    g = torch.Generator().manual_seed(1)
    B = 4096
    a = torch.nn.functional.normalize(torch.randn(B, 256, generator=g), dim=-1)
    p = torch.nn.functional.normalize(a + 0.3 * torch.randn(B, 256, generator=g), dim=-1)
    n = torch.nn.functional.normalize(a + 0.6 * torch.randn(B, 256, generator=g), dim=-1)
    ref = O.supcon_hard_rows(a.double(), p.double(), n.double())
    loss = tb.SupConWithHardNegatives()
    rows = loss.loss_rows(a.to(DEV), p.to(DEV), n.to(DEV)).cpu().double()
    assert bool(((rows - ref).abs() <= 2e-6 + 2e-5 * ref.abs()).all())      # fp32 logits / 0.07: ~1e-6 absolute
    assert abs(float(loss(a.to(DEV), p.to(DEV), n.to(DEV))) - float(ref.mean())) < 1e-5 * float(ref.mean())
    # the training use: anchor == positive (train.py:519-521)
    ref2 = O.supcon_hard_rows(a, a, n)
    assert max_abs(loss.loss_rows(a.to(DEV), a.to(DEV), n.to(DEV)).cpu(), ref2) < 1e-5
    ref_mod = reference()
    if ref_mod is not None:
        want = float(ref_mod.losses.SupConWithHardNegatives()(a, p, n))
        assert abs(float(loss(a.to(DEV), p.to(DEV), n.to(DEV))) - want) < 1e-5 * want
        want_t = float(ref_mod.losses.TCL()(a[:1024], torch.arange(1024) % 8))
        assert abs(float(tb.TCL()(a[:1024].to(DEV), (torch.arange(1024) % 8).to(DEV))) - want_t) < 2e-5 * abs(want_t)


def test_augmentations_b4096_are_exact_gathers():
    """4096 windows x 32 frames x 2596 features: the three augmentations are pure data movement -> bit-exact against the
    restated torch ops, with the reference's RNG consumption for the shuffle."""
    B, T, D = 4096, 32, 2596
    g = torch.Generator(device=DEV).manual_seed(2)
    x = torch.randn(B, T, D, device=DEV, generator=g)
    assert torch.equal(tb.reverse_sequence(x), torch.flip(x, dims=[1]))
    assert torch.equal(tb.get_static_window(x), x[:, :1].expand_as(x))
    torch.manual_seed(11)
    got = tb.partial_shuffle_within_window(x[:512])
    torch.manual_seed(11)
    want = O.partial_shuffle_within_window(x[:512].cpu())
    assert torch.equal(got.cpu(), want)
    ref_mod = reference()
    if ref_mod is not None:
        torch.manual_seed(12)
        a = ref_mod.utils.partial_shuffle_within_window(x[:64].cpu())
        torch.manual_seed(12)
        assert torch.equal(tb.partial_shuffle_within_window(x[:64]).cpu(), a)
        assert torch.equal(tb.reverse_sequence(x[:64]).cpu(), ref_mod.utils.reverse_sequence(x[:64].cpu()))
        assert torch.equal(tb.get_static_window(x[:64]).cpu(), ref_mod.utils.get_static_window(x[:64].cpu()))
    # odd feature width (padded to 16-byte pieces inside)
    x3 = torch.randn(3, 5, 7, device=DEV)
    assert torch.equal(tb.reverse_sequence(x3), torch.flip(x3, dims=[1]))


def test_hard_negative_step_matches_oracle():
    """compute_loss_components (train.py:511-524), forward only: embedding pass + three hard-negative passes + the four loss
    terms, on 96 real feature windows against the oracle encoder (tensor-core mode: 1e-3 on every term)."""
    gcase = golden_case("m5_t32")
    stats = gcase.stats()
    vids = tb.make_videos(96, 32, seed=99)
    x = torch.stack([O.window_features(vids.video(v), 0, 32, stats, gcase.mods)[0] for v in range(96)], 0)
    y = torch.arange(96) % 6
    torch.manual_seed(5)
    with torch.no_grad():
        enc = lambda t: O.encoder_forward(gcase.sd, t, gcase.dims_raw, gcase.dims_diff)[0]
        emb = enc(x)
        sh, rv, st = enc(O.partial_shuffle_within_window(x)), enc(O.reverse_sequence(x)), enc(O.get_static_window(x))
    want = {"tcl": float(O.tcl_loss(emb, y)), "hard_shuf": 10 * float(O.supcon_hard_rows(emb, emb, sh).mean()),
            "hard_rev": 10 * float(O.supcon_hard_rows(emb, emb, rv).mean()),
            "hard_stat": 10 * float(O.supcon_hard_rows(emb, emb, st).mean())}
    for precision, tol in (("fp16_tc", 1e-3), ("fp32", 1e-4)):
        model = _model(gcase, precision, max_windows=96)
        torch.manual_seed(5)
        got = tb.hard_negative_step(model, x.to(DEV), y.to(DEV))
        for k, v in want.items():
            rel = abs(float(got[k]) - v) / abs(v)
            print(f"hard_negative_step[{precision}] {k}: {float(got[k]):.6f} vs {v:.6f} (rel {rel:.2e})")
            assert rel < tol, (precision, k)
