"""CPU suite for the host-side logic and the C-ABI boundary (no GPU, no compute calls)."""
import ctypes
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

import tag_b200 as tb
from tag_b200 import _lib
from helpers import golden_case, oracle, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "tag_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(tag_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tag_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert tb.load_library().tag_abi_version() == 2


def test_fused_conv_block_shared_memory_plan():
    """Host-side plan of the fused TemporalConvBlock kernel (csrc/tcn_block_tc.cu): six 128-row tiles between shared zero halos of
    2 * dil * (128 / T) rows, then the weight ring — at least four 16 KB stages inside the 227 KB a CTA may use, else the shape stays on
    the two-kernel path."""
    lib = tb.load_library()
    ws, sm, tl = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()

    def plan(M, T, dil):
        ok = lib.tag_debug_tcn_block_plan(M, T, dil, ctypes.byref(ws), ctypes.byref(sm), ctypes.byref(tl))
        return ok, ws.value, sm.value, tl.value

    for dil, stages in ((1, 7), (2, 6), (4, 5), (8, 4)):                      # config 2: T = 32, four windows per tile
        ok, w, smem, tiles = plan(400000, 32, dil)
        halo = 2 * dil * 4
        assert ok == 1 and w == stages
        assert tiles == (6 * (halo + 128) + halo) * 128                      # multiples of 1024 here: every tile on a swizzle-atom boundary
        assert smem <= 232448 and smem >= tiles + w * 16384
    assert plan(400000, 32, 16)[0] == 0                                      # halo tiles leave no room for the weight ring
    assert plan(400000, 16, 8)[0] == 0
    assert plan(131072, 256, 1)[0] == 0                                      # a window spans two tiles
    assert plan(2400, 24, 1)[0] == 0                                         # not a power of two
    assert plan(128, 32, 1)[0] == 0 and plan(160, 32, 1)[0] == 1             # a single tile has no CTA pair
    assert plan(1000, 32, 1)[0] == 0                                         # rows must be whole windows
    ok, w, smem, tiles = plan(12800, 128, 1)                                 # halo of 2 rows: unshared layout, TMA zero fill writes the halos
    assert ok == 1 and w >= 4 and tiles % 1024 == 0 and smem <= 232448


def test_config_struct_layout_matches_header():
    # 3 arrays of 8 + 1 + 10 scalars, all int32
    assert ctypes.sizeof(_lib.tag_config) == 4 * (1 + 3 * 8 + 10)
    assert ctypes.sizeof(_lib.tag_videos) == 8 * 8 + 8 + 8


def test_create_rejects_bad_configs_without_a_gpu():
    lib = tb.load_library()
    h = ctypes.c_void_p()
    cfg = _lib.tag_config()
    cfg.n_modalities = 0
    assert lib.tag_create(ctypes.byref(h), ctypes.byref(cfg)) == 1
    assert b"n_modalities" in lib.tag_last_error(None)
    cfg.n_modalities = 1
    cfg.d_model = 128
    assert lib.tag_create(ctypes.byref(h), ctypes.byref(cfg)) == 5
    assert lib.tag_create(None, None) == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_cuda():
    r, d = tb.dims_maps()
    m = tb.HumanActionScorer(r, d).eval()
    with pytest.raises(tb.TagError):
        m(torch.zeros(1, 32, m.feat_dim))
    with pytest.raises(tb.TagError):
        tb.compute_temporal_coherence_scores({"frame_embeds": torch.zeros(2, 5, 256), "vid_names": ["a.npz", "a.npz"]})
    with pytest.raises(tb.TagError):
        tb.FeatureFuser(r, d, "cpu")


def test_product_package_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "video-gen-evals_b200")
    for fn in os.listdir(pkg_dir):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg_dir, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle|import_module\([\"']oracle|sys\.path.*reference", src, re.M), fn


def test_state_dict_contract_and_ctor_errors():
    r, d = tb.dims_maps(True)
    m = tb.HumanActionScorer(r, d)
    sd = tb.make_state_dict(r, d, seed=1)
    assert set(m.state_dict().keys()) == set(sd.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    assert m.load_state_dict(sd, strict=True)
    assert m.M == 7 and m.feat_dim == 5156 and m.one_pass_raw == 2650
    with pytest.raises(ValueError):
        tb.HumanActionScorer([1], {"a": 1})
    with pytest.raises(ValueError):
        tb.HumanActionScorer({"vit": 4}, {"pose": 4})
    # positional table == reference formula (model.py:11-15)
    pe = tb.sinusoidal_pe(64, 256)[0]
    assert abs(float(pe[3, 0]) - np.sin(3.0)) < 1e-6 and abs(float(pe[3, 1]) - np.cos(3.0)) < 1e-6


def test_window_table_matches_reference_enumeration():
    lens = [64, 64, 40, 32, 20, 33, 56, 7, 48, 100, 31, 1]
    for clip, stride in ((32, 8), (32, 1), (16, 5), (256, 8)):
        wv, ws, seg = tb.window_table(lens, clip, stride)
        ev, es = tb.enumerate_windows(lens, clip, stride)
        assert wv.tolist() == ev and ws.tolist() == es
        assert seg[-1] == len(ev) and all(seg[i + 1] - seg[i] >= 1 for i in range(len(lens)))
    g = golden_case("m5_t32")       # == what the reference's sample_all_windows_npz produced
    wins = sorted(g.gen_windows())
    ev, es = tb.enumerate_windows(g.meta["gen_lens"], 32, 8)
    assert sorted(zip(ev, es)) == wins


def test_block_plan_covers_the_batch_in_whole_passes():
    """score_stream's block plan: contiguous, complete, every block at most one encoder pass, a short ramp only at the
    start of a stream; the bench workload (5000 clips x 64 frames, 13,024-window passes, 148 SMs) lands on 43 GEMM waves."""
    wave = 74 * 256 // 32
    for lengths, first in (([64] * 5000, True), ([64] * 5000, False), ([64, 7, 200, 32, 33, 500] * 40, True), ([20], True), ([], False)):
        plan = tb.block_plan(lengths, 32, 8, 13024, first, 148)
        seg = tb.window_table(lengths, 32, 8)[2]
        assert [a for a, _ in plan] == [0] * bool(plan) + [b for _, b in plan[:-1]]          # contiguous
        assert (plan[-1][1] if plan else 0) == len(lengths)                                 # complete
        for a, b in plan:
            assert b > a and (seg[b] - seg[a] <= 13024 or b == a + 1)
    ramp = tb.block_plan([64] * 5000, 32, 8, 13024, True, 148)
    seg = tb.window_table([64] * 5000, 32, 8)[2]
    assert [int(-(-(seg[b] - seg[a]) // wave)) for a, b in ramp] == [1, 3, 15, 22, 2]
    steady = tb.block_plan([64] * 5000, 32, 8, 13024, False, 148)
    assert [int(-(-(seg[b] - seg[a]) // wave)) for a, b in steady] == [22, 21]


def test_shard_range_partitions():
    for n in (0, 1, 7, 5000, 100000):
        for world in (1, 2, 4, 8):
            spans = [tb.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_stats_vectors_and_infer_dims_order():
    g = golden_case("m7_t256")
    st = tb.ModalityStats.from_dict(g.stats())
    assert tb.infer_dims_from_stats(st) == (g.dims_raw, g.dims_diff)
    mean, std = tb.stats_vectors(st, g.mods, "cpu")
    assert mean.shape == (5156,) and std.shape == (5156,)
    assert torch.equal(mean[:1024], st.vit_raw_mean) and torch.equal(mean[1024:1033], st.gori_raw_mean)
    assert torch.equal(mean[2650:2650 + 1024], st.vit_diff_mean) and torch.equal(std[-768:], st.dino_diff_std)
    # same field names / order as the reference dataclass (utils.py:570-586)
    names = [f.name for f in tb.ModalityStats.__dataclass_fields__.values()]
    assert names[:4] == ["vit_raw_mean", "vit_raw_std", "gori_raw_mean", "gori_raw_std"] and len(names) == 28


def test_segments_and_json_writer(tmp_path):
    from tag_b200.scoring import _segments, _canonicalize_class
    vids, perm, offs = _segments(["b.npz", "a.npz", "b.npz", "c", "a.npz"])
    assert vids == ["b", "a", "c"] and perm == [0, 2, 1, 4, 3] and offs == [0, 2, 4, 5]
    assert _canonicalize_class("tennisswing") == "TennisSwing" and _canonicalize_class("Foo") == "Foo"
    out = tb.write_video_scores(str(tmp_path / "video_scores.json"), {"x": 0.1, "y": 0.2}, {"x": 0.3})
    assert out == {"x": {"ac": 0.1, "tc": 0.3}, "y": {"ac": 0.2}}
    import json
    assert json.load(open(tmp_path / "video_scores.json")) == out


def test_synth_is_deterministic_and_smooth():
    a = tb.make_videos(3, [10, 5, 8], seed=9, appearance=True)
    b = tb.make_videos(3, [10, 5, 8], seed=9, appearance=True)
    assert a.offsets == [0, 10, 15, 23] and torch.equal(a.vit, b.vit) and torch.equal(a.pose, b.pose)
    assert a.clip.shape == (23, 512) and a.dino.shape == (23, 768)
    R = a.pose[0, 0]
    assert float((R @ R.T - torch.eye(3)).abs().max()) < 1e-5
    sel = a.select([2, 0])
    assert sel.offsets == [0, 8, 18] and torch.equal(sel.vit[:8], a.vit[15:23])


_GLOO_WORKER = textwrap.dedent("""
    import os, sys, importlib, torch, torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    rank, world = int(sys.argv[1]), int(sys.argv[2])
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[3], RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tb = importlib.import_module("video-gen-evals_b200")
    O = importlib.import_module("oracle.tag_oracle")
    g = torch.Generator().manual_seed(0)
    N, C = 1001, 10
    z = torch.nn.functional.normalize(torch.randn(N, 256, generator=g), dim=-1)
    y = torch.randint(0, C, (N,), generator=g)
    lo, hi = tb.shard_range(N, rank, world)
    s, c = O.centroid_sums(z[lo:hi], y[lo:hi], C)            # stands in for the K3 kernel on this rank's shard
    sc = torch.cat([s, c[:, None]], 1)                       # packed [C, 257]
    tb.allreduce_centroid_sums(sc)                           # the path's one collective
    cen = O.centroid_finalize(sc[:, :256], sc[:, 256])
    ref, cnt = O.build_centroids(z, y, C)
    assert torch.equal(sc[:, 256], cnt), (sc[:, 256], cnt)
    assert float((cen - ref).abs().max()) < 1e-6
    dist.barrier(); dist.destroy_process_group()
    print("ok", rank)
""")


def test_centroid_allreduce_algebra_world2_gloo(tmp_path):
    """N>1 path on CPU: block-sharded class sums, packed [C,257] buffer, all-reduce, finalize == single-rank."""
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", port], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_npz_ingest_reads_the_reference_layouts(tmp_path):
    """N3: .npz (zlib-compressed, as extract_mesh.py:35 writes them, and stored) + keypoints.npy -> packed arrays, both on-disk
    layouts, ragged lengths, float64 members, a keypoint file one frame short; names -> classes as eval.py:55-74."""
    import importlib
    RR = importlib.import_module("oracle.ref_runner")
    vb = synth.make_videos(7, [64, 40, 33, 64, 7, 20, 64], seed=21, appearance=True)
    vb.names = ["Hunyuan_PushUps_01_aa.npz", "wan21_Soccerjuggling_02_bb.npz", "Opensora_768_HulaHoop_03_cc.npz", "x_TennisSwing_1.npz",
                "cog_WallPushups_9.npz", "Runway_Shotput_5.npz", "clip_Mystery_7.npz"]
    want_cls = ["PushUps", "SoccerJuggling", "HulaHoop", "TennisSwing", "WallPushups", "Shotput", "Mystery"]
    g = tmp_path / "gen"
    RR.write_set(vb, str(g / "meshes"), str(g / "generated_kps"), generated=True, clip_dir=str(g / "clip"), dino_dir=str(g / "dino"))
    # re-write two files the way the extraction tools do: compressed, and one with float64 members; one short keypoint file
    d = vb.video(1)
    np.savez_compressed(g / "meshes" / vb.names[1], pose=d["pose"].numpy(), betas=d["betas"].numpy(),
                        global_orient=d["global_orient"].numpy(), vit=d["vit"].numpy())
    d = vb.video(2)
    np.savez_compressed(g / "meshes" / vb.names[2], pose=d["pose"].numpy().astype(np.float64), betas=d["betas"].numpy(),
                        global_orient=d["global_orient"].numpy(), vit=d["vit"].numpy().astype(np.float64))
    np.save(g / "generated_kps" / "x_TennisSwing_1" / "keypoints.npy", vb.video(3)["keypoints"].numpy()[:-1])
    ing = tb.NpzIngest(str(g / "meshes"), str(g / "generated_kps"), generated=True, clip_dir=str(g / "clip"), dino_dir=str(g / "dino"), threads=3)
    items = ing.scan()
    order = sorted(range(7), key=lambda i: vb.names[i])
    assert [it.name for it in items] == [vb.names[i] for i in order]
    assert [it.cls for it in items] == [want_cls[i] for i in order]
    assert [it.length for it in items] == [vb.length(i) for i in order]
    got = ing.load(items, pin=False)
    ref = vb.select(order)
    for name in ("pose", "gori", "betas", "vit", "clip", "dino"):
        assert torch.equal(getattr(got, name), getattr(ref, name)), name
    kp_ref = ref.kp.clone()
    j = order.index(3)
    kp_ref[ref.offsets[j + 1] - 1] = kp_ref[ref.offsets[j + 1] - 2]        # short keypoint file: last frame repeated
    assert torch.equal(got.kp, kp_ref)
    assert got.offsets == ref.offsets and got.cls_idx[order.index(6)] == -1
    assert [b.n_videos for b in ing.batches(items, videos_per_batch=3)] == [3, 3, 1]
    # real layout (<Class>/<name>.npz), class filter
    r = tmp_path / "real"
    rv = synth.make_videos(6, 24, seed=22)
    RR.write_set(rv, str(r / "meshes"), str(r / "kp"), generated=False)
    ing2 = tb.NpzIngest(str(r / "meshes"), str(r / "kp"), generated=False, filter_classes=tb.ACTION_CLASSES[:4])
    it2 = ing2.scan()
    assert sorted(i.cls for i in it2) == sorted(rv.cls_name(v) for v in range(6) if rv.cls_name(v) in tb.ACTION_CLASSES[:4])
    b2 = ing2.load(it2, pin=False)
    k = [rv.names.index(i.name) for i in it2]
    assert torch.equal(b2.vit, rv.select(k).vit) and torch.equal(b2.kp, rv.select(k).kp)
    # a missing keypoint file is the reference's FileNotFoundError (utils.py:416-417)
    os.remove(r / "kp" / it2[0].cls / os.path.splitext(it2[0].name)[0] / "keypoints.npy")
    with pytest.raises(FileNotFoundError):
        ing2.scan()
