#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded
synthetic inputs. Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

What is pinned (SURVEY.md §8c — the reference has no tests/golden vectors of its own, so these
are outputs of the reference itself): ModalityStats, WindowDataset features, HumanActionScorer
forward (incl. intermediate activations through forward hooks), build_train_centroids_subset,
compute_action_consistency_scores, compute_temporal_coherence_scores, TCL loss.
Inputs/weights are NOT stored: they are rebuilt from seeds by video-gen-evals_b200/synth.py
(numpy PCG64, machine independent).
"""
import importlib
import json
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("TAG_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

synth = importlib.import_module("video-gen-evals_b200.synth")

import model as ref_model      # noqa: E402  (reference)
import utils as ref_utils      # noqa: E402
import eval as ref_eval        # noqa: E402
import losses as ref_losses    # noqa: E402
from torch.utils.data import DataLoader  # noqa: E402


def write_set(vb, mesh_dir, kp_dir, generated: bool, clip_dir=None, dino_dir=None):
    for v in range(vb.n_videos):
        d = vb.video(v)
        cls = vb.cls_name(v)
        stem = os.path.splitext(vb.names[v])[0]
        mdir = mesh_dir if generated else os.path.join(mesh_dir, cls)
        os.makedirs(mdir, exist_ok=True)
        np.savez(os.path.join(mdir, vb.names[v]), pose=d["pose"].numpy(), betas=d["betas"].numpy(),
                 global_orient=d["global_orient"].numpy(), vit=d["vit"].numpy())
        sub = (lambda root: os.path.join(root, stem)) if generated else (lambda root: os.path.join(root, cls, stem))
        os.makedirs(sub(kp_dir), exist_ok=True)
        np.save(os.path.join(sub(kp_dir), "keypoints.npy"), d["keypoints"].numpy())
        if clip_dir is not None:
            os.makedirs(sub(clip_dir), exist_ok=True)
            np.savez(os.path.join(sub(clip_dir), "clip_embeddings.npz"), embeddings=d["clip"].numpy())
        if dino_dir is not None:
            os.makedirs(sub(dino_dir), exist_ok=True)
            np.savez(os.path.join(sub(dino_dir), "dino_embeddings.npz"), embeddings=d["dino"].numpy())


def stats_to_np(stats):
    out = {}
    for k, v in vars(stats).items():
        if v is not None:
            out[f"stats.{k}"] = v.numpy()
    return out


def run_case(tag, *, appearance, real_n, real_len, gen_lens, clip_len, stride, seed_w, full_feat_windows,
             tap_windows, frame_windows):
    tmp = tempfile.mkdtemp(prefix=f"tag_golden_{tag}_")
    try:
        real_dir = os.path.join(tmp, "real_meshes")
        real_kp = os.path.join(tmp, "SAVE_REAL_KP")           # not a "generated" layout
        gen_dir = os.path.join(tmp, "generated_meshes")
        gen_kp = os.path.join(tmp, "generated_kps")           # triggers the flat layout (utils.py:411)
        clip_r = os.path.join(tmp, "clip_real") if appearance else None
        dino_r = os.path.join(tmp, "dino_real") if appearance else None
        clip_g = os.path.join(tmp, "clip_gen") if appearance else None
        dino_g = os.path.join(tmp, "dino_gen") if appearance else None

        real = synth.make_videos(real_n, real_len, seed=1337 + 100, appearance=appearance, name_prefix="real_")
        gen = synth.make_videos(len(gen_lens), gen_lens, seed=1337 + 200, appearance=appearance, name_prefix="gen_")
        write_set(real, real_dir, real_kp, generated=False, clip_dir=clip_r, dino_dir=dino_r)
        write_set(gen, gen_dir, gen_kp, generated=True, clip_dir=clip_g, dino_dir=dino_g)

        # ---- eval.py:367-375 ----
        real_ds = ref_utils.NpzVideoDataset(real_dir, filter_classes=ref_eval.ACTION_CLASSES)
        train_ds, _ = ref_utils.train_test_split(real_ds, train_ratio=0.8, seed=1337)
        stats = ref_utils.compute_stats_from_npz(train_ds.items, keypoint_dir=real_kp, clip_dir=clip_r, dino_dir=dino_r)
        dims_raw, dims_diff = ref_eval.infer_dims_from_stats(stats)
        exp_raw, exp_diff = synth.dims_maps(appearance)
        assert dims_raw == exp_raw and dims_diff == exp_diff, (dims_raw, dims_diff)

        # ---- model with seeded weights; strict=True checks the state-dict key contract ----
        mdl = ref_model.HumanActionScorer(dims_raw, dims_diff)
        sd = synth.make_state_dict(dims_raw, dims_diff, seed=seed_w)
        mdl.load_state_dict(sd, strict=True)
        mdl.eval()

        # ---- centroids: eval.py:260-286 ----
        label_dict = {c: i for i, c in enumerate(sorted({it.cls for it in real_ds.items}))}
        real_loader = ref_utils.make_test_loader(train_ds, clip_len=clip_len, stride=stride, stats=stats, seed=1337,
                                                 batch_size=64, keypoint_dir=real_kp, clip_dir=clip_r, dino_dir=dino_r,
                                                 num_workers=0)
        centroids, counts = ref_utils.build_train_centroids_subset(mdl, real_loader, label_dict, device="cpu")
        mdl.eval()
        real_windows = [(it.name, s) for it, s in real_loader.dataset.samples]

        # ---- generated windows: eval.py:394-418 ----
        gen_ds = ref_eval.create_dataset_from_generated_meshes(gen_dir)
        samples = ref_utils.sample_all_windows_npz(gen_ds, clip_len=clip_len, stride=stride)
        wds = ref_utils.WindowDataset(samples=samples, clip_len=clip_len, stats=stats, keypoint_dir=gen_kp,
                                      clip_dir=clip_g, dino_dir=dino_g)
        loader = DataLoader(wds, batch_size=32, shuffle=False, num_workers=0, collate_fn=ref_utils.safe_collate)

        feats_all = torch.stack([wds[i][0] for i in range(len(wds))], 0)

        # forward hooks for intermediate activations (tap_windows only)
        taps = {}
        hooks = []

        def hook(name):
            def fn(_m, _i, out):
                taps.setdefault(name, []).append(out.detach().clone())
            return fn
        tap_mods = ["vit", "pose", "kp2d"]
        for m in tap_mods:
            for side in ("state_enc", "motion_enc"):
                enc = getattr(mdl, side)[m]
                hooks.append(enc.stem.register_forward_hook(hook(f"{side}.{m}.stem")))
                for b in range(4):
                    hooks.append(enc.blocks[b].register_forward_hook(hook(f"{side}.{m}.blocks.{b}")))
                hooks.append(enc.register_forward_hook(hook(f"{side}.{m}.out")))
        hooks.append(mdl.fusion.register_forward_hook(hook("frame_tok")))
        for l in range(4):
            hooks.append(mdl.temporal.layers[l].register_forward_hook(hook(f"temporal.layers.{l}")))
        if tap_windows:
            with torch.no_grad():
                mdl(feats_all[tap_windows])
        for h in hooks:
            h.remove()
        tap_out = {}
        for k, v in taps.items():
            t = v[0]
            is_conv_layout = k.startswith(("state_enc.", "motion_enc.")) and (k.endswith(".stem") or ".blocks." in k)
            if is_conv_layout:
                t = t.transpose(1, 2)                     # conv layout [B,C,T] -> [B,T,C]
            tap_out[f"tap.{k}"] = t.numpy()
        if tap_windows:
            tap_out["tap.fusion.attn"] = mdl.last_attn.numpy()

        features = ref_eval.extract_window_features(mdl, loader, device="cpu")
        ac = ref_eval.compute_action_consistency_scores(features, centroids, label_dict)
        tc = ref_eval.compute_temporal_coherence_scores(features)

        # fp64 copy of the reference model = the reference's own fp32 noise floor
        mdl64 = ref_model.HumanActionScorer(dims_raw, dims_diff).double()
        mdl64.load_state_dict({k: v.double() for k, v in sd.items()}, strict=True)
        mdl64.eval()
        with torch.no_grad():
            seq64 = torch.cat([mdl64(feats_all[i:i + 16].double())[0] for i in range(0, len(feats_all), 16)], 0)

        # TCL forward on the seq embeddings (losses.py:14-34) with class labels
        # labels in a P x K pattern (every sample has positives, otherwise the reference returns NaN)
        y = torch.arange(features["seq_embeds"].shape[0]) % 4
        tcl = float(ref_losses.TCL()(features["seq_embeds"], y)) if len(y) >= 8 else None

        meta = {
            "tag": tag, "appearance": appearance, "clip_len": clip_len, "stride": stride, "seed_w": seed_w,
            "real_n": real_n, "real_len": real_len, "gen_lens": list(gen_lens),
            "real_seed": 1337 + 100, "gen_seed": 1337 + 200,
            "train_names": [it.name for it in train_ds.items],
            "real_windows": real_windows,
            "gen_windows": [(it.name, s) for it, s in samples],
            "vid_names": features["vid_names"], "cls_names": features["cls_names"],
            "label_dict": label_dict, "ac": ac, "tc": tc, "tcl": tcl,
            "full_feat_windows": list(full_feat_windows), "tap_windows": list(tap_windows),
            "frame_windows": list(frame_windows),
            "torch": torch.__version__, "numpy": np.__version__,
        }
        out = {
            "centroids": centroids.numpy(), "counts": counts.numpy(),
            "seq_embeds": features["seq_embeds"].numpy(),
            "seq_embeds_fp64": seq64.numpy(),
            "frame_embeds_sel": features["frame_embeds"][frame_windows].numpy(),
            "feats_sel": feats_all[full_feat_windows].numpy() if full_feat_windows else np.zeros((0,), np.float32),
            "feats_rowsum": feats_all.double().sum(dim=1).float().numpy(),      # [N, D] per-window column sums
            "feats_abssum": feats_all.double().abs().sum(dim=(1, 2)).numpy(),
        }
        out.update(stats_to_np(stats))
        out.update(tap_out)
        np.savez_compressed(os.path.join(HERE, f"{tag}.npz"), **out)
        with open(os.path.join(HERE, f"{tag}.json"), "w") as f:
            json.dump(meta, f, indent=1)
        print(f"[{tag}] windows={len(samples)} videos={len(ac)} AC[0..3]={list(ac.values())[:3]} "
              f"TC[0..3]={list(tc.values())[:3]} tcl={tcl} counts={counts.tolist()}")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def delta_fn_case():
    """Direct pins of the delta functions incl. edge cases (utils.py:130-217)."""
    g = np.random.default_rng(7)
    vb = synth.make_videos(3, 24, seed=77)
    out = {}
    for v in range(3):
        d = vb.video(v)
        out[f"v{v}.vit_delta"] = ref_utils._vit_delta(d["vit"]).numpy()
        out[f"v{v}.pose_delta"] = ref_utils._rotmat_delta(d["pose"]).numpy()
        out[f"v{v}.gori_delta"] = ref_utils._rotmat_delta(d["global_orient"]).numpy()
        out[f"v{v}.betas_delta"] = ref_utils._betas_delta(d["betas"]).numpy()
        out[f"v{v}.kp_delta"] = ref_utils._procrustes_kp_delta(d["keypoints"]).numpy()
    # edge: identical consecutive frames, large rotations (theta near pi), zero vit rows, invisible (-1) keypoints
    aa = torch.from_numpy(g.standard_normal((16, 4, 3)).astype(np.float32)) * 2.5
    R = synth._rodrigues(aa)
    R[5] = R[4]
    out["edge.R"] = R.numpy()
    out["edge.R_delta"] = ref_utils._rotmat_delta(R).numpy()
    x = torch.from_numpy(g.standard_normal((8, 64)).astype(np.float32))
    x[3] = 0.0
    out["edge.x"] = x.numpy()
    out["edge.x_delta"] = ref_utils._vit_delta(x).numpy()
    kp = vb.video(0)["keypoints"][:8].clone()
    kp[:, 10:20] = -1.0
    kp[4] = kp[3]
    out["edge.kp"] = kp.numpy()
    out["edge.kp_delta"] = ref_utils._procrustes_kp_delta(kp).numpy()
    np.savez_compressed(os.path.join(HERE, "deltas.npz"), **out)
    print("[deltas] done")


def mirror_case():
    """`_procrustes_kp_delta` (utils.py:177-217) in its det(H) < 0 regime (mirror-like consecutive frames): i.i.d. random
    frames (about half of the pairs) and a smooth sequence whose odd frames are left/right flipped (every pair).
    Separate file (deltas_mirror.npz) so that deltas.npz stays bit-identical to its first generation."""
    g = np.random.default_rng(11)
    out = {}
    kp = torch.from_numpy(g.random((48, 120)).astype(np.float32))
    out["iid.kp"] = kp.numpy()
    out["iid.kp_delta"] = ref_utils._procrustes_kp_delta(kp).numpy()
    kp = synth.make_videos(1, 40, seed=78).video(0)["keypoints"].clone()
    flip = kp.view(40, 60, 2).clone()
    flip[1::2, :, 0] = 1.0 - flip[1::2, :, 0]
    kp = flip.reshape(40, 120)
    out["flip.kp"] = kp.numpy()
    out["flip.kp_delta"] = ref_utils._procrustes_kp_delta(kp).numpy()
    np.savez_compressed(os.path.join(HERE, "deltas_mirror.npz"), **out)
    print("[deltas_mirror] done")


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    if len(sys.argv) > 1 and sys.argv[1] == "mirror":
        mirror_case()
        sys.exit(0)
    delta_fn_case()
    mirror_case()
    # M=5, reference defaults clip_len 32 / stride 8 (eval.py:358-359); lengths cover: exact multiple,
    # ragged tail, == clip_len, short (padded) and clip_len+1
    run_case("m5_t32", appearance=False, real_n=20, real_len=48,
             gen_lens=[64, 64, 40, 32, 20, 33, 56, 64, 7, 48, 35, 64], clip_len=32, stride=8, seed_w=0,
             full_feat_windows=[0, 11, 13], tap_windows=[13], frame_windows=[0, 5, 13, 20])
    # M=7 (vit+clip+dino appearance), long clip: one 256-frame window per video (BASELINE config 4 shape)
    run_case("m7_t256", appearance=True, real_n=10, real_len=256, gen_lens=[256, 256, 200], clip_len=256, stride=8,
             seed_w=1, full_feat_windows=[], tap_windows=[], frame_windows=[0, 2])
