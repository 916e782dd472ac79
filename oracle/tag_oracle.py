"""CPU ORACLE for the TAG scoring hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional restatement (torch CPU tensor ops, fp32 by default, fp64 on request) of the
reference algorithm: window feature construction -> encoder forward -> centroids -> AC / TC.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; the product package never does (it fails loudly when the CUDA
library is missing).

Parity pin: the reference ships no tests or golden vectors, so this oracle is
pinned against the UNMODIFIED reference executed in the build container; the resulting
vectors are committed under tests/golden/ together with the script that made them
(tests/golden/make_golden.py) and re-checked by `pytest -m "not gpu"`.

Every function cites the reference file:line it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# A1  window slicing                                              utils.py:366-381
# --------------------------------------------------------------------------------------
def slice_or_pad_index(L: int, start: int, T: int) -> torch.Tensor:
    """Frame indices that `WindowDataset._slice_or_pad` selects: arr[start:start+T]; short tail
    repeats the last frame; start outside [0, L) repeats the first / last frame."""
    if start < 0 or start >= L:
        idx = 0 if start < 0 else L - 1
        return torch.full((T,), idx, dtype=torch.long)
    t = torch.arange(T, dtype=torch.long) + start
    return t.clamp_max(L - 1)


# --------------------------------------------------------------------------------------
# A3..A6  frame-to-frame deltas                                   utils.py:130-217
# --------------------------------------------------------------------------------------
def log_so3(R: torch.Tensor) -> torch.Tensor:
    """utils.py:130-140. R [...,3,3] -> axis-angle [...,3]."""
    tr = (R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]).clamp(-1 + 1e-6, 3 - 1e-6)
    theta = torch.acos((tr - 1) / 2)
    denom = (2 * torch.sin(theta)).unsqueeze(-1).clamp_min(1e-6)
    v = torch.stack([R[..., 2, 1] - R[..., 1, 2],
                     R[..., 0, 2] - R[..., 2, 0],
                     R[..., 1, 0] - R[..., 0, 1]], dim=-1) / denom
    return theta.unsqueeze(-1) * v


def vit_delta(x: torch.Tensor) -> torch.Tensor:
    """utils.py:142-147. L2-normalise each frame (eps 1e-12), first difference, row 0 = 0."""
    v = F.normalize(x, dim=-1)
    return v - torch.cat([v[:1], v[:-1]], dim=0)


def betas_delta(b: torch.Tensor) -> torch.Tensor:
    """utils.py:161-163."""
    return b - torch.cat([b[:1], b[:-1]], dim=0)


def rotmat_delta(R: torch.Tensor) -> torch.Tensor:
    """utils.py:165-174. R [T,J,3,3] -> [T,J,3]: log(R_{t-1}^T R_t), t=0 pairs with itself."""
    Rp = torch.cat([R[:1], R[:-1]], dim=0)
    return log_so3(torch.matmul(Rp.transpose(-1, -2), R))


def procrustes_kp_delta(kp: torch.Tensor, eps: float = 1e-6) -> Tuple[torch.Tensor, int]:
    """utils.py:177-217 including its `R = Vh @ U.T` (sic, :209) and the det<0 fix-up (:210-212).
    Returns (delta [T,2K], number of frames with det(H) < 0 — the mirror regime, whose result is pinned by
    LAPACK's always-improper U; see procrustes_kp_delta_closed_form)."""
    T, D = kp.shape
    K = D // 2
    pts = kp.view(T, K, 2)
    pts_c = pts - pts.mean(dim=1, keepdim=True)
    s = torch.linalg.norm(pts_c, dim=(1, 2), keepdim=True).clamp_min(eps)
    pts_n = pts_c / s
    deltas = torch.zeros_like(pts_n)
    n_reflect = 0
    for t in range(1, T):
        X, Y = pts_n[t - 1], pts_n[t]
        H = X.t().matmul(Y)
        if float(H[0, 0] * H[1, 1] - H[0, 1] * H[1, 0]) < 0:
            n_reflect += 1
        U, _, Vh = torch.linalg.svd(H)
        R = Vh @ U.t()
        if torch.det(R) < 0:
            Vh = Vh.clone()
            Vh[:, -1] *= -1
            R = Vh @ U.t()
        deltas[t] = Y - X @ R
    return deltas.reshape(T, K * 2), n_reflect


def procrustes_kp_delta_closed_form(kp: torch.Tensor, eps: float = 1e-6) -> Tuple[torch.Tensor, torch.Tensor]:
    """Closed form of the above (SURVEY.md §8a A6 parity note), what the CUDA kernel evaluates; kept here so
    the CPU suite can check the algebra against the SVD form. LAPACK's 2x2 SVD (MKL, this image) always returns
    an IMPROPER U (a reflection, det U = -1: measured on 60 k random H in both regimes), which pins `Vh @ U.T`
    plus the det fix-up (utils.py:209-212) to R = [[c, s], [-s, c]] with
        angle = atan2(H10 - H01, H00 + H11)   for det(H) >= 0  (transpose of the polar rotation of H)
        angle = atan2(H10 + H01, H00 - H11)   for det(H) <  0  (angle of the polar reflection of H).
    Returns (delta, detH[T])."""
    T, D = kp.shape
    K = D // 2
    pts = kp.view(T, K, 2)
    pts_c = pts - pts.mean(dim=1, keepdim=True)
    s = torch.linalg.norm(pts_c, dim=(1, 2), keepdim=True).clamp_min(eps)
    P = pts_c / s
    X = torch.cat([P[:1], P[:-1]], dim=0)
    Y = P
    H = torch.einsum("tka,tkb->tab", X, Y)
    det = H[:, 0, 0] * H[:, 1, 1] - H[:, 0, 1] * H[:, 1, 0]
    mirror = det < 0
    ang = torch.where(mirror, torch.atan2(H[:, 1, 0] + H[:, 0, 1], H[:, 0, 0] - H[:, 1, 1]),
                      torch.atan2(H[:, 1, 0] - H[:, 0, 1], H[:, 0, 0] + H[:, 1, 1]))
    c, sn = torch.cos(ang), torch.sin(ang)
    # X @ R with R = [[c, s], [-s, c]]
    xr0 = X[..., 0] * c[:, None] - X[..., 1] * sn[:, None]
    xr1 = X[..., 0] * sn[:, None] + X[..., 1] * c[:, None]
    d = torch.stack([Y[..., 0] - xr0, Y[..., 1] - xr1], dim=-1)
    d[0] = 0.0
    det[0] = 1.0
    return d.reshape(T, D), det


# --------------------------------------------------------------------------------------
# N2  ModalityStats                                               utils.py:568-801
# --------------------------------------------------------------------------------------
_STAT_NAMES = {"vit": "vit", "global": "gori", "pose": "pose", "beta": "beta", "kp2d": "keypoints",
               "clip": "clip", "dino": "dino"}


def video_raw_diff(video: Dict[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor], int]:
    """raw and diff blocks of one (whole video | window), keyed by modality name.
    utils.py:396-404 (raw), :455-470 (diff). Returns (raw, diff, n_reflect)."""
    T = video["pose"].shape[0]
    pose = video["pose"].float()
    gori = video["global_orient"].float()
    raw = {"vit": video["vit"].float(), "global": gori.reshape(T, -1), "pose": pose.reshape(T, -1),
           "beta": video["betas"].float()}
    diff = {"vit": vit_delta(raw["vit"]), "global": rotmat_delta(gori).reshape(T, -1),
            "pose": rotmat_delta(pose).reshape(T, -1), "beta": betas_delta(raw["beta"])}
    n_reflect = 0
    if "keypoints" in video and video["keypoints"] is not None:
        raw["kp2d"] = video["keypoints"].float()
        diff["kp2d"], n_reflect = procrustes_kp_delta(raw["kp2d"])
    for m in ("clip", "dino"):
        if m in video and video[m] is not None:
            raw[m] = video[m].float()
            diff[m] = vit_delta(raw[m])
    return raw, diff, n_reflect


def compute_stats(videos: Sequence[Dict[str, torch.Tensor]], eps: float = 1e-6) -> Dict[str, torch.Tensor]:
    """`compute_stats_from_npz` (utils.py:595-801): float64 running sum / sum-of-squares over
    WHOLE videos (diffs taken across the full sequence, :717-732), mean = s/n,
    std = sqrt(max(ss/n - mean^2, 0) + eps) (:746-750). Keys: '<mod>_raw_mean' etc. with the
    reference's field prefixes (vit, gori, pose, beta, keypoints, clip, dino)."""
    acc: Dict[str, List] = {}
    for vid in videos:
        raw, diff, _ = video_raw_diff(vid)
        for kind, blocks in (("raw", raw), ("diff", diff)):
            for m, X in blocks.items():
                Xn = X.numpy()
                key = f"{_STAT_NAMES[m]}_{kind}"
                if key not in acc:
                    acc[key] = [np.zeros(Xn.shape[1], np.float64), np.zeros(Xn.shape[1], np.float64), 0]
                acc[key][0] += Xn.sum(axis=0, dtype=np.float64)
                acc[key][1] += (Xn.astype(np.float64) ** 2).sum(axis=0)
                acc[key][2] += Xn.shape[0]
    out = {}
    for key, (s1, s2, n) in acc.items():
        mean = s1 / max(1, n)
        var = s2 / max(1, n) - mean ** 2
        std = np.sqrt(np.maximum(var, 0.0) + eps)
        out[f"{key}_mean"] = torch.from_numpy(mean.astype(np.float32))
        out[f"{key}_std"] = torch.from_numpy(std.astype(np.float32))
    return out


# --------------------------------------------------------------------------------------
# A2, A7  window features                                         utils.py:383-516
# --------------------------------------------------------------------------------------
def window_features(video: Dict[str, torch.Tensor], start: int, T: int,
                    stats: Optional[Dict[str, torch.Tensor]], modalities: Sequence[str]) -> Tuple[torch.Tensor, int]:
    """`WindowDataset._try_one`: slice/pad every array, recompute diffs PER WINDOW (first window
    frame has zero diff), z-score `(x-mean)/(std+1e-6)` (:472-494), concat [raw blocks || diff
    blocks] in modality order (:496-514). Returns (feats [T,D] fp32, n_reflect)."""
    L = video["pose"].shape[0]
    idx = slice_or_pad_index(L, start, T)
    win = {k: v[idx] for k, v in video.items() if v is not None}
    raw, diff, n_reflect = video_raw_diff(win)
    if stats is not None:
        e = 1e-6
        for m in modalities:
            p = _STAT_NAMES[m]
            raw[m] = (raw[m] - stats[f"{p}_raw_mean"]) / (stats[f"{p}_raw_std"] + e)
            diff[m] = (diff[m] - stats[f"{p}_diff_mean"]) / (stats[f"{p}_diff_std"] + e)
    feats = torch.cat([raw[m] for m in modalities] + [diff[m] for m in modalities], dim=-1)
    return feats, n_reflect


# --------------------------------------------------------------------------------------
# A8..A14  encoder forward                                        model.py:1-193
# --------------------------------------------------------------------------------------
def _conv_encoder(sd, prefix: str, x_btf: torch.Tensor, dilations=(1, 2, 4, 8), taps=None) -> torch.Tensor:
    """`MovementConvEncoder.forward` (model.py:52-58) with `TemporalConvBlock` (:34-40):
    stem 1x1 conv (no bias) -> 4 x [conv(k=5,dil d,zero pad 2d) GELU conv +res GELU GroupNorm(1,C)]
    -> Linear proj (no bias). Dropout is identity in eval."""
    x = x_btf.transpose(1, 2)
    y = F.conv1d(x, sd[f"{prefix}.stem.weight"])
    if taps is not None:
        taps[f"{prefix}.stem"] = y.transpose(1, 2)
    for b, d in enumerate(dilations):
        p = f"{prefix}.blocks.{b}"
        res = y
        h = F.gelu(F.conv1d(y, sd[f"{p}.conv1.weight"], padding=2 * d, dilation=d))
        h = F.conv1d(h, sd[f"{p}.conv2.weight"], padding=2 * d, dilation=d)
        h = F.gelu(h + res)
        y = F.group_norm(h, 1, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], eps=1e-5)
        if taps is not None:
            taps[p] = y.transpose(1, 2)
    y = y.transpose(1, 2)
    return F.linear(y, sd[f"{prefix}.proj.weight"])


def _fusion(sd, M_tokens: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """`MinimalPerFrameFusion.forward` (model.py:79-98); `mask` is never used there."""
    B, T, M, D = M_tokens.shape
    kv = F.layer_norm(M_tokens, (D,), sd["fusion.kv_ln.weight"], sd["fusion.kv_ln.bias"], 1e-5).reshape(B * T, M, D)
    q = F.layer_norm(sd["fusion.latent"].expand(B * T, 1, D), (D,), sd["fusion.q_ln.weight"],
                     sd["fusion.q_ln.bias"], 1e-5)
    Q = F.linear(q, sd["fusion.Wq.weight"])
    K = F.linear(kv, sd["fusion.Wk.weight"])
    V = F.linear(kv, sd["fusion.Wv.weight"])
    logits = torch.matmul(Q, K.transpose(-2, -1)) / math.sqrt(D)
    tau = F.softplus(sd["fusion.logit_temp"]) + 1e-3
    logits = logits / tau.view(1, 1, M) + sd["fusion.logit_bias"].view(1, 1, M)
    A = logits.softmax(dim=-1)
    fused = F.linear(torch.matmul(A, V).squeeze(1), sd["fusion.Wo.weight"])
    return fused.view(B, T, D), A.squeeze(1)


def _transformer_layer(sd, p: str, x: torch.Tensor, n_heads: int) -> torch.Tensor:
    """`nn.TransformerEncoderLayer(256, 8, 1024, batch_first=True)` in eval (model.py:145):
    post-norm, ReLU FFN, biases, LN eps 1e-5, no mask; head_dim 32, scale 1/sqrt(32)."""
    B, S, D = x.shape
    hd = D // n_heads
    qkv = F.linear(x, sd[f"{p}.self_attn.in_proj_weight"], sd[f"{p}.self_attn.in_proj_bias"])
    q, k, v = qkv.split(D, dim=-1)
    sh = lambda t: t.reshape(B, S, n_heads, hd).transpose(1, 2)
    q, k, v = sh(q), sh(k), sh(v)
    att = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = torch.matmul(att, v).transpose(1, 2).reshape(B, S, D)
    o = F.linear(o, sd[f"{p}.self_attn.out_proj.weight"], sd[f"{p}.self_attn.out_proj.bias"])
    x = F.layer_norm(x + o, (D,), sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], 1e-5)
    f = F.linear(F.relu(F.linear(x, sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"])),
                 sd[f"{p}.linear2.weight"], sd[f"{p}.linear2.bias"])
    return F.layer_norm(x + f, (D,), sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], 1e-5)


def encoder_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, dims_map_raw: Dict[str, int],
                    dims_map_diff: Dict[str, int], n_heads: int = 8, taps: Optional[dict] = None):
    """`HumanActionScorer.forward` (model.py:162-193) ->
    (seq_embed [B,256], frame_embeds [B,T+1,256], tokens [B,T+1,256])."""
    mods = list(dims_map_raw.keys())
    raw_total = sum(dims_map_raw.values())
    diff_total = sum(dims_map_diff.values())
    has_diff = any(v > 0 for v in dims_map_diff.values())
    raw = x[:, :, :raw_total]
    rawp = dict(zip(mods, torch.split(raw, [dims_map_raw[m] for m in mods], dim=-1)))
    if has_diff:
        diff = x[:, :, raw_total:raw_total + diff_total]
        diffp = dict(zip(mods, torch.split(diff, [dims_map_diff[m] for m in mods], dim=-1)))
    per_mod = []
    for m in mods:
        s = _conv_encoder(sd, f"state_enc.{m}", rawp[m], taps=taps)
        if has_diff and dims_map_diff[m] > 0:
            s = s + _conv_encoder(sd, f"motion_enc.{m}", diffp[m], taps=taps)
        s = F.layer_norm(s, (s.size(-1),))                               # model.py:175, no affine
        per_mod.append(s.unsqueeze(2))
    M_tokens = torch.cat(per_mod, dim=2)
    frame_tok, attn = _fusion(sd, M_tokens)
    if taps is not None:
        taps["M_tokens"] = M_tokens
        taps["fusion.attn"] = attn
        taps["frame_tok"] = frame_tok
    B, T, D = frame_tok.shape
    tokens = torch.cat([sd["cls"].expand(B, 1, D), frame_tok], dim=1)
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("temporal.layers."))
    if "pos_enc.pe" in sd:
        pe = sd["pos_enc.pe"]
    else:
        pe = _sinusoidal_pe(T + 1, D).to(x.dtype)
    tokens = tokens + pe[:, :T + 1, :]
    for l in range(n_layers):
        tokens = _transformer_layer(sd, f"temporal.layers.{l}", tokens, n_heads)
        if taps is not None:
            taps[f"temporal.layers.{l}"] = tokens
    seq_embed = F.normalize(tokens[:, 0, :])
    frame_embeds = F.normalize(tokens, dim=-1)
    return seq_embed, frame_embeds, tokens


def _sinusoidal_pe(n: int, d: int) -> torch.Tensor:
    """model.py:8-16."""
    pe = torch.zeros(n, d)
    pos = torch.arange(0, n, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * (-math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0)


# --------------------------------------------------------------------------------------
# A15  centroids                                                  utils.py:1018-1045
# --------------------------------------------------------------------------------------
def centroid_sums(z: torch.Tensor, y: torch.Tensor, C: int) -> Tuple[torch.Tensor, torch.Tensor]:
    sums = torch.zeros(C, z.shape[1], dtype=z.dtype)
    counts = torch.zeros(C, dtype=z.dtype)
    sums.index_add_(0, y, z)
    counts.index_add_(0, y, torch.ones_like(y, dtype=z.dtype))
    return sums, counts


def centroid_finalize(sums: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    """utils.py:1042-1043."""
    return F.normalize(sums / counts.clamp_min(1.0).unsqueeze(1), dim=-1)


def build_centroids(z: torch.Tensor, y: torch.Tensor, C: int) -> Tuple[torch.Tensor, torch.Tensor]:
    sums, counts = centroid_sums(z, y, C)
    return centroid_finalize(sums, counts), counts


# --------------------------------------------------------------------------------------
# A16, A17  scores                                                eval.py:209-257
# --------------------------------------------------------------------------------------
def temporal_coherence_scores(features: dict) -> Dict[str, float]:
    """eval.py:209-226."""
    per_video: Dict[str, List[float]] = {}
    for i, name in enumerate(features["vid_names"]):
        vid = os.path.splitext(name)[0]
        fr = features["frame_embeds"][i][1:]
        if fr.shape[0] < 2:
            continue
        d = (fr[1:] - fr[:-1]).pow(2).sum(dim=-1).sqrt()
        per_video.setdefault(vid, []).append(float(d.mean().item()))
    return {v: float(np.mean(s)) for v, s in per_video.items()}


def action_consistency_scores(features: dict, centroids: torch.Tensor, label_dict: Dict[str, int]) -> Dict[str, float]:
    """eval.py:229-257 (class canonicalisation :36-45 is identity for canonical names)."""
    emb: Dict[str, List[torch.Tensor]] = {}
    cls_of: Dict[str, str] = {}
    for i, name in enumerate(features["vid_names"]):
        vid = os.path.splitext(name)[0]
        emb.setdefault(vid, []).append(features["seq_embeds"][i])
        cls_of[vid] = features["cls_names"][i]
    out = {}
    for vid, e in emb.items():
        c = cls_of[vid]
        if c not in label_dict:
            continue
        idx = label_dict[c]
        if idx >= len(centroids):
            continue
        zm = F.normalize(torch.stack(e, 0).mean(dim=0), p=2, dim=-1)
        out[vid] = float(torch.norm(zm - centroids[idx], p=2).item())
    return out


# --------------------------------------------------------------------------------------
# N1  TCL forward                                                 losses.py:14-34
# --------------------------------------------------------------------------------------
def tcl_loss_rows(z: torch.Tensor, targets: torch.Tensor, temperature=0.1, k1=5000.0, k2=1.0) -> torch.Tensor:
    """losses.py:14-31, per anchor row (before the final mean of :33)."""
    S = z @ z.T
    E = torch.exp(S / temperature)
    En = torch.exp(-S)
    same = targets[:, None] == targets[None, :]
    pos = same.to(z.dtype) * (1 - torch.eye(z.shape[0], dtype=z.dtype))
    neg = (~same).to(z.dtype)
    den = (E * pos).sum(1) + k1 * (En * pos).sum(1) + k2 * (E * neg).sum(1)
    return (-torch.log(E / den[:, None]) * pos).sum(1) / pos.sum(1)


def tcl_loss(z: torch.Tensor, targets: torch.Tensor, temperature=0.1, k1=5000.0, k2=1.0) -> torch.Tensor:
    """losses.py:14-34."""
    return tcl_loss_rows(z, targets, temperature, k1, k2).mean()


def supcon_hard_rows(anchor: torch.Tensor, positive: torch.Tensor, hard_negative: torch.Tensor, temperature=0.07) -> torch.Tensor:
    """losses.py:43-56 per sample: cross entropy over [a.p/t, a.h/t] with the positive as the target."""
    sim_ap = (anchor * positive).sum(-1) / temperature
    sim_ah = (anchor * hard_negative).sum(-1) / temperature
    logits = torch.stack([sim_ap, sim_ah], dim=1)
    return F.cross_entropy(logits, torch.zeros(anchor.shape[0], dtype=torch.long), reduction="none")


# hard-negative augmentations                                     utils.py:65-95
def partial_shuffle_within_window(seqs: torch.Tensor, shuffle_fraction: float = 0.7) -> torch.Tensor:
    """utils.py:65-75 (same RNG consumption: randperm(T)[:n], then randperm(n), per sample)."""
    out = seqs.clone()
    B, T, _ = seqs.shape
    for i in range(B):
        if T > 1:
            n = max(1, int(shuffle_fraction * T))
            idx = torch.randperm(T)[:n]
            out[i, idx] = out[i, idx][torch.randperm(n)]
    return out


def reverse_sequence(seqs: torch.Tensor) -> torch.Tensor:
    """utils.py:78-86."""
    return torch.flip(seqs, dims=[1])


def get_static_window(seqs: torch.Tensor) -> torch.Tensor:
    """utils.py:88-95: every frame replaced by the window's first frame."""
    return seqs[:, :1].expand_as(seqs).clone()


# --------------------------------------------------------------------------------------
# whole path on in-memory tensors (the "reference compute-only" comparator)
# --------------------------------------------------------------------------------------
def score_videos(videos: Sequence[Dict[str, torch.Tensor]], names: Sequence[str], cls_names: Sequence[str],
                 sd, dims_map_raw, dims_map_diff, stats, centroids, label_dict,
                 clip_len: int = 32, stride: int = 8, batch: int = 32):
    """eval.py:394-437 on in-memory videos: windows -> feats -> encoder -> AC + TC."""
    mods = list(dims_map_raw.keys())
    feats, wn, wc = [], [], []
    for v, vid in enumerate(videos):
        L = vid["pose"].shape[0]
        starts = [0] if L < clip_len else list(range(0, L - clip_len + 1, max(1, stride)))
        for s in starts:
            f, _ = window_features(vid, s, clip_len, stats, mods)
            feats.append(f); wn.append(names[v]); wc.append(cls_names[v])
    seq, frm = [], []
    with torch.no_grad():
        for i in range(0, len(feats), batch):
            x = torch.stack(feats[i:i + batch], 0)
            s, f, _ = encoder_forward(sd, x, dims_map_raw, dims_map_diff)
            seq.append(s); frm.append(f)
    features = {"seq_embeds": torch.cat(seq, 0), "frame_embeds": torch.cat(frm, 0),
                "cls_names": wc, "vid_names": wn}
    ac = action_consistency_scores(features, centroids, label_dict)
    tc = temporal_coherence_scores(features)
    return ac, tc, features
