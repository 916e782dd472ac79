"""Synthetic inputs and seeded encoder weights for the TAG scoring hot path.

Shapes follow what the upstream extraction tools emit (reference extract_mesh.py:25-44,
modifications/process_video.py:57): per frame `pose [23,3,3]`, `global_orient [1,3,3]`,
`betas [10]`, `vit [1024]`, `keypoints [120]` (+ optional `clip [512]`, `dino [768]`).
The generator is the recipe of SURVEY.md §8(d): smooth sequences so that the Procrustes
keypoint delta (reference utils.py:177-217) stays in its closed-form regime (det H > 0).

Everything here is deterministic in `seed` and independent of the reference, so the same
inputs/weights can be rebuilt on a box that has no /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

# reference eval.py:22-33 (already sorted; label order = sorted classes, eval.py:271)
ACTION_CLASSES = [
    "BodyWeightSquats", "HulaHoop", "JumpingJack", "PullUps", "PushUps",
    "Shotput", "SoccerJuggling", "TennisSwing", "ThrowDiscus", "WallPushups",
]

# modality order = concat order of reference utils.py:496-510 / eval.py:106-131
MODALITY_ORDER = ("vit", "global", "pose", "beta", "kp2d", "clip", "dino")
RAW_DIMS = {"vit": 1024, "global": 9, "pose": 207, "beta": 10, "kp2d": 120, "clip": 512, "dino": 768}
DIFF_DIMS = {"vit": 1024, "global": 3, "pose": 69, "beta": 10, "kp2d": 120, "clip": 512, "dino": 768}


def dims_maps(appearance: bool = False):
    """(dims_map_raw, dims_map_diff) as `infer_dims_from_stats` builds them (eval.py:104-133)."""
    mods = MODALITY_ORDER if appearance else MODALITY_ORDER[:5]
    return {m: RAW_DIMS[m] for m in mods}, {m: DIFF_DIMS[m] for m in mods}


@dataclass
class VideoBatch:
    """Frames of V videos packed back to back (frame-major), on one device.

    offsets[v] .. offsets[v+1] are the frames of video v. All float tensors are fp32.
    """
    pose: torch.Tensor            # [F, 23, 3, 3]
    gori: torch.Tensor            # [F, 1, 3, 3]
    betas: torch.Tensor           # [F, 10]
    vit: torch.Tensor             # [F, 1024]
    kp: torch.Tensor              # [F, 120]
    offsets: List[int]            # V+1
    cls_idx: List[int]            # V
    names: List[str]              # V  ("<Class>_<idx>.npz")
    clip: Optional[torch.Tensor] = None   # [F, 512]
    dino: Optional[torch.Tensor] = None   # [F, 768]
    classes: List[str] = field(default_factory=lambda: list(ACTION_CLASSES))

    @property
    def n_videos(self) -> int:
        return len(self.offsets) - 1

    @property
    def n_frames(self) -> int:
        return int(self.offsets[-1])

    def length(self, v: int) -> int:
        return int(self.offsets[v + 1] - self.offsets[v])

    def cls_name(self, v: int) -> str:
        return self.classes[self.cls_idx[v]]

    def to(self, device) -> "VideoBatch":
        mv = lambda t: None if t is None else t.to(device, non_blocking=True)
        return VideoBatch(mv(self.pose), mv(self.gori), mv(self.betas), mv(self.vit), mv(self.kp),
                          list(self.offsets), list(self.cls_idx), list(self.names),
                          mv(self.clip), mv(self.dino), list(self.classes))

    def pin(self) -> "VideoBatch":
        pm = lambda t: None if t is None else t.pin_memory()
        return VideoBatch(pm(self.pose), pm(self.gori), pm(self.betas), pm(self.vit), pm(self.kp),
                          list(self.offsets), list(self.cls_idx), list(self.names),
                          pm(self.clip), pm(self.dino), list(self.classes))

    def video(self, v: int) -> Dict[str, torch.Tensor]:
        a, b = self.offsets[v], self.offsets[v + 1]
        d = {"pose": self.pose[a:b], "global_orient": self.gori[a:b], "betas": self.betas[a:b],
             "vit": self.vit[a:b], "keypoints": self.kp[a:b]}
        if self.clip is not None:
            d["clip"] = self.clip[a:b]
        if self.dino is not None:
            d["dino"] = self.dino[a:b]
        return d

    def select(self, vids: Sequence[int]) -> "VideoBatch":
        idx = torch.cat([torch.arange(self.offsets[v], self.offsets[v + 1]) for v in vids]).to(self.pose.device)
        off = [0]
        for v in vids:
            off.append(off[-1] + self.length(v))
        g = lambda t: None if t is None else t.index_select(0, idx)
        return VideoBatch(g(self.pose), g(self.gori), g(self.betas), g(self.vit), g(self.kp), off,
                          [self.cls_idx[v] for v in vids], [self.names[v] for v in vids],
                          g(self.clip), g(self.dino), list(self.classes))

    def slice(self, v0: int, v1: int) -> "VideoBatch":
        """videos [v0, v1) as VIEWS of the packed arrays (no copy)."""
        a, b = self.offsets[v0], self.offsets[v1]
        g = lambda t: None if t is None else t[a:b]
        return VideoBatch(g(self.pose), g(self.gori), g(self.betas), g(self.vit), g(self.kp),
                          [o - a for o in self.offsets[v0:v1 + 1]], self.cls_idx[v0:v1], self.names[v0:v1],
                          g(self.clip), g(self.dino), list(self.classes))

    def input_bytes(self) -> int:
        n = 0
        for t in (self.pose, self.gori, self.betas, self.vit, self.kp, self.clip, self.dino):
            if t is not None:
                n += t.numel() * t.element_size()
        return n


def _rodrigues(aa: torch.Tensor) -> torch.Tensor:
    """axis-angle [...,3] -> rotation matrix [...,3,3] (same formula as reference utils.py:114-128)."""
    theta = aa.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    k = aa / theta
    kx, ky, kz = k[..., 0], k[..., 1], k[..., 2]
    O = torch.zeros_like(kx)
    K = torch.stack([torch.stack([O, -kz, ky], -1), torch.stack([kz, O, -kx], -1),
                     torch.stack([-ky, kx, O], -1)], -2)
    I = torch.eye(3, device=aa.device, dtype=aa.dtype).expand(aa.shape[:-1] + (3, 3))
    s = torch.sin(theta)[..., None]
    c = torch.cos(theta)[..., None]
    return I + s * K + (1.0 - c) * (K @ K)


class _NumpyRng:
    def __init__(self, seed):
        self.g = np.random.default_rng(seed)

    def normal(self, *shape):
        return torch.from_numpy(self.g.standard_normal(shape, dtype=np.float32))

    def uniform(self, *shape):
        return torch.from_numpy(self.g.random(shape, dtype=np.float32))


class _TorchRng:
    def __init__(self, seed, device):
        self.device = torch.device(device)
        self.g = torch.Generator(device=self.device)
        self.g.manual_seed(int(seed))

    def normal(self, *shape):
        return torch.randn(*shape, generator=self.g, device=self.device, dtype=torch.float32)

    def uniform(self, *shape):
        return torch.rand(*shape, generator=self.g, device=self.device, dtype=torch.float32)


def make_videos(n_videos: int, length, seed: int, appearance: bool = False,
                device="cpu", chunk: int = 512, name_prefix: str = "") -> VideoBatch:
    """SURVEY.md §8(d) generator. `length` = int or per-video list. CPU => numpy PCG64
    (bit-stable across machines, used for committed goldens); CUDA => torch generator."""
    dev = torch.device(device)
    rng = _NumpyRng(seed) if dev.type == "cpu" else _TorchRng(seed, dev)
    lens = [int(length)] * n_videos if isinstance(length, int) else [int(x) for x in length]
    assert len(lens) == n_videos
    parts = {k: [] for k in ("pose", "gori", "betas", "vit", "kp", "clip", "dino")}
    uniform_len = len(set(lens)) == 1
    v = 0
    while v < n_videos:
        nb = min(chunk, n_videos - v) if uniform_len else 1
        L = lens[v]
        aa = 0.5 * rng.normal(nb, 1, 24, 3) + torch.cumsum(0.03 * rng.normal(nb, L, 24, 3), dim=1)
        R = _rodrigues(aa)                                  # [nb, L, 24, 3, 3]
        parts["gori"].append(R[:, :, :1].reshape(nb * L, 1, 3, 3))
        parts["pose"].append(R[:, :, 1:].reshape(nb * L, 23, 3, 3))
        parts["betas"].append((rng.normal(nb, 1, 10) + 0.02 * rng.normal(nb, L, 10)).reshape(nb * L, 10))
        parts["vit"].append((rng.normal(nb, 1, 1024) +
                             torch.cumsum(0.05 * rng.normal(nb, L, 1024), dim=1)).reshape(nb * L, 1024))
        kp = (0.2 + 0.6 * rng.uniform(nb, 1, 120)) + torch.cumsum(0.005 * rng.normal(nb, L, 120), dim=1)
        parts["kp"].append(kp.clamp(0.0, 1.0).reshape(nb * L, 120))
        if appearance:
            parts["clip"].append((rng.normal(nb, 1, 512) +
                                  torch.cumsum(0.05 * rng.normal(nb, L, 512), dim=1)).reshape(nb * L, 512))
            parts["dino"].append((rng.normal(nb, 1, 768) +
                                  torch.cumsum(0.05 * rng.normal(nb, L, 768), dim=1)).reshape(nb * L, 768))
        v += nb
    cat = lambda k: torch.cat(parts[k], 0).contiguous() if parts[k] else None
    offsets = [0]
    for L in lens:
        offsets.append(offsets[-1] + L)
    cls_idx = [i % len(ACTION_CLASSES) for i in range(n_videos)]
    names = [f"{name_prefix}{ACTION_CLASSES[c]}_{i:06d}.npz" for i, c in enumerate(cls_idx)]
    return VideoBatch(cat("pose"), cat("gori"), cat("betas"), cat("vit"), cat("kp"), offsets, cls_idx, names,
                      cat("clip"), cat("dino"))


def enumerate_windows(lengths: Sequence[int], clip_len: int = 32, stride: int = 8):
    """(video, start) for every window, as reference utils.py:888-911 `sample_all_windows_npz`
    (== make_test_loader utils.py:823-837 for length > 0): slide with `stride`; a video shorter
    than clip_len yields one padded window at start 0."""
    vids, starts = [], []
    for v, L in enumerate(lengths):
        if L < clip_len:
            vids.append(v); starts.append(0)
            continue
        for s in range(0, L - clip_len + 1, max(1, stride)):
            vids.append(v); starts.append(s)
    return vids, starts


def sinusoidal_pe(max_len: int = 5000, d_model: int = 256) -> torch.Tensor:
    """reference model.py:8-16 buffer `pos_enc.pe` [1, max_len, d_model]."""
    pe = torch.zeros(max_len, d_model)
    pos = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0)


def state_dict_spec(dims_map_raw: Dict[str, int], dims_map_diff: Dict[str, int], d_model: int = 256,
                    time_layers: int = 4, ffn: Optional[int] = None, k: int = 5, n_blocks: int = 4):
    """[(key, shape, kind)] of every tensor in reference `HumanActionScorer.state_dict()`
    (model.py:102-146). kind in {w (fan-in uniform), gamma, beta, randn, temp}."""
    ffn = ffn or 4 * d_model
    spec = []
    mods = list(dims_map_raw.keys())

    def enc(prefix, d_in):
        spec.append((f"{prefix}.stem.weight", (d_model, d_in, 1), "w"))
        for b in range(n_blocks):
            spec.append((f"{prefix}.blocks.{b}.conv1.weight", (d_model, d_model, k), "w"))
            spec.append((f"{prefix}.blocks.{b}.conv2.weight", (d_model, d_model, k), "w"))
            spec.append((f"{prefix}.blocks.{b}.norm.weight", (d_model,), "gamma"))
            spec.append((f"{prefix}.blocks.{b}.norm.bias", (d_model,), "beta"))
        spec.append((f"{prefix}.proj.weight", (d_model, d_model), "w"))

    for m in mods:
        enc(f"state_enc.{m}", dims_map_raw[m])
    for m in mods:
        if dims_map_diff[m] > 0:
            enc(f"motion_enc.{m}", dims_map_diff[m])
    spec += [("fusion.latent", (1, 1, d_model), "randn"),
             ("fusion.q_ln.weight", (d_model,), "gamma"), ("fusion.q_ln.bias", (d_model,), "beta"),
             ("fusion.kv_ln.weight", (d_model,), "gamma"), ("fusion.kv_ln.bias", (d_model,), "beta"),
             ("fusion.Wq.weight", (d_model, d_model), "w"), ("fusion.Wk.weight", (d_model, d_model), "w"),
             ("fusion.Wv.weight", (d_model, d_model), "w"), ("fusion.Wo.weight", (d_model, d_model), "w"),
             ("fusion.logit_temp", (len(mods),), "temp"), ("fusion.logit_bias", (len(mods),), "temp"),
             ("cls", (1, 1, d_model), "randn")]
    for l in range(time_layers):
        p = f"temporal.layers.{l}"
        spec += [(f"{p}.self_attn.in_proj_weight", (3 * d_model, d_model), "w"),
                 (f"{p}.self_attn.in_proj_bias", (3 * d_model,), "beta"),
                 (f"{p}.self_attn.out_proj.weight", (d_model, d_model), "w"),
                 (f"{p}.self_attn.out_proj.bias", (d_model,), "beta"),
                 (f"{p}.linear1.weight", (ffn, d_model), "w"), (f"{p}.linear1.bias", (ffn,), "beta"),
                 (f"{p}.linear2.weight", (d_model, ffn), "w"), (f"{p}.linear2.bias", (d_model,), "beta"),
                 (f"{p}.norm1.weight", (d_model,), "gamma"), (f"{p}.norm1.bias", (d_model,), "beta"),
                 (f"{p}.norm2.weight", (d_model,), "gamma"), (f"{p}.norm2.bias", (d_model,), "beta")]
    return spec


def make_state_dict(dims_map_raw: Dict[str, int], dims_map_diff: Dict[str, int], seed: int = 0,
                    d_model: int = 256, time_layers: int = 4) -> Dict[str, torch.Tensor]:
    """Random-init weights with the reference's state-dict keys/shapes (no checkpoint ships with
    the reference, .gitignore:18-22). Fan-in-uniform like torch's default Conv/Linear init, but
    norm affines / biases / logit temp+bias get non-trivial values so that parity tests exercise
    them (torch's defaults of 1/0 would hide bugs). numpy PCG64 => identical on every machine."""
    g = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape, kind in state_dict_spec(dims_map_raw, dims_map_diff, d_model, time_layers):
        if kind == "w":
            fan_in = int(np.prod(shape[1:]))
            b = 1.0 / math.sqrt(fan_in)
            a = g.uniform(-b, b, size=shape)
        elif kind == "gamma":
            a = 1.0 + 0.1 * g.standard_normal(shape)
        elif kind == "beta":
            a = 0.05 * g.standard_normal(shape)
        elif kind == "randn":
            a = g.standard_normal(shape)
        elif kind == "temp":
            a = 0.3 * g.standard_normal(shape)
        else:
            raise AssertionError(kind)
        sd[key] = torch.from_numpy(np.asarray(a, dtype=np.float32))
    sd["pos_enc.pe"] = sinusoidal_pe(5000, d_model)
    return sd
