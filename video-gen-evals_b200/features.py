"""Window feature construction on the GPU (K1) behind the reference's data-layer names.

Mirrors (reference utils.py): `ModalityStats` (:570-586), `compute_stats_from_npz` (:595-801, here
`compute_stats_from_videos` on device-resident arrays), `WindowDataset` (:345-523) and `safe_collate`
(:104-110). File IO / directory scanning (utils.py:221-341, :384-393, :409-452) is out of scope
(SURVEY.md §2): inputs are packed per-frame arrays already resident in HBM (`DeviceVideos`).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, fields
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .synth import VideoBatch, enumerate_windows

# modality name -> field prefix used by the reference's ModalityStats
STAT_PREFIX = {"vit": "vit", "global": "gori", "pose": "pose", "beta": "beta", "kp2d": "keypoints",
               "clip": "clip", "dino": "dino"}


@dataclass
class ModalityStats:
    """Same fields, same order as reference utils.py:570-586 (None = modality absent)."""
    vit_raw_mean: Optional[torch.Tensor] = None; vit_raw_std: Optional[torch.Tensor] = None
    gori_raw_mean: Optional[torch.Tensor] = None; gori_raw_std: Optional[torch.Tensor] = None
    pose_raw_mean: Optional[torch.Tensor] = None; pose_raw_std: Optional[torch.Tensor] = None
    beta_raw_mean: Optional[torch.Tensor] = None; beta_raw_std: Optional[torch.Tensor] = None
    keypoints_raw_mean: Optional[torch.Tensor] = None; keypoints_raw_std: Optional[torch.Tensor] = None
    clip_raw_mean: Optional[torch.Tensor] = None; clip_raw_std: Optional[torch.Tensor] = None
    dino_raw_mean: Optional[torch.Tensor] = None; dino_raw_std: Optional[torch.Tensor] = None
    vit_diff_mean: Optional[torch.Tensor] = None; vit_diff_std: Optional[torch.Tensor] = None
    gori_diff_mean: Optional[torch.Tensor] = None; gori_diff_std: Optional[torch.Tensor] = None
    pose_diff_mean: Optional[torch.Tensor] = None; pose_diff_std: Optional[torch.Tensor] = None
    beta_diff_mean: Optional[torch.Tensor] = None; beta_diff_std: Optional[torch.Tensor] = None
    keypoints_diff_mean: Optional[torch.Tensor] = None; keypoints_diff_std: Optional[torch.Tensor] = None
    clip_diff_mean: Optional[torch.Tensor] = None; clip_diff_std: Optional[torch.Tensor] = None
    dino_diff_mean: Optional[torch.Tensor] = None; dino_diff_std: Optional[torch.Tensor] = None

    @classmethod
    def from_dict(cls, d: Dict[str, torch.Tensor]) -> "ModalityStats":
        names = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in d.items() if k in names})


def infer_dims_from_stats(stats) -> Tuple[Dict[str, int], Dict[str, int]]:
    """reference eval.py:104-133."""
    g = (lambda k: getattr(stats, k, None)) if not isinstance(stats, dict) else stats.get
    raw, diff = {}, {}
    for m in ("vit", "global", "pose", "beta"):
        p = STAT_PREFIX[m]
        raw[m] = g(f"{p}_raw_mean").shape[0] if g(f"{p}_raw_mean") is not None else 0
        diff[m] = g(f"{p}_diff_mean").shape[0] if g(f"{p}_diff_mean") is not None else 0
    for m in ("kp2d", "clip", "dino"):
        p = STAT_PREFIX[m]
        if g(f"{p}_raw_mean") is not None:
            raw[m] = g(f"{p}_raw_mean").shape[0]
            diff[m] = g(f"{p}_diff_mean").shape[0] if g(f"{p}_diff_mean") is not None else 0
    return raw, diff


def stats_vectors(stats, modalities: Sequence[str], device) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """(mean[D], std[D]) in feats column order [raw blocks || diff blocks] (utils.py:496-514)."""
    if stats is None:
        return None, None
    g = (lambda k: getattr(stats, k)) if not isinstance(stats, dict) else (lambda k: stats[k])
    means, stds = [], []
    for kind in ("raw", "diff"):
        for m in modalities:
            p = STAT_PREFIX[m]
            means.append(torch.as_tensor(g(f"{p}_{kind}_mean"), dtype=torch.float32).reshape(-1))
            stds.append(torch.as_tensor(g(f"{p}_{kind}_std"), dtype=torch.float32).reshape(-1))
    return torch.cat(means).to(device).contiguous(), torch.cat(stds).to(device).contiguous()


class DeviceVideos:
    """Packed per-frame arrays of V videos resident on one GPU + the `tag_videos` view of them."""

    def __init__(self, vb: VideoBatch, modalities: Sequence[str], device, frame_offset: Optional[torch.Tensor] = None):
        """frame_offset: optional device int64 [V+1] tensor prepared by the caller (the streaming path stages it on its
        copy stream, so that no host->device copy is ever queued on the compute stream)."""
        self.device = torch.device(device)
        self.modalities = list(modalities)
        self.vb = vb if vb.pose.device == self.device else vb.to(self.device)
        F = self.vb.n_frames
        src = {"vit": self.vb.vit, "global": self.vb.gori.reshape(F, -1), "pose": self.vb.pose.reshape(F, -1),
               "beta": self.vb.betas, "kp2d": self.vb.kp, "clip": self.vb.clip, "dino": self.vb.dino}
        self.src = []
        for m in self.modalities:
            t = src[m]
            if t is None:
                raise ValueError(f"modality '{m}' requested but the video batch has no such array")
            self.src.append(t.to(torch.float32).contiguous())
        self.lengths = [self.vb.offsets[i + 1] - self.vb.offsets[i] for i in range(self.vb.n_videos)]
        V = self.vb.n_videos
        L0 = self.lengths[0] if V else 0
        self.uniform_len = L0 if V and all(x == L0 for x in self.lengths) else None
        if frame_offset is not None:
            self.frame_offset = frame_offset
        elif self.uniform_len is not None and self.device.type == "cuda" and self.vb.offsets[0] == 0:
            # clips of one length: the offsets are an arithmetic progression — generated on the device, no host copy
            self.frame_offset = torch.arange(V + 1, device=self.device, dtype=torch.int64) * int(L0)
        else:
            self.frame_offset = torch.tensor(self.vb.offsets, dtype=torch.int64, device=self.device)
        self.c = _lib.tag_videos()
        for i, t in enumerate(self.src):
            self.c.src[i] = t.data_ptr()
        self.c.frame_offset = self.frame_offset.data_ptr()
        self.c.n_videos = self.vb.n_videos

    @property
    def n_videos(self) -> int:
        return self.vb.n_videos


class FeatureFuser:
    """Owns a weight-less native handle configured for a modality set; runs K1 (tag_feature_fuse)."""

    def __init__(self, dims_map_raw: Dict[str, int], dims_map_diff: Dict[str, int], device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.TagError("FeatureFuser needs a CUDA (sm_100a) device; there is no CPU path")
        self.modalities = list(dims_map_raw.keys())
        self.dims_raw = {m: int(dims_map_raw[m]) for m in self.modalities}
        self.dims_diff = {m: int(dims_map_diff[m]) for m in self.modalities}
        self.D = sum(self.dims_raw.values()) + sum(self.dims_diff.values())
        lib = _lib.load()
        cfg = _lib.tag_config()
        cfg.n_modalities = len(self.modalities)
        for i, m in enumerate(self.modalities):
            cfg.raw_dims[i], cfg.diff_dims[i], cfg.kinds[i] = self.dims_raw[m], self.dims_diff[m], _lib.KIND_OF[m]
        cfg.d_model, cfg.n_heads, cfg.n_layers, cfg.ffn_dim, cfg.n_blocks, cfg.conv_kernel = 256, 8, 0, 1024, 4, 5
        cfg.precision, cfg.max_windows, cfg.max_T = _lib.PRECISION_FP32, 1, 1
        cfg.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._h = C.c_void_p()
        _lib.check(None, lib.tag_create(C.byref(self._h), C.byref(cfg)), "tag_create")

    def __del__(self):
        try:
            if self._h:
                _lib.load().tag_destroy(self._h)
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def fuse(self, dv: DeviceVideos, win_video: torch.Tensor, win_start: torch.Tensor, T: int,
             mean: Optional[torch.Tensor], std: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (feats [N,T,D] fp32 on device, flags int32[1] = #frames in the Procrustes reflection regime)."""
        lib = _lib.load()
        N = int(win_video.numel())
        feats = torch.empty(N, T, self.D, device=self.device, dtype=torch.float32)
        flags = torch.zeros(1, device=self.device, dtype=torch.int32)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(self._h, lib.tag_feature_fuse(self._h, C.byref(dv.c), _lib.ptr(mean), _lib.ptr(std),
                                                    win_video.data_ptr(), win_start.data_ptr(), N, T,
                                                    feats.data_ptr(), flags.data_ptr(), stream), "tag_feature_fuse")
        return feats, flags


def compute_stats_from_videos(vb_or_dv, dims_map_raw: Dict[str, int], dims_map_diff: Dict[str, int], device,
                              eps: float = 1e-6, fuser: Optional[FeatureFuser] = None) -> ModalityStats:
    """`compute_stats_from_npz` (reference utils.py:595-801) on device-resident videos: un-normalised
    [raw || diff] rows of WHOLE videos (diffs across the full sequence, :717-732) from K1, float64 column
    sums / sums of squares on the GPU (:589-593), then mean = s/n, std = sqrt(max(ss/n - mean^2, 0) + eps)
    (:746-750)."""
    lib = _lib.load()
    fuser = fuser or FeatureFuser(dims_map_raw, dims_map_diff, device)
    dv = vb_or_dv if isinstance(vb_or_dv, DeviceVideos) else DeviceVideos(vb_or_dv, fuser.modalities, device)
    dev = fuser.device
    D = fuser.D
    s1 = torch.zeros(D, device=dev, dtype=torch.float64)
    s2 = torch.zeros(D, device=dev, dtype=torch.float64)
    n_rows = 0
    by_len: Dict[int, List[int]] = {}
    for v, L in enumerate(dv.lengths):
        by_len.setdefault(int(L), []).append(v)
    stream = torch.cuda.current_stream(dev).cuda_stream
    for L, vids in by_len.items():
        if L <= 0:
            continue
        for i in range(0, len(vids), 256):
            chunk = vids[i:i + 256]
            wv = torch.tensor(chunk, dtype=torch.int32, device=dev)
            ws = torch.zeros(len(chunk), dtype=torch.int32, device=dev)
            feats, _ = fuser.fuse(dv, wv, ws, L, None, None)
            rows = len(chunk) * L
            with torch.cuda.device(dev):
                _lib.check(fuser.handle, lib.tag_stats_accumulate(fuser.handle, feats.data_ptr(), rows, D, s1.data_ptr(),
                                                                  s2.data_ptr(), stream), "tag_stats_accumulate")
            n_rows += rows
    n = max(1, n_rows)
    mean = s1 / n
    var = s2 / n - mean ** 2
    std = torch.sqrt(torch.clamp(var, min=0.0) + eps)
    mean32, std32 = mean.to(torch.float32).cpu(), std.to(torch.float32).cpu()
    out, off = {}, 0
    for kind, dims in (("raw", fuser.dims_raw), ("diff", fuser.dims_diff)):
        for m in fuser.modalities:
            d = dims[m]
            if d > 0:
                out[f"{STAT_PREFIX[m]}_{kind}_mean"] = mean32[off:off + d].clone()
                out[f"{STAT_PREFIX[m]}_{kind}_std"] = std32[off:off + d].clone()
            off += d
    return ModalityStats.from_dict(out)


def safe_collate(batch):
    """reference utils.py:104-110."""
    batch = [b for b in batch if b is not None]
    if not batch:
        return None
    feats, cls_names, vids = zip(*batch)
    return torch.stack(feats, dim=0), list(cls_names), list(vids)


class WindowDataset:
    """Mirror of reference `WindowDataset` (utils.py:345-523) over device-resident videos.

    samples: [(video_index, start)]; `ds[i] -> (feats[T,D] fp32 (cuda), cls, name)`;
    `ds.batches(bs)` yields what `DataLoader(ds, bs, collate_fn=safe_collate)` yields in the reference:
    `(feats[B,T,D], cls_names, vid_names)` — one K1 launch per batch."""

    def __init__(self, samples: Sequence[Tuple[int, int]], clip_len: int = 32, stats=None, *, videos: DeviceVideos,
                 dims_map_raw: Dict[str, int], dims_map_diff: Dict[str, int], fuser: Optional[FeatureFuser] = None):
        self.samples = list(samples)
        self.clip_len = int(clip_len)
        self.videos = videos
        self.fuser = fuser or FeatureFuser(dims_map_raw, dims_map_diff, videos.device)
        self.stats = stats
        self.mean, self.std = stats_vectors(stats, self.fuser.modalities, videos.device)
        self._flags_acc = None

    @classmethod
    def all_windows(cls, videos: DeviceVideos, clip_len: int = 32, stride: int = 8, **kw) -> "WindowDataset":
        """reference `sample_all_windows_npz` (utils.py:888-911) + WindowDataset."""
        vids, starts = enumerate_windows(videos.lengths, clip_len, stride)
        return cls(list(zip(vids, starts)), clip_len, videos=videos, **kw)

    def __len__(self):
        return len(self.samples)

    def _fuse(self, idx: Sequence[int]) -> torch.Tensor:
        dev = self.videos.device
        wv = torch.tensor([self.samples[i][0] for i in idx], dtype=torch.int32, device=dev)
        ws = torch.tensor([self.samples[i][1] for i in idx], dtype=torch.int32, device=dev)
        feats, flags = self.fuser.fuse(self.videos, wv, ws, self.clip_len, self.mean, self.std)
        self._last_flags = flags
        self._flags_acc = flags if self._flags_acc is None else self._flags_acc + flags      # on the device: no sync here
        return feats

    @property
    def reflect_frames(self) -> int:
        """Frames (over all windows built so far) whose keypoint cross-covariance had det(H) < 0 — mirror-like consecutive
        frames, e.g. a detector's left/right swap. K1 reproduces the reference there too (polar-reflection closed form);
        the count is informational (a high count usually means noisy keypoints). Reading it synchronises."""
        return 0 if self._flags_acc is None else int(self._flags_acc.item())

    def __getitem__(self, i: int):
        v, _ = self.samples[i]
        return self._fuse([i])[0], self.videos.vb.cls_name(v), self.videos.vb.names[v]

    def batches(self, batch_size: int = 32) -> Iterator[Tuple[torch.Tensor, List[str], List[str]]]:
        for i in range(0, len(self.samples), batch_size):
            idx = range(i, min(len(self.samples), i + batch_size))
            feats = self._fuse(idx)
            yield (feats, [self.videos.vb.cls_name(self.samples[j][0]) for j in idx],
                   [self.videos.vb.names[self.samples[j][0]] for j in idx])
