"""Drop-ins for the scoring functions of the reference (same names, arguments and return values):

  extract_window_features(model, dataloader, device, save_path=None)        eval.py:168-206
  compute_temporal_coherence_scores(features)                               eval.py:209-226
  compute_action_consistency_scores(features, centroids, label_dict)        eval.py:229-257
  build_train_centroids_subset(model, loader, label_dict, device, ...)      utils.py:1018-1045
  TCL(temperature, k1, k2)(projections, targets)                            losses.py:6-34 (forward only)

All arithmetic runs in libtag_b200.so kernels (K2/K3/K4/N1); Python only groups window names into
per-video segments, exactly the bookkeeping the reference does with dicts.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .synth import ACTION_CLASSES

_util_handles: Dict[int, C.c_void_p] = {}


def util_handle(device) -> C.c_void_p:
    """A weight-less native handle per device for the model-independent kernels (K3/K4/N1/N2)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.TagError("TAG scoring kernels need a CUDA (sm_100a) device; there is no CPU path")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _util_handles:
        lib = _lib.load()
        cfg = _lib.tag_config()
        cfg.n_modalities = 1
        cfg.raw_dims[0], cfg.diff_dims[0], cfg.kinds[0] = 1, 0, _lib.KIND_PLAIN
        cfg.d_model, cfg.n_heads, cfg.n_layers, cfg.ffn_dim, cfg.n_blocks, cfg.conv_kernel = 256, 8, 0, 1024, 4, 5
        cfg.precision, cfg.max_windows, cfg.max_T, cfg.device = _lib.PRECISION_FP32, 1, 1, idx
        h = C.c_void_p()
        _lib.check(None, lib.tag_create(C.byref(h), C.byref(cfg)), "tag_create")
        _util_handles[idx] = h
    return _util_handles[idx]


def _cuda_device(*tensors) -> torch.device:
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    if not torch.cuda.is_available():
        raise _lib.TagError("TAG scoring kernels need a CUDA (sm_100a) device; there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _canonicalize_class(name: str) -> str:
    """reference eval.py:36-45."""
    for cls in ACTION_CLASSES:
        if name.lower() == cls.lower():
            return cls
    return {"soccerjuggling": "SoccerJuggling", "tennisswing": "TennisSwing"}.get(name.lower(), name)


# ---------------------------------------------------------------------------------------------
def extract_window_features(model, dataloader: Iterable, device=None, save_path: Optional[str] = None) -> dict:
    """eval.py:168-206. `dataloader` yields `(feats[B,T,D], cls_names, vid_names)` or None."""
    seqs, frames, cls_all, vid_all = [], [], [], []
    model.eval()
    dev = torch.device(device) if device is not None else next(model.parameters()).device
    with torch.no_grad():
        for batch in dataloader:
            if batch is None:
                continue
            feats, cls_names, vid_names = batch
            feats = feats.to(dev, non_blocking=True)
            seq, frm, _tok = model(feats)
            seqs.append(seq.cpu())
            frames.append(frm.cpu())
            cls_all.extend(cls_names)
            vid_all.extend(vid_names)
    features = {"seq_embeds": torch.cat(seqs, 0), "frame_embeds": torch.cat(frames, 0),
                "cls_names": cls_all, "vid_names": vid_all}
    if save_path:
        torch.save(features, save_path)
        print(f"Saved features to {save_path}")
    return features


def _segments(vid_names: Sequence[str]) -> Tuple[List[str], List[int], List[int]]:
    """video ids in first-occurrence order (dict order of the reference), a permutation that makes each
    video's windows contiguous, and the segment offsets [V+1]."""
    order: Dict[str, List[int]] = {}
    for i, n in enumerate(vid_names):
        order.setdefault(os.path.splitext(n)[0], []).append(i)
    perm, offs = [], [0]
    for idx in order.values():
        perm.extend(idx)
        offs.append(len(perm))
    return list(order.keys()), perm, offs


def compute_temporal_coherence_scores(features: dict) -> Dict[str, float]:
    """eval.py:209-226: per window mean_t ||f_{t+1}-f_t|| over frame embeds (CLS dropped), per video mean."""
    lib = _lib.load()
    frame_embeds = features["frame_embeds"]
    dev = _cuda_device(frame_embeds)
    h = util_handle(dev)
    fe = frame_embeds.to(dev, dtype=torch.float32).contiguous()
    N, S, D = fe.shape
    if D != 256:
        raise ValueError("frame_embeds must be [N, T+1, 256]")
    vids, perm, offs = _segments(features["vid_names"])
    stream = torch.cuda.current_stream(dev).cuda_stream
    tcw = torch.empty(N, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(h, lib.tag_window_tc(h, fe.data_ptr(), N, S, tcw.data_ptr(), stream), "tag_window_tc")
        tcw = tcw.index_select(0, torch.tensor(perm, device=dev, dtype=torch.long)).contiguous()
        seg = torch.tensor(offs, device=dev, dtype=torch.int64)
        V = len(vids)
        tc = torch.empty(V, device=dev, dtype=torch.float32)      # per-video mean through K4 (TC-only mode)
        _lib.check(h, lib.tag_score(h, None, tcw.data_ptr(), seg.data_ptr(), None, None, 1, V, None, tc.data_ptr(),
                                    stream), "tag_score")
    tc = tc.cpu().tolist()
    return {v: float(s) for v, s in zip(vids, tc) if s == s}       # NaN = window had < 2 frames (skipped, eval.py:220)


def compute_action_consistency_scores(features: dict, centroids, label_dict: dict) -> Dict[str, float]:
    """eval.py:229-257."""
    lib = _lib.load()
    seq = features["seq_embeds"]
    dev = _cuda_device(seq, centroids)
    h = util_handle(dev)
    seq = seq.to(dev, dtype=torch.float32).contiguous()
    cen = torch.as_tensor(centroids).detach().to(dev, dtype=torch.float32).contiguous()
    vids, perm, offs = _segments(features["vid_names"])
    # class of a video = class of its LAST window (dict overwrite at eval.py:242)
    last_cls: Dict[str, str] = {}
    for n, c in zip(features["vid_names"], features["cls_names"]):
        last_cls[os.path.splitext(n)[0]] = _canonicalize_class(c)
    C_ = int(cen.shape[0])
    labels = []
    for v in vids:
        idx = label_dict.get(last_cls[v], -1)
        labels.append(idx if 0 <= idx < C_ else -1)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        seq = seq.index_select(0, torch.tensor(perm, device=dev, dtype=torch.long)).contiguous()
        seg = torch.tensor(offs, device=dev, dtype=torch.int64)
        lab = torch.tensor(labels, device=dev, dtype=torch.int32)
        V = len(vids)
        ac = torch.empty(V, device=dev, dtype=torch.float32)
        _lib.check(h, lib.tag_score(h, seq.data_ptr(), None, seg.data_ptr(), lab.data_ptr(), cen.data_ptr(), C_, V,
                                    ac.data_ptr(), None, stream), "tag_score")
    ac = ac.cpu().tolist()
    return {v: float(s) for v, s, l in zip(vids, ac, labels) if l >= 0}


def centroid_accumulate(z: torch.Tensor, y: torch.Tensor, sums_counts: torch.Tensor):
    """K3: sums_counts[C,257] += per-class (sum of z rows || count)."""
    lib = _lib.load()
    h = util_handle(z.device)
    C_ = sums_counts.shape[0]
    stream = torch.cuda.current_stream(z.device).cuda_stream
    with torch.cuda.device(z.device):
        _lib.check(h, lib.tag_centroid_accumulate(h, z.data_ptr(), y.data_ptr(), z.shape[0], C_, sums_counts.data_ptr(),
                                                  stream), "tag_centroid_accumulate")


def centroid_finalize(sums_counts: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    dev = sums_counts.device
    h = util_handle(dev)
    C_ = sums_counts.shape[0]
    cen = torch.empty(C_, 256, device=dev, dtype=torch.float32)
    cnt = torch.empty(C_, device=dev, dtype=torch.float32)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(h, lib.tag_centroid_finalize(h, sums_counts.data_ptr(), C_, cen.data_ptr(), cnt.data_ptr(), stream),
                   "tag_centroid_finalize")
    return cen, cnt


def allreduce_centroid_sums(sums_counts: torch.Tensor, group=None) -> torch.Tensor:
    """The path's ONE collective: sum the packed [C,257] buffer across ranks, in place,
    on the compute stream (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums_counts, op=dist.ReduceOp.SUM, group=group)
    return sums_counts


@torch.no_grad()
def build_train_centroids_subset(model, small_loader: Iterable, label_dict: dict, device, feature_selector=None,
                                 group=None):
    """utils.py:1018-1045; returns (centroids[C,256], counts[C]) on `device`. When torch.distributed is
    initialised, the class sums/counts are all-reduced across ranks before normalisation."""
    model.eval()
    dev = torch.device(device)
    C_ = len(label_dict)
    sc = torch.zeros(C_, 257, device=dev, dtype=torch.float32)
    for packed in small_loader:
        if packed is None:
            continue
        feats, cls_names, _ = packed
        feats = feats.to(dev, non_blocking=True)
        if feature_selector is not None:
            feats = feature_selector.select(feats)
        z, _, _ = model(feats)
        y = torch.tensor([label_dict[c] for c in cls_names], device=dev, dtype=torch.int32)
        centroid_accumulate(z.contiguous(), y, sc)
    allreduce_centroid_sums(sc, group)
    centroids, counts = centroid_finalize(sc)
    model.train()
    return centroids, counts


from .losses import TCL  # noqa: E402,F401  (losses.py:6-34; kept importable from here)


def write_video_scores(path: str, ac: Dict[str, float], tc: Dict[str, float]) -> Dict[str, dict]:
    """`video_scores.json` exactly as eval.py:439-451 writes it (consumed by process_scores.py:113-127)."""
    combined = {}
    for vid in sorted(set(ac) | set(tc)):
        entry = {}
        if vid in ac:
            entry["ac"] = ac[vid]
        if vid in tc:
            entry["tc"] = tc[vid]
        combined[vid] = entry
    with open(path, "w") as f:
        json.dump(combined, f, indent=2)
    return combined
