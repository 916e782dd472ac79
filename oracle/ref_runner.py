"""The UNMODIFIED reference as a checker and CPU baseline — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference (XThomasBU/video-gen-evals) is pure Python: nothing to compile. `build_ref()` — called by
`__graft_entry__.build()` in the build container, where /root/reference exists — places the reference's own
`model.py`, `utils.py`, `eval.py`, `losses.py`, `process_scores.py` (+ the human-score table) as ONE archive into the git-ignored
`oracle/_ref/` (not gpurun-ignored, so it travels to the GPU box next to the built .so); `load_ref()` unpacks it into a scratch
directory outside the repository and imports from there. Nothing under `oracle/_ref/` is ever committed, imported by the
product package, or modified.

Users (and only these): `tests/`, `bench.py --impl reference` / `cpu_baseline`. `load_ref()` returns None when
`oracle/_ref/` is absent (a fresh clone without the reference): callers then fall back to the oracle port
(`oracle/tag_oracle.py`) and say so (`cpu_baseline.kind = "port"`).

`train.py` is deliberately not taken: importing it mutates env/seeds and creates directories (train.py:12-13,
:60-66, :114-115); nothing on the scoring path needs it.
"""
from __future__ import annotations

import importlib
import os
import shutil
import sys
import tempfile
import time
from types import SimpleNamespace
from typing import Dict, Optional, Sequence

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(REF_DIR, "reference_modules.tar.gz")
FILES = ("model.py", "utils.py", "eval.py", "losses.py", "process_scores.py", "TAG_final_human_scores.json")


def build_ref(src: Optional[str] = None) -> Optional[str]:
    """Pack the reference's scoring-path modules, byte-identical, into ONE archive under the git-ignored oracle/_ref/
    (it travels to the GPU box with the built .so; no loose copies of reference sources lie in the tree). No-op when the
    reference is not present (GPU box: the archive arrived with the snapshot)."""
    import tarfile
    src = src or os.environ.get("TAG_REFERENCE", "/root/reference")
    if not os.path.isdir(src):
        return ARCHIVE if os.path.exists(ARCHIVE) else None
    os.makedirs(REF_DIR, exist_ok=True)
    for f in os.listdir(REF_DIR):                      # loose files of an earlier layout
        if f != os.path.basename(ARCHIVE):
            os.remove(os.path.join(REF_DIR, f))
    tmp = ARCHIVE + ".tmp"
    with tarfile.open(tmp, "w:gz") as tar:
        for f in FILES:
            tar.add(os.path.join(src, f), arcname=f)
    os.replace(tmp, ARCHIVE)
    return ARCHIVE


_cached = None
_unpacked = None


def unpack_dir() -> Optional[str]:
    """The archive extracted into a per-archive scratch directory (outside the repository), or None."""
    global _unpacked
    import hashlib
    import tarfile
    if _unpacked is not None:
        return _unpacked
    if not os.path.exists(ARCHIVE):
        return None
    with open(ARCHIVE, "rb") as f:
        tag = hashlib.sha256(f.read()).hexdigest()[:16]
    d = os.path.join(tempfile.gettempdir(), f"tag_reference_{tag}_{os.getuid()}")
    if not all(os.path.exists(os.path.join(d, f)) for f in FILES):
        tmp = tempfile.mkdtemp(prefix="tag_reference_unpack_")
        with tarfile.open(ARCHIVE, "r:gz") as tar:
            tar.extractall(tmp, filter="data")
        try:
            os.replace(tmp, d)
        except OSError:                                # another process won the race
            shutil.rmtree(tmp, ignore_errors=True)
    _unpacked = d
    return d


def human_scores_path() -> Optional[str]:
    """300 TAG-Bench file names + human MOS (eval.py:297-347), from the archive."""
    d = unpack_dir()
    return None if d is None else os.path.join(d, "TAG_final_human_scores.json")


def load_ref():
    """-> namespace(model, utils, eval, losses, process_scores) of the unmodified reference, or None."""
    global _cached
    if _cached is not None:
        return _cached
    d = unpack_dir()
    if d is None:
        return None
    if d not in sys.path:
        sys.path.insert(0, d)                # the reference's modules import each other by bare name (eval.py:7, :16)
    mods = {}
    for name in ("model", "utils", "eval", "losses", "process_scores"):
        m = importlib.import_module(name)
        if os.path.dirname(os.path.abspath(m.__file__)) != os.path.abspath(d):
            raise ImportError(f"module '{name}' resolved to {m.__file__}, not to the unpacked reference")
        mods[name] = m
    _cached = SimpleNamespace(**mods)
    return _cached


# ---------------------------------------------------------------------------------------------
# synthetic sets in the reference's on-disk layout (what extract_mesh.py:25-44 / DWpose write)
# ---------------------------------------------------------------------------------------------
def write_set(vb, mesh_dir: str, kp_dir: str, generated: bool, clip_dir=None, dino_dir=None, names=None):
    """real layout: <mesh_dir>/<Class>/<name>.npz + <kp_dir>/<Class>/<stem>/keypoints.npy;
    generated layout (utils.py:411): <mesh_dir>/<name>.npz + <kp_dir>/<stem>/keypoints.npy."""
    for v in range(vb.n_videos):
        d = vb.video(v)
        cls = vb.cls_name(v)
        name = names[v] if names is not None else vb.names[v]
        stem = os.path.splitext(name)[0]
        mdir = mesh_dir if generated else os.path.join(mesh_dir, cls)
        os.makedirs(mdir, exist_ok=True)
        np.savez(os.path.join(mdir, name), pose=d["pose"].numpy(), betas=d["betas"].numpy(),
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


def scratch_dir(prefix="tag_ref_") -> str:
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    return tempfile.mkdtemp(prefix=prefix, dir=base)


def stats_object(ref, stats: Dict[str, torch.Tensor]):
    """{'vit_raw_mean': ...} -> the reference's ModalityStats dataclass (utils.py:570-586)."""
    from dataclasses import fields
    kw = {f.name: None for f in fields(ref.utils.ModalityStats)}
    kw.update({k: torch.as_tensor(v, dtype=torch.float32) for k, v in stats.items() if k in kw})
    return ref.utils.ModalityStats(**kw)


def reference_model(ref, sd, dims_raw, dims_diff):
    """The reference's HumanActionScorer with our seeded state dict (strict: the key contract), eval mode."""
    mdl = ref.model.HumanActionScorer(dims_raw, dims_diff)
    mdl.load_state_dict(sd, strict=True)
    return mdl.eval()


def generated_loader(ref, gen_dir: str, kp_dir: str, stats, clip_len: int, stride: int, batch_size=32, workers=0,
                     clip_dir=None, dino_dir=None):
    """eval.py:394-418: generated-mesh dataset -> all windows -> WindowDataset -> DataLoader (batch 32)."""
    from torch.utils.data import DataLoader
    gen_ds = ref.eval.create_dataset_from_generated_meshes(gen_dir)
    samples = ref.utils.sample_all_windows_npz(gen_ds, clip_len=clip_len, stride=stride)
    wds = ref.utils.WindowDataset(samples=samples, clip_len=clip_len, stats=stats, keypoint_dir=kp_dir,
                                  clip_dir=clip_dir, dino_dir=dino_dir)
    loader = DataLoader(wds, batch_size=batch_size, shuffle=False, num_workers=workers,
                        collate_fn=ref.utils.safe_collate, persistent_workers=False)
    return loader, samples


def reference_scoring_pass(ref, mdl, loader, centroids, label_dict, device="cpu"):
    """The timed unit of the CPU baseline = eval.py:421-437 on an existing loader: the reference's own
    `WindowDataset.__getitem__` (file read + per-window deltas + z-score) in the loader, `extract_window_features`
    (model forward), `compute_action_consistency_scores`, `compute_temporal_coherence_scores`.
    Returns (seconds, ac dict, tc dict, features)."""
    t0 = time.perf_counter()
    features = ref.eval.extract_window_features(mdl, loader, device=device)
    ac = ref.eval.compute_action_consistency_scores(features, centroids, label_dict)
    tc = ref.eval.compute_temporal_coherence_scores(features)
    return time.perf_counter() - t0, ac, tc, features
