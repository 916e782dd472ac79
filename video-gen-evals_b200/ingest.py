"""File ingest for the scoring path: the reference's on-disk inputs -> pinned host arrays -> the
device staging ring of `TagScorer.score_stream`, so that the end-to-end call can start at files.

What the reference does per WINDOW (utils.py:384-393, :409-452: `np.load(npz)` + zlib inflate of all four arrays + `np.load`
of keypoints.npy, once for every window of every video, inside DataLoader workers) happens here once per VIDEO, straight
into one pinned buffer per modality (no per-file temporaries for float32 inputs), on a thread pool (zlib and file reads
release the GIL). The window index is not built from files at all: windows are (video, start) pairs derived from the frame
counts (`window_table`, utils.py:888-911) — on the device when all clips have one length.

Layouts (the two the reference reads):
  generated  <mesh_dir>/<name>.npz, <kp_dir>/<stem>/keypoints.npy                    (eval.py:48-101, utils.py:411-412)
  real       <mesh_dir>/<Class>/<name>.npz, <kp_dir>/<Class>/<stem>/keypoints.npy   (utils.py:229-319, :413-414)
`.npz` members: pose [T,23,3,3], betas [T,10], global_orient [T,1,3,3], vit [T,1024] (extract_mesh.py:35-43, written with
np.savez_compressed); optional clip_embeddings.npz / dino_embeddings.npz with member `embeddings`.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import zipfile
from dataclasses import dataclass
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .synth import ACTION_CLASSES, VideoBatch

_ALIASES = {"soccerjuggling": "SoccerJuggling", "tennisswing": "TennisSwing"}


def canonicalize_class(name: str) -> str:
    """reference eval.py:36-45."""
    for cls in ACTION_CLASSES:
        if name.lower() == cls.lower():
            return cls
    return _ALIASES.get(name.lower(), name)


def class_from_filename(stem: str) -> str:
    """reference eval.py:55-74: first '_'-separated token that names an action class, else the first capitalised
    word-like token, else 'Unknown'."""
    parts = stem.split("_")
    for part in parts:
        canon = canonicalize_class(part)
        if canon in ACTION_CLASSES:
            return canon
    for part in parts:
        if part and part[0].isupper() and not part.isdigit() and len(part) > 3 and part.lower() not in ("videos", "npz"):
            return canonicalize_class(part)
    return "Unknown"


@dataclass
class FileItem:
    """one video on disk (the reference's VideoItem, utils.py:221-227, plus the side files)"""
    cls: str
    name: str            # file name with .npz
    path: str
    kp_path: Optional[str]
    clip_path: Optional[str] = None
    dino_path: Optional[str] = None
    length: int = 0


def _npy_header(f) -> Tuple[tuple, np.dtype, bool]:
    major, minor = np.lib.format.read_magic(f)
    if (major, minor) == (1, 0):
        shape, fortran, dtype = np.lib.format.read_array_header_1_0(f)
    else:
        shape, fortran, dtype = np.lib.format.read_array_header_2_0(f)
    return shape, dtype, fortran


def _npz_member_shape(path: str, member: str) -> tuple:
    with zipfile.ZipFile(path) as z:
        with z.open(member + ".npy") as f:
            return _npy_header(f)[0]


def _read_into(f, out: np.ndarray):
    """fill `out` (C-contiguous float32 view of a pinned buffer) from an open .npy stream positioned at its magic"""
    shape, dtype, fortran = _npy_header(f)
    n = int(np.prod(shape)) if len(shape) else 1
    if n != out.size:
        raise ValueError(f"array of shape {shape} does not fill a destination of {out.shape}")
    if dtype == np.float32 and not fortran:
        mv = memoryview(out.reshape(-1)).cast("B")
        got = 0
        while got < len(mv):
            k = f.readinto(mv[got:])
            if not k:
                raise EOFError("truncated .npy payload")
            got += k
    else:                                  # other dtypes / orders: one temporary, then a converting copy
        a = np.frombuffer(f.read(n * dtype.itemsize), dtype=dtype, count=n).reshape(shape, order="F" if fortran else "C")
        out[...] = a.reshape(out.shape).astype(np.float32)


class NpzIngest:
    def __init__(self, mesh_dir: str, kp_dir: Optional[str], generated: bool = True, clip_dir: Optional[str] = None,
                 dino_dir: Optional[str] = None, filter_classes: Optional[Sequence[str]] = None, threads: Optional[int] = None):
        self.mesh_dir, self.kp_dir, self.generated = mesh_dir, kp_dir, generated
        self.clip_dir, self.dino_dir = clip_dir, dino_dir
        self.filter_classes = set(filter_classes) if filter_classes is not None else None
        self.threads = threads or min(32, os.cpu_count() or 1)

    # ------------------------------------------------------------------ scanning (eval.py:48-101 / utils.py:273-319)
    def _side(self, root, cls, stem, fname):
        if root is None:
            return None
        return os.path.join(root, stem, fname) if self.generated else os.path.join(root, cls, stem, fname)

    def scan(self) -> List[FileItem]:
        items: List[FileItem] = []
        if self.generated:
            for f in sorted(os.listdir(self.mesh_dir)):
                if f.endswith(".npz"):
                    stem = os.path.splitext(f)[0]
                    cls = class_from_filename(stem)
                    items.append(FileItem(cls, f, os.path.join(self.mesh_dir, f), self._side(self.kp_dir, cls, stem, "keypoints.npy"),
                                          self._side(self.clip_dir, cls, stem, "clip_embeddings.npz"),
                                          self._side(self.dino_dir, cls, stem, "dino_embeddings.npz")))
        else:
            for cls in sorted(d for d in os.listdir(self.mesh_dir) if os.path.isdir(os.path.join(self.mesh_dir, d))):
                if self.filter_classes is not None and cls not in self.filter_classes:
                    continue
                for f in sorted(os.listdir(os.path.join(self.mesh_dir, cls))):
                    if f.endswith(".npz"):
                        stem = os.path.splitext(f)[0]
                        items.append(FileItem(cls, f, os.path.join(self.mesh_dir, cls, f), self._side(self.kp_dir, cls, stem, "keypoints.npy"),
                                              self._side(self.clip_dir, cls, stem, "clip_embeddings.npz"),
                                              self._side(self.dino_dir, cls, stem, "dino_embeddings.npz")))
        with cf.ThreadPoolExecutor(self.threads) as ex:
            lens = list(ex.map(lambda it: _npz_member_shape(it.path, "pose")[0], items))
        for it, L in zip(items, lens):
            it.length = int(L)
            if it.kp_path is not None and not os.path.exists(it.kp_path):
                raise FileNotFoundError(f"Expected keypoints at '{it.kp_path}' for video '{os.path.splitext(it.name)[0]}' but file does not exist.")
        return [it for it in items if it.length > 0]

    # ------------------------------------------------------------------ loading
    def load(self, items: Sequence[FileItem], pin: bool = True, classes: Sequence[str] = ACTION_CLASSES) -> VideoBatch:
        """-> VideoBatch of pinned host arrays (frames of all videos back to back). A keypoint file shorter / longer than the
        mesh arrays is cut / padded with its last frame (the reference slices both per window, utils.py:366-381)."""
        offs = [0]
        for it in items:
            offs.append(offs[-1] + it.length)
        F = offs[-1]
        has_clip = self.clip_dir is not None and all(it.clip_path and os.path.exists(it.clip_path) for it in items)
        has_dino = self.dino_dir is not None and all(it.dino_path and os.path.exists(it.dino_path) for it in items)

        def alloc(*shape):
            t = torch.empty(shape, dtype=torch.float32)
            return t.pin_memory() if pin and torch.cuda.is_available() else t

        pose, gori, betas, vit = alloc(F, 23, 3, 3), alloc(F, 1, 3, 3), alloc(F, 10), alloc(F, 1024)
        kp = alloc(F, 120)
        clip = alloc(F, 512) if has_clip else None
        dino = alloc(F, 768) if has_dino else None
        views = {"pose": pose.numpy(), "global_orient": gori.numpy(), "betas": betas.numpy(), "vit": vit.numpy()}
        kpn = kp.numpy()

        def one(i):
            it = items[i]
            a, b = offs[i], offs[i + 1]
            with zipfile.ZipFile(it.path) as z:
                for member, dst in views.items():
                    with z.open(member + ".npy") as f:
                        _read_into(f, dst[a:b])
            if it.kp_path is not None:
                arr = np.load(it.kp_path)
                n = min(arr.shape[0], b - a)
                kpn[a:a + n] = arr[:n].reshape(n, -1)
                if n < b - a:
                    kpn[a + n:b] = kpn[a + n - 1]
            else:
                kpn[a:b] = 0.0
            for path, dst in ((it.clip_path, clip), (it.dino_path, dino)):
                if dst is not None:
                    with zipfile.ZipFile(path) as z, z.open("embeddings.npy") as f:
                        _read_into(f, dst.numpy()[a:b])

        with cf.ThreadPoolExecutor(self.threads) as ex:
            list(ex.map(one, range(len(items))))
        cidx = {c: i for i, c in enumerate(classes)}
        return VideoBatch(pose, gori, betas, vit, kp, offs, [cidx.get(it.cls, -1) for it in items], [it.name for it in items],
                          clip, dino, list(classes))

    def batches(self, items: Optional[Sequence[FileItem]] = None, videos_per_batch: int = 5000, prefetch: int = 1) -> Iterator[VideoBatch]:
        """pinned VideoBatches of `videos_per_batch` videos, the next one(s) loading on a background thread while the caller
        (score_stream) consumes the current one"""
        items = list(items) if items is not None else self.scan()
        chunks = [items[i:i + videos_per_batch] for i in range(0, len(items), videos_per_batch)]
        with cf.ThreadPoolExecutor(1) as bg:
            pending = []
            for c in chunks:
                pending.append(bg.submit(self.load, c))
                if len(pending) > prefetch:
                    yield pending.pop(0).result()
            for p in pending:
                yield p.result()
