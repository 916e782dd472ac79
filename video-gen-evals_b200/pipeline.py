"""The fused, device-resident TAG scoring pipeline: what `python eval.py` (reference eval.py:350-466)
does between "arrays loaded" and "video_scores.json", as three native calls per batch of videos:

    tag_encode_windows   K1 feature fuse + K2 encoder (+ fused per-window TC), chunked in HBM
    tag_centroid_*       K3 per-action centroid sums -> [allreduce across ranks] -> normalise
    tag_score            K4 per-video AC (distance to centroid) and TC (mean of window TCs)

Multi-GPU: videos are block-sharded across ranks (`shard_range`), every rank scores
its own block with no data-path communication; the only collective is the all-reduce of the packed
[C,257] centroid sums/counts buffer during the centroid build.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .features import DeviceVideos, stats_vectors
from .model import HumanActionScorer
from .scoring import allreduce_centroid_sums, centroid_accumulate, centroid_finalize, util_handle
from .synth import VideoBatch


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of `n` videos owned by `rank` (keeps a video's windows on one rank)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def window_table(lengths: Sequence[int], clip_len: int, stride: int):
    """(win_video int32[N], win_start int32[N], seg_offsets int64[V+1]) for `sample_all_windows_npz`
    (reference utils.py:888-911), vectorised."""
    L = np.asarray(lengths, dtype=np.int64)
    stride = max(1, int(stride))
    nwin = np.where(L < clip_len, 1, (L - clip_len) // stride + 1)
    seg = np.zeros(len(L) + 1, dtype=np.int64)
    np.cumsum(nwin, out=seg[1:])
    win_video = np.repeat(np.arange(len(L), dtype=np.int32), nwin)
    win_start = (np.arange(seg[-1], dtype=np.int64) - seg[:-1][win_video]) * stride
    return win_video.astype(np.int32), win_start.astype(np.int32), seg


def block_plan(lengths: Sequence[int], clip_len: int, stride: int, max_windows: int, first: bool,
               n_sms: int = 148) -> List[Tuple[int, int]]:
    """Cut a batch of videos into contiguous blocks `(lo, hi)` for `TagScorer.score_stream`: every block is at most one
    encoder pass (<= max_windows windows; a video's windows never straddle two blocks), and the FIRST batch of a stream
    starts with short blocks (1, 3, 15 waves of the CTA-pair GEMM) so that only a sliver of host->device copy is exposed
    before the encoder has work. A wave = (n_sms // 2) tile pairs of 256 rows = (n_sms // 2) * 256 / clip_len windows."""
    seg = window_table(lengths, clip_len, stride)[2]                           # cumulative windows per video
    wave = max(1, (n_sms // 2) * 256 // max(1, clip_len))
    cap = max(1, int(max_windows))
    steps = [wave, 3 * wave, 15 * wave] if first else []
    blocks, lo, V = [], 0, len(lengths)
    while lo < V:
        want = min(steps.pop(0), cap) if steps else cap
        hi = int(np.searchsorted(seg, seg[lo] + want, side="right")) - 1       # last video that still fits
        hi = max(hi, lo + 1)
        blocks.append((lo, min(hi, V)))
        lo = min(hi, V)
    return blocks


class TagScorer:
    def __init__(self, model: HumanActionScorer, stats, clip_len: int = 32, stride: int = 8, device=None):
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise _lib.TagError("TagScorer needs a CUDA (sm_100a) device; there is no CPU path")
        self.clip_len, self.stride = int(clip_len), int(stride)
        self.mean, self.std = stats_vectors(stats, model.modalities, self.device)
        self.model.eval()

    # ------------------------------------------------------------------
    def to_device(self, vb: VideoBatch) -> DeviceVideos:
        return DeviceVideos(vb, self.model.modalities, self.device)

    def _table(self, dv: DeviceVideos):
        """window table of a video batch, cached ON the batch object (never keyed by id(): ids are reused)."""
        key = (self.clip_len, self.stride)
        cache = dv.__dict__.setdefault("_window_tables", {})
        if key not in cache and dv.uniform_len is not None and dv.uniform_len >= self.clip_len and self.device.type == "cuda":
            # clips of one length on the regular grid: the table is arithmetic — built on the device, no host copies
            V, L, T, st = dv.n_videos, int(dv.uniform_len), self.clip_len, max(1, self.stride)
            wpv = (L - T) // st + 1
            idx = torch.arange(V * wpv, device=self.device, dtype=torch.int32)
            cache[key] = (torch.div(idx, wpv, rounding_mode="floor").to(torch.int32), (idx % wpv) * st,
                          torch.arange(V + 1, device=self.device, dtype=torch.int64) * wpv, V * wpv)
        if key not in cache:
            wv, ws, seg = window_table(dv.lengths, self.clip_len, self.stride)
            cache[key] = (torch.from_numpy(wv).to(self.device), torch.from_numpy(ws).to(self.device),
                          torch.from_numpy(seg).to(self.device), int(seg[-1]))
        return cache[key]

    def encode(self, dv: DeviceVideos, want_frames: bool = False, use_clips: Optional[bool] = None):
        """-> dict(seq [N,256], tc_window [N], seg [V+1], frames [N,T+1,256]|None, flags int32[1]).
        use_clips: None = tag_encode_clips whenever all videos have the same length (>= clip_len), False = always the
        explicit window table (tag_encode_windows); both give the same windows in the same order."""
        lib = _lib.load()
        wv, ws, seg, N = self._table(dv)
        T = self.clip_len
        h = self.model.handle(self.device, T)
        seq = torch.empty(N, 256, device=self.device, dtype=torch.float32)
        tcw = torch.empty(N, device=self.device, dtype=torch.float32)
        frames = torch.empty(N, T + 1, 256, device=self.device, dtype=torch.float32) if want_frames else None
        flags = torch.zeros(1, device=self.device, dtype=torch.int32)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        lens = dv.lengths
        L = int(lens[0]) if len(lens) else 0
        uniform = dv.uniform_len is not None and L >= T and use_clips is not False
        with torch.cuda.device(self.device):
            if uniform:
                # clips of equal length on the reference's regular window grid: the library can build the features once per
                # source frame and let the stem GEMMs gather the (overlapping) windows
                _lib.check(h, lib.tag_encode_clips(h, C.byref(dv.c), _lib.ptr(self.mean), _lib.ptr(self.std), len(lens), L, T,
                                                   self.stride, seq.data_ptr(), _lib.ptr(frames), None, tcw.data_ptr(),
                                                   flags.data_ptr(), stream), "tag_encode_clips")
            else:
                _lib.check(h, lib.tag_encode_windows(h, C.byref(dv.c), _lib.ptr(self.mean), _lib.ptr(self.std), wv.data_ptr(),
                                                     ws.data_ptr(), N, T, seq.data_ptr(), _lib.ptr(frames), None, tcw.data_ptr(),
                                                     flags.data_ptr(), stream), "tag_encode_windows")
        return {"seq": seq, "tc_window": tcw, "seg": seg, "frames": frames, "flags": flags, "win_video": wv}

    # ------------------------------------------------------------------ centroid build (config 3)
    def centroid_sums(self, dv: DeviceVideos, n_classes: int) -> torch.Tensor:
        """Local [C,257] sums||counts of this rank's videos (window label = its video's class)."""
        enc = self.encode(dv)
        vid_label = torch.tensor(dv.vb.cls_idx, device=self.device, dtype=torch.int32)
        y = vid_label.index_select(0, enc["win_video"].long()).contiguous()
        sc = torch.zeros(n_classes, 257, device=self.device, dtype=torch.float32)
        centroid_accumulate(enc["seq"], y, sc)
        return sc

    def build_centroids(self, dv: DeviceVideos, n_classes: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """build_real_centroids (eval.py:260-286) over this rank's shard + the one all-reduce."""
        sc = self.centroid_sums(dv, n_classes)
        allreduce_centroid_sums(sc, group)
        return centroid_finalize(sc)

    # ------------------------------------------------------------------ scoring (config 2)
    def score(self, dv: DeviceVideos, centroids: torch.Tensor, labels: Optional[torch.Tensor] = None):
        """-> (ac [V], tc [V]) device tensors; ac is NaN for videos whose class has no centroid."""
        lib = _lib.load()
        enc = self.encode(dv)
        V = dv.n_videos
        if labels is None:                                   # class index per video, cached on the batch object
            labels = dv.__dict__.get("_labels")
            if labels is None:
                labels = dv.__dict__["_labels"] = torch.tensor(dv.vb.cls_idx, device=self.device, dtype=torch.int32)
        ac = torch.empty(V, device=self.device, dtype=torch.float32)
        tc = torch.empty(V, device=self.device, dtype=torch.float32)
        h = util_handle(self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(h, lib.tag_score(h, enc["seq"].data_ptr(), enc["tc_window"].data_ptr(), enc["seg"].data_ptr(),
                                        labels.data_ptr(), centroids.data_ptr(), int(centroids.shape[0]), V, ac.data_ptr(),
                                        tc.data_ptr(), stream), "tag_score")
        self.last_flags = enc["flags"]
        return ac, tc

    def _block_plan(self, lengths: Sequence[int], first: bool) -> List[Tuple[int, int]]:
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        return block_plan(lengths, self.clip_len, self.stride, int(self.model.max_windows), first, sms)

    def score_stream(self, batches, centroids: torch.Tensor, pieces: Optional[int] = None, prefetch: int = 3):
        """Streaming end-to-end call: `batches` is an iterable of HOST (ideally pinned) VideoBatch objects; yields
        one `(ac [V], tc [V])` pair of CPU tensors per batch, in order. This is the `e2e` leg of bench.py.

        Every batch is cut into contiguous blocks of videos (`pieces` equal blocks, or — default — one block per
        encoder pass with a short ramp at the start of the stream, see `_block_plan`). Blocks are copied host->device
        on a side stream up to `prefetch` blocks ahead of the block being scored (across batch boundaries), and the
        8 B/video results go back through a pinned buffer one batch behind the compute, so in steady state the PCIe
        time (350 KB per video) hides behind the encoder. At most prefetch+2 blocks of inputs are alive on the device."""
        from collections import deque
        dev = self.device
        main = torch.cuda.current_stream(dev)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=dev)
        cs = self._copy_stream
        cs.wait_stream(main)

        def jobs():
            for b, vb in enumerate(batches):
                if vb.n_videos == 0:                          # an empty batch still gets its (empty) result, in order
                    yield vb, 0, 0, True, 0
                    continue
                if pieces is None:
                    plan = self._block_plan([vb.length(v) for v in range(vb.n_videos)], first=(b == 0))
                else:
                    n = max(1, min(int(pieces), vb.n_videos))
                    plan = [shard_range(vb.n_videos, i, n) for i in range(n)]
                cap = max((vb.offsets[hi] - vb.offsets[lo] for lo, hi in plan), default=0)     # frames of the largest block
                for i, (lo, hi) in enumerate(plan):
                    yield vb, lo, hi, i == len(plan) - 1, cap

        it = jobs()
        staged, done = deque(), []
        n_blocks = [0]
        # device staging ring: block j lives in slot j % ring_slots, and its copy waits for the block that used the slot
        # before (done[j - ring_slots]) — no allocator traffic and no cross-stream frees inside the stream
        ring_slots = prefetch + 2

        def stage():
            job = next(it, None)
            if job is None:
                return
            vb, lo, hi, last, cap = job
            if vb.n_videos == 0:
                staged.append((vb, 0, 0, True, None, None))
                return
            j = n_blocks[0]                                   # index of this block in the stream
            n_blocks[0] += 1
            if j - prefetch - 2 >= 0:
                cs.wait_event(done[j - prefetch - 2])         # bound the device copies that are alive
            with torch.cuda.stream(cs):
                # everything the block needs from the host goes over the COPY stream: a host->device copy queued on the
                # compute stream would sit behind the prefetched blocks in the copy engine's queue and stall the encoder
                hb = vb.slice(lo, hi)
                piece = self._stage_block(hb, j % ring_slots, cap)
                meta = vb.__dict__.get("_stream_meta")
                if meta is None:                              # pinned once per host batch: class index and frame offsets
                    meta = vb.__dict__["_stream_meta"] = (torch.tensor(vb.cls_idx, dtype=torch.int32).pin_memory(),
                                                          torch.tensor(vb.offsets, dtype=torch.int64).pin_memory())
                labels = meta[0][lo:hi].to(dev, non_blocking=True)
                offs = meta[1][lo:hi + 1].to(dev, non_blocking=True) - int(vb.offsets[lo])
                lens = [hb.length(v) for v in range(hb.n_videos)]
                table = None
                if not (all(x == lens[0] for x in lens) and lens[0] >= self.clip_len):
                    # ragged block: its explicit window table travels with it (equal-length blocks build theirs on the device)
                    wv, ws, seg = window_table(lens, self.clip_len, self.stride)
                    table = tuple(torch.from_numpy(x).pin_memory().to(dev, non_blocking=True) for x in (wv, ws, seg)) + (int(seg[-1]),)
                ev = torch.cuda.Event()
                ev.record(cs)
            staged.append((vb, lo, hi, last, (piece, labels, offs, table), ev))

        def finish(p):
            ev, host, flags = p
            ev.synchronize()
            self.last_flags = flags
            return host[0], host[1]

        for _ in range(prefetch + 1):
            stage()
        pending, out, flags = None, None, None
        while staged:
            vb, lo, hi, last, piece, ev = staged.popleft()
            if piece is None:                                 # empty batch
                stage()
                e = torch.cuda.Event()
                e.record(main)
                if pending is not None:
                    yield finish(pending)
                pending = (e, torch.empty(2, 0, dtype=torch.float32), None)
                continue
            main.wait_event(ev)
            if lo == 0:
                out = torch.empty(2, vb.n_videos, device=dev, dtype=torch.float32)
                flags = None
            piece, labels, offs, table = piece
            dv = DeviceVideos(piece, self.model.modalities, dev, frame_offset=offs)
            for t in (labels, offs) + (table[:3] if table is not None else ()):
                t.record_stream(main)                         # allocated on the copy stream, read by kernels of this one
            if table is not None:
                dv.__dict__.setdefault("_window_tables", {})[(self.clip_len, self.stride)] = table
            a, t_ = self.score(dv, centroids, labels=labels)
            out[0, lo:hi].copy_(a)
            out[1, lo:hi].copy_(t_)
            flags = self.last_flags if flags is None else flags + self.last_flags
            d = torch.cuda.Event()
            d.record(main)
            done.append(d)
            stage()                                           # keep the copy window full
            if last:
                host = torch.empty(2, vb.n_videos, dtype=torch.float32, pin_memory=True)
                host.copy_(out, non_blocking=True)
                e = torch.cuda.Event()
                e.record(main)
                if pending is not None:
                    yield finish(pending)                     # one batch behind: the GPU never waits for the host
                pending = (e, host, flags)
        if pending is not None:
            yield finish(pending)

    def _stage_block(self, hb: VideoBatch, slot: int, cap_frames: int) -> VideoBatch:
        """Copy a host block into staging slot `slot` (device buffers of at least `cap_frames` frames, re-allocated only
        when a larger batch plan arrives) on the current stream; returns a VideoBatch of views."""
        ring = self.__dict__.setdefault("_ring", {})
        bufs = ring.setdefault(slot, {})
        out = {}
        for name in ("pose", "gori", "betas", "vit", "kp", "clip", "dino"):
            src = getattr(hb, name)
            if src is None:
                out[name] = None
                continue
            buf = bufs.get(name)
            if buf is None or buf.shape[0] < src.shape[0] or buf.shape[1:] != src.shape[1:]:
                buf = torch.empty((max(int(cap_frames), src.shape[0]),) + tuple(src.shape[1:]), device=self.device, dtype=src.dtype)
                bufs[name] = buf
            out[name] = buf[:src.shape[0]]
            out[name].copy_(src, non_blocking=True)
        return VideoBatch(out["pose"], out["gori"], out["betas"], out["vit"], out["kp"], list(hb.offsets), list(hb.cls_idx),
                          list(hb.names), out["clip"], out["dino"], list(hb.classes))

    def score_host(self, vb_host: VideoBatch, centroids: torch.Tensor, pieces: int = 4) -> Tuple[torch.Tensor, torch.Tensor]:
        """End-to-end call with HOST buffers for one batch: H2D of every input array, score, D2H of the per-video
        results (`score_stream` over a single batch). Returns CPU tensors (ac [V], tc [V])."""
        return next(iter(self.score_stream([vb_host], centroids, pieces=pieces, prefetch=pieces)))

    def score_files(self, ingest, centroids: torch.Tensor, items=None, videos_per_batch: int = 5000) -> Dict[str, Dict[str, float]]:
        """eval.py:394-451 from FILES: `ingest` (ingest.NpzIngest over the reference's on-disk layout) parses `.npz` /
        `keypoints.npy` into pinned host batches on background threads, `score_stream` moves them through the device staging
        ring and scores them; returns {video_id: {"ac","tc"}} (what eval.py writes to video_scores.json)."""
        out: Dict[str, Dict[str, float]] = {}
        held = []

        def feed():
            for vb in ingest.batches(items, videos_per_batch):
                held.append(vb)
                yield vb

        for i, (ac, tc) in enumerate(self.score_stream(feed(), centroids)):
            out.update(self.scores_dict(held[i], ac, tc))
            held[i] = None
        return out

    def scores_dict(self, vb: VideoBatch, ac: torch.Tensor, tc: torch.Tensor) -> Dict[str, Dict[str, float]]:
        """{video_id: {"ac","tc"}} as eval.py:439-447 (a key is absent when the reference would skip it)."""
        ac, tc = ac.cpu().tolist(), tc.cpu().tolist()
        out = {}
        for v, name in enumerate(vb.names):
            e = {}
            if ac[v] == ac[v]:
                e["ac"] = float(ac[v])
            if tc[v] == tc[v]:
                e["tc"] = float(tc[v])
            out[os.path.splitext(name)[0]] = e
        return out
