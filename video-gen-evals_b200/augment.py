"""Hard-negative window augmentations of the reference (utils.py:65-95) as ONE device kernel: every one of them is a
per-sample permutation / selection of frames, i.e. a gather `out[b, t] = x[b, idx[b, t]]` (tag_gather_frames).

  partial_shuffle_within_window(seqs, shuffle_fraction=0.7)    utils.py:65-75
  reverse_sequence(seqs)                                       utils.py:78-86
  get_static_window(seqs)                                      utils.py:88-95

`partial_shuffle_within_window` draws its permutations exactly as the reference does (two `torch.randperm` calls per
sample on the CPU generator, in batch order), so under the same `torch.manual_seed` it returns the same windows.
"""
from __future__ import annotations

import torch

from . import _lib


def gather_frames(seqs: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """out[b, t, :] = seqs[b, idx[b, t], :] on the GPU. seqs [B,T,D] fp32 (cuda), idx [B,T] integer."""
    from .scoring import util_handle
    if not seqs.is_cuda:
        raise _lib.TagError("augmentations run on a CUDA (sm_100a) device; there is no CPU path")
    lib = _lib.load()
    B, T, D = seqs.shape
    x = seqs.detach().to(torch.float32).contiguous()
    pad = (-D) % 4
    if pad:                                         # the kernel moves 16-byte pieces
        x = torch.nn.functional.pad(x, (0, pad))
    ix = idx.to(seqs.device, dtype=torch.int32).contiguous()
    out = torch.empty_like(x)
    h = util_handle(seqs.device)
    stream = torch.cuda.current_stream(seqs.device).cuda_stream
    with torch.cuda.device(seqs.device):
        _lib.check(h, lib.tag_gather_frames(h, x.data_ptr(), ix.data_ptr(), B, T, D + pad, out.data_ptr(), stream),
                   "tag_gather_frames")
    return out[..., :D] if pad else out


def shuffle_indices(batch_size: int, length: int, shuffle_fraction: float = 0.7) -> torch.Tensor:
    """The frame permutation utils.py:65-75 applies to every sample, with the reference's RNG consumption."""
    idx = torch.arange(length).repeat(batch_size, 1)
    if length > 1:
        n = max(1, int(shuffle_fraction * length))
        for i in range(batch_size):
            sel = torch.randperm(length)[:n]
            idx[i, sel] = sel[torch.randperm(n)]
    return idx


def partial_shuffle_within_window(seqs: torch.Tensor, shuffle_fraction: float = 0.7) -> torch.Tensor:
    B, T, _ = seqs.shape
    return gather_frames(seqs, shuffle_indices(B, T, shuffle_fraction))


def reverse_sequence(seqs: torch.Tensor) -> torch.Tensor:
    B, T, _ = seqs.shape
    return gather_frames(seqs, torch.arange(T - 1, -1, -1, device=seqs.device).repeat(B, 1))


def get_static_window(seqs: torch.Tensor) -> torch.Tensor:
    B, T, _ = seqs.shape
    return gather_frames(seqs, torch.zeros(B, T, dtype=torch.int32, device=seqs.device))
