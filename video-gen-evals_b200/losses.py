"""Forward passes of the reference's training losses (losses.py) on the GPU — BASELINE config 5, SURVEY.md §8f row N1.

  TCL(temperature, k1, k2)(projections, targets)                losses.py:6-34
  SupConWithHardNegatives(temperature)(anchor, positive, hard)  losses.py:37-56
  hard_negative_step(model, feats, labels)                      train.py:511-524 (`compute_loss_components`, forward only)

Backward / optimiser are out of scope (the north star is inference scoring); these are the embedding pass and the loss
values of one training step. `Z Z^T` of TCL runs as a tcgen05 GEMM with the masked row sums in its epilogue (B >= 512).
"""
from __future__ import annotations

from typing import Dict

import torch

from . import _lib
from .augment import get_static_window, partial_shuffle_within_window, reverse_sequence


def _dev(*tensors) -> torch.device:
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    if not torch.cuda.is_available():
        raise _lib.TagError("the TAG loss kernels need a CUDA (sm_100a) device; there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


class TCL(torch.nn.Module):
    """Forward of reference losses.py:6-34 in one pass over the similarity matrix (never materialised)."""

    def __init__(self, temperature=0.1, k1=5000.0, k2=1.0):
        super().__init__()
        self.temperature, self.k1, self.k2 = float(temperature), float(k1), float(k2)

    def loss_rows(self, projections: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        from .scoring import util_handle
        lib = _lib.load()
        dev = _dev(projections)
        h = util_handle(dev)
        z = projections.detach().to(dev, dtype=torch.float32).contiguous()
        if z.dim() != 2 or z.shape[1] != 256:
            raise ValueError("projections must be [B, 256]")
        y = targets.to(dev, dtype=torch.int32).contiguous()
        out = torch.empty(z.shape[0], device=dev, dtype=torch.float32)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _lib.check(h, lib.tag_tcl_forward(h, z.data_ptr(), y.data_ptr(), z.shape[0], self.temperature, self.k1, self.k2,
                                              out.data_ptr(), stream), "tag_tcl_forward")
        return out

    def forward(self, projections, targets):
        return self.loss_rows(projections, targets).mean()


class SupConWithHardNegatives(torch.nn.Module):
    """Forward of reference losses.py:37-56: 2-way cross entropy (positive vs hard negative) per anchor, mean over the batch."""

    def __init__(self, temperature=0.07):
        super().__init__()
        self.temperature = float(temperature)

    def loss_rows(self, anchor, positive, hard_negative) -> torch.Tensor:
        from .scoring import util_handle
        lib = _lib.load()
        dev = _dev(anchor, positive, hard_negative)
        h = util_handle(dev)
        a, p, n = (t.detach().to(dev, dtype=torch.float32).contiguous() for t in (anchor, positive, hard_negative))
        if a.dim() != 2 or a.shape[1] != 256 or p.shape != a.shape or n.shape != a.shape:
            raise ValueError("anchor, positive and hard_negative must all be [B, 256]")
        out = torch.empty(a.shape[0], device=dev, dtype=torch.float32)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _lib.check(h, lib.tag_supcon_hard_forward(h, a.data_ptr(), p.data_ptr(), n.data_ptr(), a.shape[0], self.temperature,
                                                      out.data_ptr(), stream), "tag_supcon_hard_forward")
        return out

    def forward(self, anchor, positive, hard_negative):
        return self.loss_rows(anchor, positive, hard_negative).mean()


@torch.no_grad()
def hard_negative_step(model, feats: torch.Tensor, labels: torch.Tensor, tcl: TCL = None, hard: SupConWithHardNegatives = None,
                       hard_weight: float = 10.0, shuffle_fraction: float = 0.7) -> Dict[str, torch.Tensor]:
    """The forward half of one training step of the reference (train.py:488-524): the embedding pass on the batch plus the
    three hard-negative passes on shuffled / reversed / static copies of it (utils.py:65-95), then
    {"tcl", "hard_shuf", "hard_rev", "hard_stat"} exactly as `compute_loss_components` (hard weight 10, train.py:24-27).
    `feats` [B,T,D] on the model's device; the model runs in eval mode (dropout is the only train-time difference)."""
    tcl = tcl or TCL()
    hard = hard or SupConWithHardNegatives()
    emb, _, _ = model(feats)
    sh_emb, _, _ = model(partial_shuffle_within_window(feats, shuffle_fraction))
    rev_emb, _, _ = model(reverse_sequence(feats))
    st_emb, _, _ = model(get_static_window(feats))
    return {"tcl": tcl(emb, labels), "hard_shuf": hard_weight * hard(emb, emb, sh_emb),
            "hard_rev": hard_weight * hard(emb, emb, rev_emb), "hard_stat": hard_weight * hard(emb, emb, st_emb)}
