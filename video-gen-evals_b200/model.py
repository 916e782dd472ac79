"""Drop-in for the reference `model.py`: `HumanActionScorer` with the same constructor, the same
state_dict keys (so reference checkpoints load, eval.py:136-165) and the same forward contract
`model(x[B,T,D]) -> (seq_embed[B,256], frame_embeds[B,T+1,256], tokens[B,T+1,256])`
(reference model.py:102-193) — but the forward runs entirely in libtag_b200.so on sm_100a.

The nn.Module tree below only HOLDS parameters (names/shapes identical to the reference); there is
no PyTorch compute path and no fallback: forward raises if the CUDA library is unavailable.
"""
from __future__ import annotations

import ctypes as C
import math
import typing as T

import torch
import torch.nn as nn

from . import _lib
from .synth import sinusoidal_pe


def _uniform_(t: torch.Tensor, fan_in: int) -> torch.Tensor:
    b = 1.0 / math.sqrt(max(1, fan_in))
    return t.uniform_(-b, b)


class _Weight(nn.Module):
    def __init__(self, shape, bias: bool = False):
        super().__init__()
        fan_in = int(torch.tensor(shape[1:]).prod().item()) if len(shape) > 1 else shape[0]
        self.weight = nn.Parameter(_uniform_(torch.empty(*shape), fan_in))
        if bias:
            self.bias = nn.Parameter(_uniform_(torch.empty(shape[0]), fan_in))


class _Affine(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.bias = nn.Parameter(torch.zeros(d))


class _ConvBlock(nn.Module):                       # reference TemporalConvBlock, model.py:21-40
    def __init__(self, c, k):
        super().__init__()
        self.conv1 = _Weight((c, c, k))
        self.conv2 = _Weight((c, c, k))
        self.norm = _Affine(c)


class _ConvEncoder(nn.Module):                     # reference MovementConvEncoder, model.py:43-58
    def __init__(self, d_in, d_out, k=5, n_blocks=4):
        super().__init__()
        self.stem = _Weight((d_out, d_in, 1))
        self.blocks = nn.ModuleList([_ConvBlock(d_out, k) for _ in range(n_blocks)])
        self.proj = _Weight((d_out, d_out))


class _Fusion(nn.Module):                          # reference MinimalPerFrameFusion, model.py:61-98
    def __init__(self, d, m):
        super().__init__()
        self.latent = nn.Parameter(torch.randn(1, 1, d))
        self.q_ln = _Affine(d)
        self.kv_ln = _Affine(d)
        self.Wq = _Weight((d, d))
        self.Wk = _Weight((d, d))
        self.Wv = _Weight((d, d))
        self.Wo = _Weight((d, d))
        self.logit_temp = nn.Parameter(torch.zeros(m))
        self.logit_bias = nn.Parameter(torch.zeros(m))


class _SelfAttn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.in_proj_weight = nn.Parameter(_uniform_(torch.empty(3 * d, d), d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = _Weight((d, d), bias=True)


class _TransformerLayer(nn.Module):                # nn.TransformerEncoderLayer parameter layout, model.py:145
    def __init__(self, d, ffn):
        super().__init__()
        self.self_attn = _SelfAttn(d)
        self.linear1 = _Weight((ffn, d), bias=True)
        self.linear2 = _Weight((d, ffn), bias=True)
        self.norm1 = _Affine(d)
        self.norm2 = _Affine(d)


class _Temporal(nn.Module):
    def __init__(self, d, ffn, n):
        super().__init__()
        self.layers = nn.ModuleList([_TransformerLayer(d, ffn) for _ in range(n)])


class _PosEnc(nn.Module):                          # reference SinusoidalPositionalEmbedding, model.py:8-19
    def __init__(self, d, max_len=5000):
        super().__init__()
        self.register_buffer("pe", sinusoidal_pe(max_len, d))


class HumanActionScorer(nn.Module):
    """Same signature as reference model.py:103-110, plus keyword-only runtime knobs:
    precision 'fp16_tc' (tcgen05 tensor cores, default) or 'fp32' (CUDA cores, reference precision);
    max_windows = windows per internal pass (workspace size)."""

    def __init__(self,
                 dims_map_raw: T.Dict[str, int],
                 dims_map_diff: T.Dict[str, int],
                 d_model: int = 256,
                 latent_dim: int = 128,
                 time_layers: int = 4,
                 time_heads: int = 8,
                 dropout: float = 0.1,
                 *, precision: str = "fp16_tc", max_windows: int = 1024):
        super().__init__()
        if not isinstance(dims_map_raw, dict) or not isinstance(dims_map_diff, dict):
            raise ValueError("dims_map_raw and dims_map_diff must be dicts of {modality_name: dim}.")
        if set(dims_map_raw.keys()) != set(dims_map_diff.keys()):
            raise ValueError("dims_map_raw and dims_map_diff must have the same modality keys.")
        if precision not in ("fp16_tc", "fp32"):
            raise ValueError("precision must be 'fp16_tc' or 'fp32'")
        self.modalities = list(dims_map_raw.keys())
        if len(self.modalities) > _lib.TAG_MAX_MODALITIES:
            raise ValueError(f"at most {_lib.TAG_MAX_MODALITIES} modalities are supported")
        self.dim_map_raw = {m: int(dims_map_raw[m]) for m in self.modalities}
        self.dim_map_diff = {m: int(dims_map_diff[m]) for m in self.modalities}
        self.one_pass_raw = sum(self.dim_map_raw.values())
        self.one_pass_diff = sum(self.dim_map_diff.values())
        self.has_diff = any(v > 0 for v in self.dim_map_diff.values())
        self.M = len(self.modalities)
        self.d_model, self.time_layers, self.time_heads = d_model, time_layers, time_heads
        self.precision, self.max_windows = precision, int(max_windows)

        self.state_enc = nn.ModuleDict({m: _ConvEncoder(self.dim_map_raw[m], d_model) for m in self.modalities})
        if self.has_diff:
            self.motion_enc = nn.ModuleDict({m: _ConvEncoder(self.dim_map_diff[m], d_model)
                                             for m in self.modalities if self.dim_map_diff[m] > 0})
        else:
            self.motion_enc = None
        self.fusion = _Fusion(d_model, self.M)
        self.cls = nn.Parameter(torch.randn(1, 1, d_model))
        self.pos_enc = _PosEnc(d_model)
        self.temporal = _Temporal(d_model, 4 * d_model, time_layers)
        self.last_attn = None

        self._h = None            # tag_handle*
        self._h_key = None        # (device index, max_T, precision, max_windows)
        self._w_sig = None        # parameter versions the handle was packed from

    # ------------------------------------------------------------------ handle management
    @property
    def feat_dim(self) -> int:
        return self.one_pass_raw + self.one_pass_diff

    def _kinds(self):
        try:
            return [_lib.KIND_OF[m] for m in self.modalities]
        except KeyError as e:
            raise ValueError(f"unknown modality {e}; known: {sorted(_lib.KIND_OF)}") from None

    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _release(self):
        if self._h is not None:
            _lib.load().tag_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def handle(self, device: torch.device, T_frames: int):
        """Create (or re-create) the native handle for this device / clip length and pack weights."""
        lib = _lib.load()
        if device.type != "cuda":
            raise _lib.TagError("HumanActionScorer runs only on a CUDA (sm_100a) device; there is no CPU path")
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        sig = self._signature()
        same_shape = self._h is not None and self._h_key[0] == dev_index and self._h_key[1] >= T_frames and \
            self._h_key[2:] == (self.precision, self.max_windows)
        if same_shape and self._w_sig == sig:
            return self._h
        if same_shape:
            # parameter values changed (e.g. an optimiser step): re-pack the weights into the existing handle, keep its workspace
            h, max_T = self._h, self._h_key[1]
            _lib.check(h, lib.tag_reload_weights_begin(h), "tag_reload_weights_begin")
            try:
                self._load_all(lib, h, max_T)
            except Exception:
                self._release()
                raise
            self._w_sig = sig
            return h
        self._release()
        max_T = max(32, int(T_frames))
        cfg = _lib.tag_config()
        cfg.n_modalities = self.M
        kinds = self._kinds()
        for i, m in enumerate(self.modalities):
            cfg.raw_dims[i] = self.dim_map_raw[m]
            cfg.diff_dims[i] = self.dim_map_diff[m]
            cfg.kinds[i] = kinds[i]
        cfg.d_model, cfg.n_heads, cfg.n_layers, cfg.ffn_dim = self.d_model, self.time_heads, self.time_layers, 4 * self.d_model
        cfg.n_blocks, cfg.conv_kernel = 4, 5
        cfg.precision = _lib.PRECISION_FP16_TC if self.precision == "fp16_tc" else _lib.PRECISION_FP32
        cfg.max_windows, cfg.max_T, cfg.device = self.max_windows, max_T, dev_index
        h = C.c_void_p()
        _lib.check(None, lib.tag_create(C.byref(h), C.byref(cfg)), "tag_create")
        try:
            self._load_all(lib, h, max_T)
        except Exception:
            lib.tag_destroy(h)
            raise
        self._h, self._h_key, self._w_sig = h, (dev_index, max_T, self.precision, self.max_windows), sig
        return h

    def _load_all(self, lib, h, max_T: int):
        """every state_dict tensor -> tag_load_weight (modality names become indices in the keys), then finalize"""
        idx = {m: str(i) for i, m in enumerate(self.modalities)}
        for key, t in self.state_dict().items():
            parts = key.split(".")
            if parts[0] in ("state_enc", "motion_enc"):
                parts[1] = idx[parts[1]]
            if key == "pos_enc.pe":
                t = t[:, :max_T + 1, :]
            t = t.detach().to(torch.float32).contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(h, lib.tag_load_weight(h, ".".join(parts).encode(), t.data_ptr(), shape, t.dim()),
                       f"tag_load_weight({key})")
        _lib.check(h, lib.tag_finalize_weights(h), "tag_finalize_weights")

    # ------------------------------------------------------------------ reference forward contract
    def forward(self, x: torch.Tensor, modality_mask=None):
        if self.training:
            raise _lib.TagError("libtag_b200 implements the inference (eval / no-grad) forward only; call model.eval()")
        if x.dim() != 3 or x.shape[-1] != self.feat_dim:
            raise ValueError(f"expected x of shape [B, T, {self.feat_dim}], got {tuple(x.shape)}")
        B, Tn, _ = x.shape
        h = self.handle(x.device, Tn)
        lib = _lib.load()
        x = x.detach().to(torch.float32).contiguous()
        seq = torch.empty(B, self.d_model, device=x.device, dtype=torch.float32)
        frames = torch.empty(B, Tn + 1, self.d_model, device=x.device, dtype=torch.float32)
        tokens = torch.empty(B, Tn + 1, self.d_model, device=x.device, dtype=torch.float32)
        attn = torch.empty(B * Tn, self.M, device=x.device, dtype=torch.float32)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            _lib.check(h, lib.tag_set_fusion_attn_out(h, attn.data_ptr()), "tag_set_fusion_attn_out")
            _lib.check(h, lib.tag_encode(h, x.data_ptr(), B, Tn, seq.data_ptr(), frames.data_ptr(), tokens.data_ptr(),
                                         None, stream), "tag_encode")
        self.last_attn = attn      # [B*T, M] fusion softmax, as the reference keeps it (model.py:94, :185)
        return seq, frames, tokens

    def launch_count(self) -> int:
        return 0 if self._h is None else int(_lib.load().tag_launch_count(self._h))
