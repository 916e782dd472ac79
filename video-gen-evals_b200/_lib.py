"""ctypes binding of libtag_b200.so (include/tag_b200.h). There is NO fallback: if the library is
missing or cannot be loaded, every product entry point raises."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtag_b200.so")

TAG_MAX_MODALITIES = 8
TAG_OK = 0
KIND_COSINE, KIND_ROTMAT, KIND_PLAIN, KIND_PROCRUSTES = 0, 1, 2, 3
PRECISION_FP32, PRECISION_FP16_TC = 0, 1

# modality name -> delta kind (reference utils.py:455-470)
KIND_OF = {"vit": KIND_COSINE, "clip": KIND_COSINE, "dino": KIND_COSINE, "global": KIND_ROTMAT,
           "pose": KIND_ROTMAT, "beta": KIND_PLAIN, "kp2d": KIND_PROCRUSTES}


class TagError(RuntimeError):
    pass


class tag_config(C.Structure):
    _fields_ = [("n_modalities", C.c_int32),
                ("raw_dims", C.c_int32 * TAG_MAX_MODALITIES),
                ("diff_dims", C.c_int32 * TAG_MAX_MODALITIES),
                ("kinds", C.c_int32 * TAG_MAX_MODALITIES),
                ("d_model", C.c_int32), ("n_heads", C.c_int32), ("n_layers", C.c_int32), ("ffn_dim", C.c_int32),
                ("n_blocks", C.c_int32), ("conv_kernel", C.c_int32), ("precision", C.c_int32),
                ("max_windows", C.c_int32), ("max_T", C.c_int32), ("device", C.c_int32)]


class tag_videos(C.Structure):
    _fields_ = [("src", C.c_void_p * TAG_MAX_MODALITIES),
                ("frame_offset", C.c_void_p),
                ("n_videos", C.c_int64)]


# every symbol include/tag_b200.h declares: name -> (restype, argtypes)
_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SIGNATURES = {
    "tag_create": (C.c_int, [C.POINTER(_P), C.POINTER(tag_config)]),
    "tag_destroy": (None, [_P]),
    "tag_last_error": (C.c_char_p, [_P]),
    "tag_abi_version": (C.c_int, []),
    "tag_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(_I64), _I32]),
    "tag_finalize_weights": (C.c_int, [_P]),
    "tag_reload_weights_begin": (C.c_int, [_P]),
    "tag_feature_fuse": (C.c_int, [_P, C.POINTER(tag_videos), _P, _P, _P, _P, _I64, _I32, _P, _P, _P]),
    "tag_encode": (C.c_int, [_P, _P, _I64, _I32, _P, _P, _P, _P, _P]),
    "tag_set_fusion_attn_out": (C.c_int, [_P, _P]),
    "tag_encode_windows": (C.c_int, [_P, C.POINTER(tag_videos), _P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P, _P, _P]),
    "tag_centroid_accumulate": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _P]),
    "tag_centroid_finalize": (C.c_int, [_P, _P, _I32, _P, _P, _P]),
    "tag_score": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I64, _P, _P, _P]),
    "tag_window_tc": (C.c_int, [_P, _P, _I64, _I32, _P, _P]),
    "tag_stats_accumulate": (C.c_int, [_P, _P, _I64, _I32, _P, _P, _P]),
    "tag_tcl_forward": (C.c_int, [_P, _P, _P, _I64, _F, _F, _F, _P, _P]),
    "tag_supcon_hard_forward": (C.c_int, [_P, _P, _P, _P, _I64, _F, _P, _P]),
    "tag_gather_frames": (C.c_int, [_P, _P, _P, _I64, _I32, _I32, _P, _P]),
    "tag_launch_count": (_I64, [_P]),
    "tag_set_profiling": (C.c_int, [_P, _I32]),
    "tag_get_profile": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "tag_get_profile_kinds": (C.c_int, [_P, C.POINTER(C.c_double), _I32]),
    "tag_debug_gemm_f32": (C.c_int, [_P, _P, _I32, _P, _I32, _I64, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _I32, _P]),
    "tag_encode_clips": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "tag_debug_feature_fuse16": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P]),
    "tag_debug_poison_workspace": (C.c_int, [_P, _P]),
    "tag_debug_tlayer_tail": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "tag_debug_tcn_block": (C.c_int, [_P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P]),
    "tag_debug_tcn_block_plan": (C.c_int, [_I64, _I32, _I32, _P, _P, _P]),
    "tag_debug_gemm_tc": (C.c_int, [_P, _P, _I32, _P, _I64, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _P]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libtag_b200.so (built in-tree by build.py / __graft_entry__.build). Raises TagError when
    it is missing — the product has no CPU or PyTorch fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TagError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). There is no CPU fallback for the TAG scoring path.")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:
        raise TagError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(h, rc: int, what: str):
    if rc != TAG_OK:
        msg = load().tag_last_error(h)
        raise TagError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t) -> Optional[int]:
    """device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
