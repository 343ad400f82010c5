"""ctypes binding of libggad_b200.so (include/ggad_b200.h).

The library is the product: if it is missing or cannot be loaded this module raises --
there is no CPU / eager fallback anywhere in ggad_b200.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GGAD_B200_LIB") or os.path.join(_HERE, "libggad_b200.so")

GGAD_TILE_ITEMS = 2048
GGAD_MAX_WIDTH = 768

_vp = C.c_void_p


class GatherDesc(C.Structure):
    """Mirror of ggad_gather_desc_t."""
    _fields_ = [
        ("rowptr", _vp), ("col", _vp), ("val", _vp), ("n_rows", C.c_int64), ("nnz", C.c_int64),
        ("x", _vp), ("ldx", C.c_int64), ("xmap", _vp), ("col_scale", _vp), ("row_scale", _vp),
        ("d", C.c_int32), ("relu", C.c_int32),
        ("bias", _vp), ("prelu_slope", _vp), ("y", _vp), ("z", _vp), ("ldy", C.c_int64),
        ("sumsq", _vp), ("dot_mat", _vp), ("lddot", C.c_int64), ("dot_rows", _vp), ("dot_scale", _vp),
        ("dot_out", _vp),
        ("tile_row", _vp), ("tile_edge", _vp), ("n_tiles", C.c_int64), ("ws", _vp),
        ("y_peer", _vp * 7), ("n_peer", C.c_int32), ("tile_epoch", C.c_int32), ("y_multicast", _vp), ("peer_need", _vp),
        ("tile_done", _vp), ("mc_min_peers", C.c_int32), ("n_x_rows", C.c_int32),
    ]


class ChaseDesc(C.Structure):
    """Mirror of ggad_chase_desc_t."""
    _fields_ = [
        ("y", _vp), ("ldy", C.c_int64), ("d", C.c_int32), ("n_peer", C.c_int32), ("rowptr", _vp), ("n_rows", C.c_int64),
        ("tile_row", _vp), ("tile_edge", _vp), ("n_tiles", C.c_int64), ("tile_done", _vp), ("tile_epoch", C.c_int32),
        ("n_ctas", C.c_int32), ("peer_need", _vp), ("y_peer", _vp * 7), ("y_multicast", _vp),
        ("mc_min_peers", C.c_int32), ("reserved", C.c_int32),
    ]


class TailDesc(C.Structure):
    """Mirror of ggad_tail_desc_t."""
    _fields_ = [
        ("combined", _vp), ("ld_combined", C.c_int64), ("ego", _vp), ("ld_ego", C.c_int64), ("fc", _vp), ("weight", _vp),
        ("labels", _vp), ("batch", C.c_int32), ("h", C.c_int32), ("rows", _vp), ("apre_src", _vp), ("apre_own", _vp),
        ("scores", _vp), ("bce", _vp), ("cos", _vp), ("dist", _vp), ("norms", _vp), ("src", _vp), ("out", _vp),
        ("grad_total", _vp), ("d_combined", _vp), ("ld_d_combined", C.c_int64), ("d_apre", _vp), ("d_ego", _vp),
        ("d_scores", _vp), ("d_weight", _vp),
    ]


class ResidentCSR(C.Structure):
    """Mirror of ggad_resident_csr_t."""
    _fields_ = [
        ("rowptr", _vp), ("col", _vp), ("val", _vp), ("row_scale", _vp), ("col_scale", _vp),
        ("n_rows", C.c_int64), ("n_cols", C.c_int64), ("nnz", C.c_int64),
        ("tile_row", _vp), ("tile_edge", _vp), ("n_tiles", C.c_int64),
    ]


# name -> (restype, argtypes); must list every symbol include/ggad_b200.h declares
_i64, _i32, _f = C.c_int64, C.c_int32, C.c_float
SIGNATURES = {
    "ggad_version": (C.c_int, []),
    "ggad_last_error": (C.c_char_p, []),
    "ggad_launch_count": (_i64, []),
    "ggad_trim_workspace": (C.c_int, []),
    "ggad_reserve_workspace": (C.c_int, [_i64, _vp]),
    "ggad_device_info": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "ggad_gather_reduce": (C.c_int, [C.POINTER(GatherDesc), _vp]),
    "ggad_halo_push": (C.c_int, [_vp, _i64, _i64, _i32, _vp, _vp, _i32, _vp]),
    "ggad_halo_chase": (C.c_int, [C.POINTER(ChaseDesc), _vp]),
    "ggad_dense_matmul": (C.c_int, [_i32, _i32, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _f, _f, _i32, _i32, _vp]),
    "ggad_plan_num_tiles": (_i64, [_i64, _i64]),
    "ggad_plan_build": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "ggad_plan_build_padded": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "ggad_normalize_backward": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _i64, _i32, _vp]),
    "ggad_row_inv_norm": (C.c_int, [_vp, _i64, _i64, _i32, _vp, _vp, _vp]),
    "ggad_coo_keys_to_csr": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "ggad_csr_transpose": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "ggad_csr_extract_rows": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "ggad_csr_row_sum_f64": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "ggad_csr_add_identity_rowptr": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "ggad_csr_scale_add_identity": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "ggad_col_histogram": (C.c_int, [_vp, _i64, _vp, _i64, _vp]),
    "ggad_block_rowptr": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp]),
    "ggad_block_fill": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp]),
    "ggad_unique_sorted": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "ggad_block_remap": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp]),
    "ggad_block_col_weights": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "ggad_minibatch_tail_fwd": (C.c_int, [C.POINTER(TailDesc), _vp]),
    "ggad_minibatch_tail_bwd": (C.c_int, [C.POINTER(TailDesc), _vp]),
    "ggad_rmat_keys": (C.c_int, [_vp, _i64, _i64, _i32, _i32, C.c_uint64, _f, _f, _f, _i64, _i64, _vp, _vp]),
    "ggad_spmm_fwd_bwd_host": (C.c_int, [C.POINTER(ResidentCSR), C.POINTER(ResidentCSR), _vp, _vp, _vp, _vp, _i32,
                                         _vp, _vp, _vp, _vp, _vp]),
    "ggad_spmm_fwd_bwd_host_enqueue": (C.c_int, [C.POINTER(ResidentCSR), C.POINTER(ResidentCSR), _vp, _vp, _vp, _vp, _i32,
                                                 _vp, _vp, _vp, _vp, _vp]),
}

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load (once) and return the C-ABI library; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"ggad_b200: CUDA library {LIB_PATH} is not built; run `python -m ggad_b200.build` "
                "(there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().ggad_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"ggad_b200 error {rc}: {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"ggad_b200: {what} must be a CUDA tensor (no CPU fallback); got {t.device}")


def launch_count() -> int:
    return int(lib().ggad_launch_count())
