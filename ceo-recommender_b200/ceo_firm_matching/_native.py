"""ctypes binding of libcfm_b200.so (the C ABI declared in include/cfm_b200.h).

There is NO fallback: if the shared library is missing, or a CUDA device is not
present when an op is called, the call raises.  Build the library with
``python __graft_entry__.py`` (or ``make -C ceo-recommender_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

CFM_MAX_TABLES = 16
CFM_MAX_PEERS = 8
CFM_TOPK_CAP = 384
CFM_ABI_VERSION = 3

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libcfm_b200.so")

p_f32 = C.c_void_p
i64 = C.c_int64


class Tower(C.Structure):
    """Mirror of ``cfm_tower_t``."""
    _fields_ = [
        ("n_num", i64), ("n_tables", i64), ("emb_dim", i64), ("h1", i64), ("h2", i64), ("d_out", i64),
        ("bn2", i64), ("drop1", C.c_double), ("drop2", C.c_double), ("tower_id", i64), ("precision", i64),
        ("x_num", C.c_void_p), ("x_cat", C.c_void_p),
        ("tables", C.c_void_p * CFM_MAX_TABLES), ("table_rows", i64 * CFM_MAX_TABLES),
        ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("w3", C.c_void_p), ("b3", C.c_void_p),
        ("bn1_w", C.c_void_p), ("bn1_b", C.c_void_p), ("bn2_w", C.c_void_p), ("bn2_b", C.c_void_p),
        ("bn1_rm", C.c_void_p), ("bn1_rv", C.c_void_p), ("bn2_rm", C.c_void_p), ("bn2_rv", C.c_void_p),
        ("bn1_nbt", C.c_void_p), ("bn2_nbt", C.c_void_p),
        ("h1_raw", C.c_void_p), ("h2_raw", C.c_void_p), ("out", C.c_void_p),
        ("bn1_stat", C.c_void_p), ("bn2_stat", C.c_void_p), ("scratch", C.c_void_p), ("wimg", C.c_void_p),
        ("a1", C.c_void_p), ("a2", C.c_void_p), ("xstash", C.c_void_p),
    ]


class TowerGrads(C.Structure):
    """Mirror of ``cfm_tower_grads_t``."""
    _fields_ = [
        ("g_out", C.c_void_p),
        ("dw1", C.c_void_p), ("db1", C.c_void_p), ("dw2", C.c_void_p), ("db2", C.c_void_p),
        ("dw3", C.c_void_p), ("db3", C.c_void_p),
        ("dbn1_w", C.c_void_p), ("dbn1_b", C.c_void_p), ("dbn2_w", C.c_void_p), ("dbn2_b", C.c_void_p),
        ("dy1", C.c_void_p), ("dy2", C.c_void_p), ("dx_emb", C.c_void_p), ("dx_num", C.c_void_p),
    ]


class EmbGroup(C.Structure):
    """Mirror of ``cfm_emb_group_t`` (one tower's tables in the joint embedding-gradient reduce)."""
    _fields_ = [
        ("x_cat", C.c_void_p), ("dx_emb", C.c_void_p), ("n_tables", i64), ("emb_dim", i64),
        ("grad_tables", C.c_void_p * CFM_MAX_TABLES), ("table_rows", i64 * CFM_MAX_TABLES),
    ]


class PeerTable(C.Structure):
    """Mirror of ``cfm_peer_table_t`` (one owned table of the NVLink table-sharded mode)."""
    _fields_ = [
        ("n_cols", i64), ("col", i64), ("col0", i64), ("rows", i64), ("grad", C.c_void_p),
        ("x_cat", C.c_void_p * CFM_MAX_PEERS), ("dx_emb", C.c_void_p * CFM_MAX_PEERS),
    ]


class Projector(C.Structure):
    """Mirror of ``cfm_projector_t``."""
    _fields_ = [("d_in", i64), ("d_hid", i64), ("d_out", i64),
                ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
                ("x", C.c_void_p), ("hid", C.c_void_p), ("raw", C.c_void_p), ("out", C.c_void_p)]


class ProjectorGrads(C.Structure):
    """Mirror of ``cfm_projector_grads_t``."""
    _fields_ = [("g_out", C.c_void_p), ("dx", C.c_void_p), ("dw1", C.c_void_p), ("db1", C.c_void_p),
                ("dw2", C.c_void_p), ("db2", C.c_void_p), ("scratch", C.c_void_p)]


class AdamTensor(C.Structure):
    """Mirror of ``cfm_adam_tensor_t``."""
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", i64)]


class PeerGroup(C.Structure):
    """Mirror of ``cfm_peer_group_t``."""
    _fields_ = [("owned", C.POINTER(PeerTable)), ("n_owned", i64), ("emb_dim", i64), ("width", i64)]


# name -> (restype, argtypes); every symbol include/cfm_b200.h declares
_V, _I, _D, _U64 = C.c_void_p, i64, C.c_double, C.c_uint64
PROTOTYPES = {
    "cfm_abi_version": (C.c_int, []),
    "cfm_last_error": (C.c_char_p, []),
    "cfm_device_info": (C.c_int, [C.POINTER(i64)] * 4),
    "cfm_tower_scratch_floats": (i64, [C.POINTER(Tower)]),
    "cfm_tower_wimg_floats": (i64, [C.POINTER(Tower)]),
    "cfm_tower_xstash_floats": (i64, [C.POINTER(Tower), _I]),
    "cfm_launch_count": (i64, [_I]),
    "cfm_profile_enable": (C.c_int, [_I]),
    "cfm_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(i64), _I]),
    "cfm_towers_fwd": (C.c_int, [C.POINTER(Tower), _I, _I, _I, _U64, _U64, _V, _V, _V]),
    "cfm_counter_advance": (C.c_int, [_V, _U64, _V]),
    "cfm_towers_bwd": (C.c_int, [C.POINTER(Tower), C.POINTER(TowerGrads), _I, _I, _I, _U64, _U64, _V, _V]),
    "cfm_dropout_mask": (C.c_int, [_V, _I, _I, _D, _I, _I, _U64, _U64, _V]),
    "cfm_emb_grad_tmp_bytes": (i64, [_I, _I]),
    "cfm_emb_grad_segment_reduce": (C.c_int, [_V, _V, _I, _I, _I, C.POINTER(C.c_void_p), C.POINTER(i64),
                                              _V, _V, _V, _V, _V, _I, _V]),
    "cfm_emb_grad_rezero": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(i64), _I, _I, _V, _I, _V]),
    "cfm_emb_grad_joint_reduce": (C.c_int, [C.POINTER(EmbGroup), _I, _I, _I, _V, _V, _V, _V, _V, _I, _V]),
    "cfm_emb_grad_joint_rezero": (C.c_int, [C.POINTER(EmbGroup), _I, _I, _V, _V]),
    "cfm_enable_peer_access": (C.c_int, [_I]),
    "cfm_ipc_export": (C.c_int, [_V, C.c_char_p, C.POINTER(i64)]),
    "cfm_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "cfm_ipc_close": (C.c_int, [_V]),
    "cfm_emb_gather_rows": (C.c_int, [_V, _I, _I, _I, _I, C.POINTER(C.c_void_p), C.POINTER(i64), _V, _V, _V]),
    "cfm_emb_grad_peer_reduce": (C.c_int, [C.POINTER(PeerGroup), _I, _I, _I, _I, _V, _V, _V, _V, _V, _I, _V]),
    "cfm_emb_grad_peer_rezero": (C.c_int, [C.POINTER(PeerGroup), _I, _I, _I, _V, _V]),
    "cfm_adam_step": (C.c_int, [C.POINTER(AdamTensor), _I, _V, _I, _D, _D, _D, _D, _I, _V]),
    "cfm_cosine_head_fwd": (C.c_int, [_V, _V, _V, _I, _I, _D, _V, _V, _V, _V, _V, _V, _V, _V]),
    "cfm_cosine_head_bwd": (C.c_int, [_V, _V, _V, _V, _V, _V, _V, _V, _V, _I, _I, _D, _V, _V, _V, _V, _V]),
    "cfm_projector_scratch_floats": (i64, [C.POINTER(Projector), _I]),
    "cfm_projector_fwd": (C.c_int, [C.POINTER(Projector), _I, _I, _D, _V]),
    "cfm_projector_bwd": (C.c_int, [C.POINTER(Projector), C.POINTER(ProjectorGrads), _I, _I, _D, _V]),
    "cfm_structural_head": (C.c_int, [_V, _V, _V, _V, _V, _V, _I, _D, _V, _V, _V, _V, _V, _V]),
    "cfm_simtile_chunks": (i64, [_I, _I]),
    "cfm_simtile_set_rb": (C.c_int, [_I]),
    "cfm_simtile_set_poly": (C.c_int, [_I]),
    "cfm_pack_rows_bf16": (C.c_int, [_V, _I, _I, _I, _V, _V]),
    "cfm_infonce_rowsum": (C.c_int, [_V, _V, _I, _I, _I, _D, _I, _V, _V, _V, _V]),
    "cfm_infonce_colpart_floats": (i64, [_I, _I]),
    "cfm_infonce_rowcolsum": (C.c_int, [_V, _V, _I, _I, _I, _D, _I, _V, _V, _V, _V, _V, _V]),
    "cfm_infonce_loss": (C.c_int, [_V, _V, _V, _I, _D, _I, _V, _V]),
    "cfm_infonce_grad": (C.c_int, [_V, _V, _I, _I, _I, _I, _D, _I, _I, _V, _V, _V, _V, _V, _V, _V]),
    "cfm_simtile_scores": (C.c_int, [_V, _V, _I, _I, _I, _V, _V]),
    "cfm_pack_rows_f16": (C.c_int, [_V, _I, _I, _I, _V, _V]),
    "cfm_allpairs_topk": (C.c_int, [_V, _V, _V, _V, _I, _I, _I, _I, _I, _I, _D, _D, _I, _V, _V, _V, _V, _V, _V, _V, _V]),
    "cfm_topk_merge": (C.c_int, [_V, _I, _V, _I, _I, _I, _V, _V, _V]),
    "cfm_allpairs_rank": (C.c_int, [_V, _V, _I, _I, _I, _V, _V, _V]),
    "cfm_allpairs_diag_rank": (C.c_int, [_V, _V, _V, _V, _I, _I, _I, _I, _I, _I, _D, _V, _V, _V, _V, _V, _V, _V, _I, _V, _V]),
    "cfm_debug_set_trace": (C.c_int, [_V, _I]),
    "cfm_tc_mma_probe": (C.c_int, [_V, _I, _I, _I, _I, _V]),
    "cfm_tc_selftest": (C.c_int, [_V, _V, _V, _I, _I, _I, _I, _I, _V]),
}

_lib: Optional[C.CDLL] = None


class CfmError(RuntimeError):
    """A libcfm_b200 entry point returned a negative status."""


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the CUDA library is not built (run `python __graft_entry__.py` or "
                f"`make -C ceo-recommender_b200/csrc`). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)      # AttributeError if the .so lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        if handle.cfm_abi_version() != CFM_ABI_VERSION:
            raise ImportError(f"libcfm_b200 ABI {handle.cfm_abi_version()} != binding {CFM_ABI_VERSION}")
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc == 0:
        return
    msg = lib().cfm_last_error().decode("utf-8", "replace")
    if rc == -4:
        raise ValueError(msg)      # torch raises ValueError for train-mode BatchNorm with one row
    raise CfmError(f"libcfm_b200 error {rc}: {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Raw device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("ceo_firm_matching (B200 build) has no CPU path: tensor is on " + str(t.device))
    if not t.is_contiguous():
        raise RuntimeError("libcfm_b200 needs contiguous tensors")
    return t.data_ptr()


class PeerView:
    """A buffer of another process of the node, mapped into this one through CUDA IPC (``cfm_ipc_open``): just
    enough of the tensor interface for ``ptr()``.  The memory is owned by the exporting process."""
    is_cuda = True

    def __init__(self, address: int, shape, dtype, owner_rank: int):
        self._address, self.shape, self.dtype, self.owner_rank = int(address), tuple(shape), dtype, owner_rank

    def data_ptr(self) -> int:
        return self._address

    def is_contiguous(self) -> bool:
        return True

    @property
    def device(self):
        return f"peer:{self.owner_rank}"


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


_dev_info = {}


def device_info() -> dict:
    dev = torch.cuda.current_device()
    if dev not in _dev_info:
        vals = [i64() for _ in range(4)]
        check(lib().cfm_device_info(*[C.byref(v) for v in vals]))
        _dev_info[dev] = dict(zip(("sm_count", "cc_major", "cc_minor", "tower_ctas"), (v.value for v in vals)))
    return _dev_info[dev]
