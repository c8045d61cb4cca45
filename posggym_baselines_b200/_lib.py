"""ctypes binding of libposg_b200.so (C-ABI declared in include/posg_b200.h)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("POSG_B200_LIB") or os.path.join(_HERE, "libposg_b200.so")  # override: kernel A/B runs


class PosgB200Error(RuntimeError):
    """Raised for every non-zero return of the C-ABI (and a missing library)."""


_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p
_i64 = C.c_int64
_u64 = C.c_uint64
_u32 = C.c_uint32
_int = C.c_int
_dbl = C.c_double

# name -> (restype, argtypes); mirrors include/posg_b200.h one to one
SIGNATURES = {
    "posg_last_error": (C.c_char_p, []),
    "posg_version": (_int, []),
    "posg_device_count": (_int, []),
    "posg_model_destroy": (_int, [_vp]),
    "posg_model_blob_bytes": (_int, [_vp]),
    "posg_model_num_agents": (_int, [_vp]),
    "posg_driving_create": (_int, [C.c_char_p, _int, _int, _int, _int, _int, C.POINTER(_vp)]),
    "posg_driving_sample_initial_state": (_int, [_vp, _i64, _vp, _u64, _u64, _vp, _vp]),
    "posg_driving_observe": (_int, [_vp, _i64, _vp, _vp, _vp]),
    "posg_driving_step": (
        _int,
        [_vp, _i64, _vp, _vp, _vp, _u64, _u32, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    ),
    "posg_driving_rollout": (
        _int,
        [_vp, _i64, _vp, _int, _int, _dbl, _u64, _u32, _u64, _vp, _vp, _vp, _vp, _vp],
    ),
    "posg_driving_rollout_host": (
        _int,
        [_vp, _i64, _vp, _int, _int, _dbl, _u64, _u32, _u64, _vp, _vp, _vp, _vp],
    ),
    "posg_driving_sweep_host": (_int, [_vp, _i64, _vp, _int, _u64, _u64, _vp, _vp, _vp, _vp, _vp]),
    "posg_driving_sweep_host_compact": (_int, [_vp, _i64, _vp, _int, _u64, _u64, _vp, _vp, _vp, _vp, _vp]),
    "posg_driving_decode_rewards_host": (_int, [_vp, _i64, _vp, _vp, _vp]),
    "posg_pp_create": (_int, [_int, C.c_char_p, _int, _int, _int, _int, _int, C.POINTER(_vp)]),
    "posg_lbf_create": (_int, [_int, _int, _int, _int, _int, _int, _int, _vp, _vp, C.POINTER(_vp)]),
    "posg_model_sample_initial_state": (_int, [_vp, _i64, _vp, _u64, _u64, _vp, _vp]),
    "posg_model_observe": (_int, [_vp, _i64, _vp, _vp, _vp]),
    "posg_model_step": (
        _int, [_vp, _i64, _vp, _vp, _vp, _u64, _u32, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "posg_model_step_host": (_int, [_vp, _i64, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp]),
    "posg_model_rollout": (
        _int, [_vp, _i64, _vp, _int, _int, _dbl, _u64, _u32, _u64, _vp, _vp, _vp, _vp, _vp]),
    "posg_trace_hash": (_int, [_i64, _vp, _int, _u64, _vp, _int, _vp, _int, _vp, _int, _vp, _vp, _vp]),
    "posg_nested_create": (_int, [_vp, C.POINTER(_vp)]),
    "posg_search_nested": (_int, [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _int, _int, _int, _vp, _i64, _i64, _u64,
                                   _vp, _vp]),
    "posg_search_multi": (_int, [_vp, _vp, _i64, _i64, _int, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _u64, _vp, _vp]),
    "posg_tree_advance_batch": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "posg_tree_node_info": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "posg_belief_reinvigorate_nested": (
        _int, [_vp, _vp, _i64, _i64, _int, _vp, _i64, _int, _int, _vp, _i64, _i64, _u64, _u64, _vp, _vp, _vp, _vp]),
    "posg_vec_flat_dim": (_int, [_vp]),
    "posg_vec_step": (_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _u64, _u32, _u64, _int, _int, _vp, _vp, _vp, _vp,
                              _vp, _vp, _vp, _dbl, _vp, _vp, _vp, _vp]),
    "posg_vec_embed": (_int, [_vp, _i64, _vp, _vp, _vp, _int, _int, _vp, _vp]),
    "posg_vec_step_embed": (_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _u64, _u32, _u64, _int, _vp, _vp, _vp, _vp, _vp, _vp,
                                   _vp, _vp, _vp, _int, _int, _vp, _vp, _vp]),
    "posg_policy_create": (_int, [_int, _int] + [_vp] * 14 + [C.POINTER(_vp)]),
    "posg_policy_destroy": (_int, [_vp]),
    "posg_policy_set_trunk0": (_int, [_vp, _int, _vp, _vp]),
    "posg_policy_eval_host": (_int, [_vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "posg_policy_step": (_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _u32, _vp, _u64, _vp, _vp, _vp, _vp, _vp,
                                 _vp, _vp, _vp, _vp]),
    "posg_seq_chunks": (_int, [_int, _i64, _vp, _int, _vp, _vp, _vp, _i64, _vp, _vp]),
    "posg_seq_gather": (_int, [_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _int, _int, _vp]),
    "posg_gae": (_int, [_int, _i64, _vp, _vp, _vp, _dbl, _dbl, _vp, _vp, _vp]),
    "posg_model_info": (_int, [_vp, C.POINTER(_int), C.POINTER(_int), C.POINTER(_int)]),
    "posg_tree_arena_bytes": (_i64, [_vp, _i64, _i64]),
    "posg_tree_reset": (_int, [_vp, _vp, _i64, _i64, _int, _dbl, _dbl, _vp]),
    "posg_search": (_int, [_vp, _vp, _i64, _i64, _int, _int, _vp, _i64, _i64, _u64, _vp, _vp]),
    "posg_search_waves": (_int, [_vp, _vp, _i64, _i64, _int, _int, _vp, _i64, _i64, _int, _u64, _vp, _vp]),
    "posg_search_warmup": (_int, [_vp]),
    "posg_tree_update": (_int, [_vp, _vp, _vp, _i64, _i64, _int, _int, _vp, _int, _vp, _i64, _vp, _i64, _i64, _i64, _dbl,
                                _u64, _u64, _vp, _vp, _vp, _vp]),
    "posg_search_forest": (_int, [_vp, _vp, _int, _i64, _i64, _int, _int, _int, _u64, _vp, _vp]),
    "posg_forest_stats": (_int, [_vp, _vp, _int, _i64, _i64, _vp, _vp]),
    "posg_tree_node_stats": (_int, [_vp, _vp, _i64, _i64, _int, _vp, _vp]),
    "posg_tree_root_stats_device": (_int, [_vp, _vp, _i64, _i64, _int, _vp, _vp]),
    "posg_tree_advance": (_int, [_vp, _vp, _i64, _i64, _int, _int, _vp, _int, _vp, _vp]),
    "posg_tree_reroot": (_int, [_vp, _vp, _vp, _i64, _i64, _int, _vp, _vp, _vp]),
    "posg_tree_gather_belief": (_int, [_vp, _vp, _i64, _i64, _int, _vp, _i64, _vp, _vp]),
    "posg_belief_reinvigorate": (
        _int, [_vp, _vp, _i64, _int, _int, _vp, _i64, _i64, _u64, _u64, _vp, _vp, _vp, _vp]),
    "posg_belief_initial": (_int, [_vp, _int, _vp, _i64, _i64, _u64, _u64, _vp, _vp, _vp]),
}


class SearchParams(C.Structure):
    """Mirror of posg_search_params (include/posg_b200.h)."""

    _fields_ = [
        ("discount", _dbl), ("c", _dbl), ("known_min", _dbl), ("known_max", _dbl),
        ("pucb_exploration_fraction", _dbl), ("depth_limit", _int), ("step_limit", _int),
        ("selection", _int), ("ego_agent", _int), ("intmcp_depth_rule", _int), ("seed", _u64),
        ("n_waves", _int), ("pad_", _int),
    ]

_lib: Optional[C.CDLL] = None


def library_path() -> str:
    return _LIB_PATH


def lib() -> C.CDLL:
    """Load (once) and return the C-ABI library.  Fails loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise PosgB200Error(
                f"{_LIB_PATH} is missing: build it with `python -m posggym_baselines_b200.build` "
                "(nvcc, sm_100a).  There is no CPU fallback."
            )
        handle = C.CDLL(_LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the header and the .so diverge
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> int:
    if rc < 0:
        msg = lib().posg_last_error()
        raise PosgB200Error(f"libposg_b200 error {rc}: {msg.decode() if msg else '?'}")
    return rc


def ptr(t) -> Optional[int]:
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        assert t.is_contiguous(), "libposg_b200 needs contiguous buffers"
        return t.data_ptr()
    assert t.flags["C_CONTIGUOUS"]
    return t.ctypes.data


def stream_ptr(stream=None) -> Optional[int]:
    import torch

    s = torch.cuda.current_stream() if stream is None else stream
    return s.cuda_stream or None
