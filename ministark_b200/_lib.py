"""ctypes binding of libministark.so (include/ministark.h).  No fallback: if the CUDA library is
missing or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MINISTARK_LIB") or os.path.join(HERE, "libministark.so")  # env: A/B builds

MS_OK = 0
ERR_NAMES = {
    1: "MS_ERR_BAD_SHAPE", 2: "MS_ERR_CUDA", 3: "MS_ERR_NCCL", 4: "MS_ERR_QUOTIENT_NONZERO",
    5: "MS_ERR_TRANSCRIPT", 6: "MS_ERR_UNSUPPORTED", 7: "MS_ERR_LEAF_NOT_FOUND", 8: "MS_ERR_BUFFER_TOO_SMALL",
}


class MiniStarkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class StarkParams(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("security_bits", "blowup_factor", "steps", "trace_columns", "inner_children")]


TRACE_COMMIT_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint8))
LDE_COMMIT_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint8))


class CommitHooks(C.Structure):
    """ms_commit_hooks (include/ministark.h)"""
    _fields_ = [("user", C.c_void_p), ("trace_commit", TRACE_COMMIT_FN), ("lde_commit", LDE_COMMIT_FN),
                ("replica_only", C.c_int32), ("download_rank", C.c_int32), ("download_world", C.c_int32)]


# every symbol include/ministark.h declares: name -> (restype, argtypes)
_u64, _i32, _vp, _sz = C.c_uint64, C.c_int32, C.c_void_p, C.c_size_t
SIGNATURES = {
    "ms_version": (_i32, []),
    "ms_ctx_create": (_i32, [_i32, _i32, _vp, C.POINTER(_vp)]),
    "ms_ctx_destroy": (None, [_vp]),
    "ms_last_error": (C.c_char_p, [_vp]),
    "ms_sync": (_i32, [_vp]),
    "ms_launch_count": (_u64, [_vp]),
    "ms_set_zero_display": (_i32, [_vp, _i32]),
    "ms_set_transcript_option": (_i32, [_vp, _i32, _i32]),
    "ms_set_profiling": (_i32, [_vp, _i32]),
    "ms_profile_collect": (_i32, [_vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_uint32), _i32]),
    "ms_selftest_field_ops": (_i32, [_vp, _u64, C.POINTER(_u64)]),
    "ms_dev_alloc": (_i32, [_vp, _sz, C.POINTER(_vp)]),
    "ms_dev_free": (_i32, [_vp, _vp]),
    "ms_h2d": (_i32, [_vp, _vp, _vp, _sz]),
    "ms_d2h": (_i32, [_vp, _vp, _vp, _sz]),
    "ms_transpose_rm_to_cm": (_i32, [_vp, _vp, _u64, _u64, _vp]),
    "ms_transpose_cm_to_rm": (_i32, [_vp, _vp, _u64, _u64, _vp]),
    "ms_trace_synth": (_i32, [_vp, _u64, _u64, _u64, _vp]),
    "ms_trace_recurrence": (_i32, [_vp, _vp, _vp, _u64, _u64, _u64, _u64, _vp]),
    "ms_merkle_commit": (_i32, [_vp, _vp, _u64, _u64, _u64, _i32, _u64, _u64, _vp, _vp]),
    "ms_merkle_node_count": (_u64, [_u64, _u64]),
    "ms_intt_columns": (_i32, [_vp, _vp, _u64, _u64, _u64, _vp, _u64]),
    "ms_linear_constraints": (_i32, [_vp, _vp, _u64, _u64, _u64, _vp, _u64, _vp, _u64]),
    "ms_coset_lde": (_i32, [_vp, _vp, _u64, _u64, _u64, _u64, _u64, _vp, _u64]),
    "ms_coset_lde_host": (_i32, [_vp, _vp, _u64, _u64, _u64, _u64, _vp]),
    "ms_mix": (_i32, [_vp, _vp, _u64, _u64, _u64, _u64, _vp]),
    "ms_deep_open": (_i32, [_vp, _vp, _u64, _u64, _u64, _vp, _u64, _vp]),
    "ms_fri_commit": (_i32, [_vp, _vp, _u64, _u64, _u64, _vp, _u64, _vp, _vp]),
    "ms_fri_deep_coeffs": (_i32, [_vp, _vp, _u64, _u64, _vp, _vp]),
    "ms_fri_fold": (_i32, [_vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _u64]),
    "ms_fri_query": (_i32, [_vp, _vp, _u64, _u64, _vp, _u64, _u64, _vp, _vp, _u64, _vp, _u64, _vp, _vp, _vp, _vp, _vp]),
    "ms_stark_derive": (_i32, [_i32, C.POINTER(StarkParams), C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]),
    "ms_stark_proof_bound": (_u64, [_i32, C.POINTER(StarkParams), _u64, _u64]),
    "ms_stark_prove": (_i32, [_vp, C.POINTER(StarkParams), _vp, _u64, _u64, _vp, _u64, _vp, C.POINTER(_u64)]),
    "ms_stark_prove_affine": (_i32, [_vp, C.POINTER(StarkParams), _vp, _u64, _u64, _vp, _vp, _u64, _vp, C.POINTER(_u64)]),
    "ms_stark_prove_device": (_i32, [_vp, C.POINTER(StarkParams), _vp, _u64, _u64, _vp, _u64, _vp, C.POINTER(_u64)]),
    "ms_stark_prove_hooked": (_i32, [_vp, C.POINTER(StarkParams), _vp, _u64, _u64, _vp, _u64, _vp, _vp, C.POINTER(_u64)]),
    "ms_stark_verify": (_i32, [_vp, C.POINTER(StarkParams), _vp, _u64, _u64, _u64, _vp, _u64, _i32, C.POINTER(_i32), C.POINTER(_i32)]),
    "ms_comm_unique_id": (_i32, [_vp]),
    "ms_comm_init_nccl": (_i32, [_vp, _vp, _i32, _i32]),
    "ms_comm_init_local": (_i32, [C.POINTER(_vp), _i32]),
    "ms_comm_destroy": (_i32, [_vp]),
    "ms_comm_info": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(C.c_char_p)]),
    "ms_set_shard_mask": (_i32, [_vp, _i32]),
    "ms_shard_plan": (_i32, [_u64, _u64, _u64, _i32, _i32, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]),
    "ms_stark_prove_multi": (_i32, [_vp, C.POINTER(StarkParams), _vp, _u64, _u64, _vp, _u64, _vp, C.POINTER(_u64), _i32]),
    "ms_merkle_subtree": (_i32, [_vp, _vp, _u64, _u64, _u64, _i32, _u64, _u64, _vp, C.POINTER(_u64)]),
    "ms_merkle_reduce": (_i32, [_vp, _vp, _u64, _u64, _vp]),
    "ms_merkle_subtree_gather": (_i32, [_vp, _vp, _u64, _u64, _i32, _u64, _u64, _vp, C.POINTER(_u64)]),
    "ms_peer_alloc": (_i32, [_vp, _u64, C.POINTER(_vp)]),
    "ms_peer_free": (_i32, [_vp, _vp]),
    "ms_peer_export": (_i32, [_vp, _vp, _vp]),
    "ms_peer_open": (_i32, [_vp, _vp, C.POINTER(_vp)]),
    "ms_peer_close": (_i32, [_vp, _vp]),
    "ms_host_register": (_i32, [_vp, _vp, _u64]),
    "ms_host_unregister": (_i32, [_vp, _vp]),
    "ms_stark_last_timings": (_i32, [_vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), _i32]),
}

_lib = None


def load():
    """dlopen the in-tree library and bind every declared symbol (raises if anything is missing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m ministark_b200.build` (nvcc, sm_100a). "
                "There is no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib
