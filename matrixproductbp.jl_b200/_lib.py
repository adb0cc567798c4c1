"""ctypes binding of include/mpbp.h.  The CUDA library is mandatory: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmpbp_b200.so")

_lib = None

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_dp = C.POINTER(C.c_double)

SIGNATURES = {
    "mpbp_last_error": (C.c_char_p, []),
    "mpbp_version": (C.c_int, []),
    "mpbp_create": (C.c_int, [C.c_int64, C.c_int64, C.c_int, c_i32p, c_i64p, c_i64p, c_i64p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "mpbp_create_periodic": (C.c_int, [C.c_int64, C.c_int64, C.c_int, c_i32p, c_i64p, c_i64p, c_i64p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "mpbp_create_infinite": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "mpbp_create_infinite_bipartite": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "mpbp_destroy": (C.c_int, [C.c_void_p]),
    "mpbp_add_node_class": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_i32p, C.c_int, c_i32p, c_dp, C.c_int, c_i32p, c_i32p, c_dp, c_dp, c_dp, c_dp, c_i32p]),
    "mpbp_set_node_classes": (C.c_int, [C.c_void_p, c_i32p]),
    "mpbp_clear_node_classes": (C.c_int, [C.c_void_p]),
    "mpbp_add_generic_class": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_i32p, C.c_int, c_dp, c_i32p]),
    "mpbp_set_phi": (C.c_int, [C.c_void_p, c_dp]),
    "mpbp_set_psi": (C.c_int, [C.c_void_p, c_dp]),
    "mpbp_get_message": (C.c_int, [C.c_void_p, C.c_int64, c_i32p, c_dp, C.c_int64, c_i64p]),
    "mpbp_set_message": (C.c_int, [C.c_void_p, C.c_int64, c_i32p, c_dp]),
    "mpbp_reset_messages": (C.c_int, [C.c_void_p]),
    "mpbp_iterate": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, c_i64p, C.c_int64, c_i64p, c_dp, C.POINTER(C.c_int), c_dp]),
    "mpbp_beliefs": (C.c_int, [C.c_void_p, c_dp]),
    "mpbp_pair_beliefs": (C.c_int, [C.c_void_p, c_dp, c_dp]),
    "mpbp_free_energy": (C.c_int, [C.c_void_p, c_dp]),
    "mpbp_twovar_marginals": (C.c_int, [C.c_void_p, c_dp]),
    "mpbp_alternate_marginals": (C.c_int, [C.c_void_p, c_dp]),
    "mpbp_sample_prior": (C.c_int, [C.c_void_p, C.c_uint64, c_i32p]),
    "mpbp_message_slot_bytes": (C.c_int64, [C.c_void_p]),
    "mpbp_pack_messages_dev": (C.c_int, [C.c_void_p, C.c_int64, c_i64p, C.c_void_p]),
    "mpbp_unpack_messages_dev": (C.c_int, [C.c_void_p, C.c_int64, c_i64p, C.c_void_p]),
    "mpbp_counters": (C.c_int, [C.c_void_p, c_dp, C.c_int]),
    "mpbp_family_flops": (C.c_int, [C.c_void_p, c_dp]),
    "mpbp_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "mpbp_kernel_times": (C.c_int, [C.c_void_p, c_dp, C.c_int, C.c_int]),
    "mpbp_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mpbp_measure_fp64_peak": (C.c_int, [C.c_int, c_dp]),
    "mpbp_test_qr": (C.c_int, [c_dp, C.c_int, C.c_int, C.c_int, c_dp]),
    "mpbp_test_qr_ft": (C.c_int, [c_dp, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp]),
    "mpbp_test_svd": (C.c_int, [c_dp, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp]),
    "mpbp_test_jacobi": (C.c_int, [c_dp, C.c_int, C.c_int, C.c_int, c_dp, c_i32p]),
}


class MPBPError(RuntimeError):
    pass


def lib():
    """Load libmpbp_b200.so (built by __graft_entry__.build()).  Fails loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MPBPError(f"{LIB_PATH} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise MPBPError(lib().mpbp_last_error().decode())
