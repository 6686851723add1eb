"""ctypes binding of the C ABI in include/itb200.h (libitb200.so).

The library is the product; this module only declares its entry points for the Python host mirror,
the tests and bench.py. Loading fails loudly when the shared object has not been built
(`python -c "import __graft_entry__ as g; g.build()"` or `make -C itensor_b200/csrc`): there is no
Python/NumPy/torch fallback for any of these calls.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libitb200.so")

ITB_OK = 0
ITB_ERR_INVALID, ITB_ERR_CUDA, ITB_ERR_UNSUPPORTED, ITB_ERR_NOMEM = -1, -2, -3, -4
ITB_F64, ITB_C64 = 0, 1


class TensorDesc(C.Structure):
    _fields_ = [
        ("order", C.c_int32),
        ("dtype", C.c_int32),
        ("nsect", C.POINTER(C.c_int32)),
        ("sect", C.POINTER(C.c_int64)),
        ("nblocks", C.c_int64),
        ("blocks", C.POINTER(C.c_int32)),
        ("offsets", C.POINTER(C.c_int64)),
        ("nelems", C.c_int64),
    ]


class ContractInfo(C.Structure):
    _fields_ = [
        ("c_order", C.c_int32),
        ("c_dtype", C.c_int32),
        ("c_nblocks", C.c_int64),
        ("c_nelems", C.c_int64),
        ("npairs", C.c_int64),
        ("flops", C.c_double),
        ("n_gemm_tiles", C.c_int64),
        ("n_skinny", C.c_int64),
        ("n_dot", C.c_int64),
        ("table_bytes", C.c_int64),
        ("class_flops", C.c_double * 5),
    ]


class ItbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"itb200 error {code}: {msg}")
        self.code = code


# every symbol include/itb200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_I32P, _I64P, _DP = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)
_DESCP = C.POINTER(TensorDesc)
SYMBOLS = {
    "itb_version": (C.c_char_p, []),
    "itb_last_error": (C.c_char_p, []),
    "itb_device_count": (C.c_int, []),
    "itb_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "itb_ctx_destroy": (C.c_int, [_P]),
    "itb_ctx_stream": (_P, [_P]),
    "itb_ctx_set_stream": (C.c_int, [_P, _P]),
    "itb_synchronize": (C.c_int, [_P]),
    "itb_launch_count": (C.c_int64, [_P]),
    "itb_malloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "itb_free": (C.c_int, [_P, _P]),
    "itb_memcpy_h2d": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "itb_memcpy_d2h": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "itb_memcpy_d2d": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "itb_memset0": (C.c_int, [_P, _P, C.c_size_t]),
    "itb_pool_trim": (C.c_int, [_P]),
    "itb_contract_plan_create": (C.c_int, [_DESCP, _I32P, _DESCP, _I32P, C.POINTER(_P)]),
    "itb_contract_plan_destroy": (C.c_int, [_P]),
    "itb_contract_plan_info": (C.c_int, [_P, C.POINTER(ContractInfo)]),
    "itb_contract_plan_shape": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _I64P, _I64P, _I64P, _DP]),
    "itb_contract_plan_c_labels": (C.c_int, [_P, _I32P]),
    "itb_contract_plan_c_nsect": (C.c_int, [_P, _I32P]),
    "itb_contract_plan_c_sect": (C.c_int, [_P, _I64P]),
    "itb_contract_plan_c_blocks": (C.c_int, [_P, _I32P]),
    "itb_contract_plan_c_offsets": (C.c_int, [_P, _I64P]),
    "itb_contract_plan_pairs": (C.c_int, [_P, _I64P]),
    "itb_contract_plan_set_cblock_range": (C.c_int, [_P, C.c_int64, C.c_int64]),
    "itb_contract_plan_set_cblock_mask": (C.c_int, [_P, C.POINTER(C.c_uint8)]),
    "itb_contract_plan_set_index_slices": (C.c_int, [_P, C.c_int32, _I64P, _I64P]),
    "itb_contract_plan_cblock_flops": (C.c_int, [_P, _DP]),
    "itb_comm_unique_id": (C.c_int, [_P]),
    "itb_comm_create": (C.c_int, [_P, C.c_int32, C.c_int32, _P, C.POINTER(_P)]),
    "itb_comm_world": (C.c_int32, [_P]),
    "itb_comm_rank": (C.c_int32, [_P]),
    "itb_comm_allgather": (C.c_int, [_P, _P, _P, _P, C.c_int64]),
    "itb_comm_destroy": (C.c_int, [_P]),
    "itb_ctx_device": (C.c_int, [_P]),
    "itb_p2p_alloc": (C.c_int, [_P, C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]),
    "itb_p2p_open": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_void_p)]),
    "itb_p2p_close": (C.c_int, [_P, _P]),
    "itb_p2p_free": (C.c_int, [_P, _P]),
    "itb_p2p_barrier_init": (C.c_int, [_P, _P]),
    "itb_p2p_barrier": (C.c_int, [_P, _P, C.POINTER(C.c_void_p), C.c_int32, C.c_int32]),
    "itb_p2p_barrier_status": (C.c_int, [_P, _P, _I64P, _I64P]),
    "itb_contract_plan_tiles": (C.c_int64, [_P, _I32P, C.c_int64]),
    "itb_contract_plan_cta_begin": (C.c_int64, [_P, _I32P, C.c_int64]),
    "itb_contract_plan_rowgroups": (C.c_int64, [_P, _I64P, C.c_int64]),
    "itb_contract_plan_cblks": (C.c_int64, [_P, _I64P, C.c_int64]),
    "itb_flux_blocks": (C.c_int64, [C.c_int32, _I32P, _I32P, C.c_int32, _I32P, _I32P, _I32P, _I32P, C.c_int64]),
    "itb_contract_run": (C.c_int, [_P, _P, _P, _P, _P]),
    "itb_contract_host": (C.c_int, [_P, _P, _P, _P, _P]),
    "itb_permute_plan_create": (C.c_int, [_DESCP, _DESCP, _I32P, C.POINTER(_P)]),
    "itb_blockcopy_plan_create": (C.c_int, [C.c_int64, _P, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "itb_permute_plan_destroy": (C.c_int, [_P]),
    "itb_permute_plan_bytes": (C.c_int64, [_P]),
    "itb_permute_run": (C.c_int, [_P, _P, _P, _P, C.c_double, C.c_double, C.c_int]),
    "itb_permute_host": (C.c_int, [_P, _P, _P, _P, C.c_double, C.c_double, C.c_int]),
    "itb_nrm2": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _DP]),
    "itb_scal": (C.c_int, [_P, C.c_int32, C.c_int64, _P, C.c_double, C.c_double]),
    "itb_axpy": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_double, C.c_double, _P, _P]),
    "itb_fill": (C.c_int, [_P, C.c_int32, C.c_int64, _P, C.c_double, C.c_double]),
    "itb_conj": (C.c_int, [_P, C.c_int64, _P]),
    "itb_real_to_cplx": (C.c_int, [_P, C.c_int64, _P, _P]),
    "itb_take_part": (C.c_int, [_P, C.c_int64, _P, _P, C.c_int]),
    "itb_get_elt": (C.c_int, [_P, C.c_int32, _P, C.c_int64, _DP]),
    "itb_dot": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P, C.c_int, _DP]),
    "itb_syevd_host": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _DP, C.POINTER(C.c_int32)]),
    "itb_gesvd_host": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P, _DP, _P, _P, C.POINTER(C.c_int32)]),
    "itb_solver_ready": (C.c_int, []),
    "itb_svd_batch_run": (C.c_int, [_P, C.c_int32, C.c_int64, _I64P, _I32P, _I32P, _P, C.POINTER(_P)]),
    "itb_svd_batch_values": (C.c_int, [_P, _DP]),
    "itb_svd_batch_copy_u": (C.c_int, [_P, C.c_int64, C.c_int32, _P]),
    "itb_svd_batch_copy_v": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_int]),
    "itb_svd_batch_destroy": (C.c_int, [_P]),
    "itb_svd_batch_stats": (C.c_double, [_I64P]),
    "itb_eigh_batch_run": (C.c_int, [_P, C.c_int32, C.c_int64, _I64P, _I32P, _P, C.c_int, C.POINTER(_P)]),
    "itb_eigh_batch_values": (C.c_int, [_P, _DP]),
    "itb_eigh_batch_copy_vectors": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_int]),
    "itb_eigh_batch_destroy": (C.c_int, [_P]),
    "itb_peak_fp64": (C.c_int, [_P, C.c_int, C.c_int, _DP]),
    "itb_contract_plan_model_work": (C.c_int, [_P, _DP, _DP]),
    "itb_contract_run_mirrored": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.POINTER(C.c_void_p)]),
    "itb_contract_plan_refine": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.POINTER(C.c_double)]),
    "itb_ctx_set_profile": (C.c_int, [_P, C.c_int]),
    "itb_contract_last_cta_cycles": (C.c_int64, [_P, _I64P, C.c_int64]),
    "itb_contract_last_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "itb_contract_last_item_cycles": (C.c_int64, [_P, _I64P, C.c_int64]),
    "itb_timer_start": (C.c_int, [_P]),
    "itb_timer_stop_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libitb200.so once; raise (never fall back) if it is missing."""
    global _lib
    if _lib is None:
        path = LIB_PATH
        # tests only: ITB200_LIB_PATH points the bindings at the oracle-backed mock of the C ABI
        # (oracle/_ref/libitb200_mock.so) to exercise host logic without a GPU; never set in production
        path = os.environ.get("ITB200_LIB_PATH", path)
        if not os.path.exists(path):
            raise ImportError(
                f"{path} not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). itensor_b200 has no CPU fallback."
            )
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != ITB_OK:
        raise ItbError(rc, lib().itb_last_error().decode())
