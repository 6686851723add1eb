"""itensor_b200 — B200-native (sm_100a) block-sparse contraction path for ITensor storage types.

Product = libitb200.so (C ABI in include/itb200.h, CUDA kernels in itensor_b200/csrc) plus the C++
storage-type plugin in itensor_b200/plugin. This Python package is the host mirror used by tests/bench.
"""
from ._lib import ITB_C64, ITB_F64, ItbError, lib  # noqa: F401
from .tensor import (BlockStruct, Context, ContractPlan, Index, PermutePlan, QTensor, add, contract, dag, elt, fill,  # noqa: F401
                     flux_blocks, norm, permute, permuted_struct, scale)
