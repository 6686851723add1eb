"""The C-ABI library loads, exports every symbol include/itb200.h declares, and refuses to run
without a device (no CPU fallback). No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import itensor_b200 as itb
from itensor_b200 import _lib
from itensor_b200.tensor import BlockStruct, Index, PermutePlan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_header_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "itb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(itb_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 45
    L = itb.lib()
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in itb200.h but not exported by libitb200.so"
    assert names == set(_lib.SYMBOLS), (names ^ set(_lib.SYMBOLS))


def test_version_and_error_string():
    L = itb.lib()
    assert b"sm_100a" in L.itb_version()


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    h = C.c_void_p()
    rc = itb.lib().itb_ctx_create(0, C.byref(h))
    assert rc == _lib.ITB_ERR_CUDA
    assert b"no CPU fallback" in itb.lib().itb_last_error()
    with pytest.raises(RuntimeError):
        itb.Context(0)


def test_planner_rejects_malformed_input():
    i = Index(1, (2, 3), ((0,), (1,)), 1)
    j = Index(2, (2, 3), ((0,), (1,)), 1)
    bad = Index(1, (2, 4), ((0,), (1,)), -1)  # same identity as i, different sector sizes
    A = BlockStruct([i, j], np.array([[0, 0], [1, 1]], np.int32))
    B = BlockStruct([bad], np.array([[0]], np.int32))
    with pytest.raises(itb.ItbError) as e:
        itb.ContractPlan(A, B)
    assert e.value.code == _lib.ITB_ERR_INVALID
    with pytest.raises(itb.ItbError):
        PermutePlan(A, A, [0, 0])  # not a permutation
    # destination lacking the image of a source block
    D = BlockStruct([j, i], np.array([[0, 0]], np.int32))
    with pytest.raises(itb.ItbError):
        PermutePlan(A, D, [1, 0])
    # block coordinate out of range
    bad_blocks = BlockStruct([i, j], np.array([[0, 0]], np.int32))
    bad_blocks.blocks[0, 1] = 5  # corrupt after construction: the C planner must catch it
    with pytest.raises(itb.ItbError):
        itb.ContractPlan(bad_blocks, BlockStruct([j.dag()], np.array([[0]], np.int32)))
