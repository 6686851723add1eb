#!/usr/bin/env python
"""Generate golden vectors from the UNMODIFIED reference (oracle/_ref/libitref.so).

Run in the build container (where /root/reference exists and `make -C oracle` has been run):
    python tests/golden/make_golden.py
Writes tests/golden/golden_v1.npz: for every case the inputs (structure + seeds) are implicit in the
generator below (seeded NumPy), the OUTPUTS stored are exactly what the reference returned:
block lists, offsets, index order and element values of  A*B,  permute(T),  A += alpha*B,  norm(T).
tests/test_golden.py replays the same cases against the oracle (CPU) and the CUDA path (GPU).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from cases import contract_cases, permute_cases, pluseq_cases  # noqa: E402

from itensor_b200 import synth  # noqa: E402
from oracle import orc  # noqa: E402


def main():
    out = {}
    for name, A, B in contract_cases():
        a, b = synth.random_values(A, 1), synth.random_values(B, 2)
        R = orc.ref_contract(A, a, B, b)
        out[f"{name}/labels"] = R.labels
        out[f"{name}/blocks"] = R.blocks
        out[f"{name}/offsets"] = R.offsets
        out[f"{name}/data"] = R.data
        out[f"{name}/is_qn"] = np.array([R.is_qn, R.dtype, R.nelems])
    for name, S, new_inds in permute_cases():
        s = synth.random_values(S, 3)
        R = orc.ref_permute(S, s, new_inds)
        out[f"{name}/labels"] = R.labels
        out[f"{name}/blocks"] = R.blocks
        out[f"{name}/offsets"] = R.offsets
        out[f"{name}/data"] = R.data
        out[f"{name}/norm"] = np.array([orc.ref_norm(S, s)])
    for name, A, B, alpha in pluseq_cases():
        a, b = synth.random_values(A, 4), synth.random_values(B, 5)
        R = orc.ref_pluseq(A, a, B, b, alpha)
        out[f"{name}/labels"] = R.labels
        out[f"{name}/blocks"] = R.blocks
        out[f"{name}/offsets"] = R.offsets
        out[f"{name}/data"] = R.data
        out[f"{name}/dtype"] = np.array([R.dtype])
    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
