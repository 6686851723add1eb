"""Every device table the sm_100a kernels consume (block-pair tables, stream-K tile items and their split-K
reduction, row groups, C-stationary streaming items, split-K dot items; permute chunk / tile / zero-fill items) is
executed on the host by the table-walking
executor of the mock ABI (oracle/mock_itb200.cc, test infrastructure) and compared with the CPU oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "oracle", "_ref", "libitb200_mock.so")


@pytest.mark.skipif(not os.path.exists(MOCK), reason="oracle/_ref/libitb200_mock.so not built (make -C oracle mock)")
def test_device_tables_reproduce_the_oracle():
    env = dict(os.environ, ITB200_LIB_PATH=MOCK, ITB_MOCK_TABLES="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emulate_tables.py")], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "tables emulation ok" in out.stdout
    stats = eval(out.stdout.strip().split("tables emulation ok:")[-1])
    # the run really covered every kernel class
    assert stats["cases"] > 170 and stats["tiles"] > 200 and stats["split_pieces"] > 10
    assert stats["rowgroups"] > 10 and stats["skinny"] > 10 and stats["dots"] > 5
    assert stats["permutes"] > 200 and stats["zero_filled"] > 10
    assert stats["refined"] > 25  # re-cut partitions (itb_contract_plan_refine) walked as well
