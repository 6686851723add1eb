"""Multi-GPU sharding logic on CPU: world_size-2 gloo processes, each running its shard of the H_eff*phi
chain through the C ABI (oracle-backed mock, host pointers) and re-replicating H*phi with one all-reduce.
The result must equal the unsharded oracle chain bit for bit (every element has exactly one owner)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "oracle", "_ref", "libitb200_mock.so")


def _worker(rank, world, port, q):
    os.environ["ITB200_LIB_PATH"] = MOCK
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import ctypes as C

    import torch
    import torch.distributed as dist

    import itensor_b200 as itb
    from itensor_b200 import synth
    from itensor_b200._lib import check, lib
    from itensor_b200.shard import shard_chain

    dist.init_process_group("gloo", rank=rank, world_size=world)
    structs = synth.heff_chain([2, 5, 9, 6, 3], [3, 6, 8, 4, 1])
    hosts = [synth.random_values(s, 10 + i) for i, s in enumerate(structs)]
    plans, s = [], structs[0]
    for t in structs[1:]:
        p = itb.ContractPlan(s, t)
        plans.append(p)
        s = p.C
    sh = shard_chain(plans, world, rank)
    assert 0 < sh.my_flops < sh.total_flops
    ctx = C.c_void_p()
    check(lib().itb_ctx_create(0, C.byref(ctx)))
    cur = hosts[0]
    for k, p in enumerate(plans):
        # unowned blocks stay NaN in the intermediates: they must never be read by this rank
        out = np.full(p.C.nelems, np.nan) if k < 3 else np.zeros(p.C.nelems)
        check(lib().itb_contract_run(ctx, p._h, cur.ctypes.data_as(C.c_void_p), hosts[k + 1].ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        cur = out
    t = torch.from_numpy(cur)
    sh.allgather(t)
    q.put((rank, t.numpy().copy(), sh.owner.copy(), sh.my_flops / sh.total_flops))
    dist.destroy_process_group()


@pytest.mark.skipif(not os.path.exists(MOCK), reason="oracle/_ref/libitb200_mock.so not built")
def test_sharded_chain_equals_unsharded_world2():
    sys.path.insert(0, ROOT)
    from itensor_b200 import synth
    from oracle import orc

    world = 2
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_worker, args=(r, world, 29533, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    structs = synth.heff_chain([2, 5, 9, 6, 3], [3, 6, 8, 4, 1])
    hosts = [synth.random_values(s, 10 + i) for i, s in enumerate(structs)]
    s, v = structs[0], hosts[0]
    import itensor_b200 as itb

    for t, h in zip(structs[1:], hosts[1:]):
        cs, tr, v = orc.contract(s, v, t, h)
        s = itb.ContractPlan(s, t).C
    for rank, got, owner, share in res:
        assert np.array_equal(got, v)          # exact: one owner per element, zeros elsewhere
        assert set(owner.tolist()) == {0, 1}
        assert 0.2 < share < 0.8
    assert abs(sum(r[3] for r in res) - 1.0) < 1e-12


def test_sector_assignment_is_lpt():
    sys.path.insert(0, ROOT)
    from itensor_b200.shard import sector_assignment

    w = [36, 2209, 45796, 257049, 401956, 174724, 21025, 729, 9]
    for world in (2, 4, 8):
        owner = sector_assignment(w, world)
        loads = np.bincount(owner, weights=w, minlength=world)
        assert loads.max() <= max(max(w), sum(w) / world * 1.34)
