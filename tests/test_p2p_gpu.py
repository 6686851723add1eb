"""Two GPUs, one process each: the row-sharded H_eff*phi chain with H*phi exchanged by DIRECT peer-memory stores
(ChainShard.prepare_p2p / push: itb_p2p_alloc + itb_p2p_open + block-copy kernel + arrival barrier) must equal the unsharded
chain on every rank, in both buffers, and agree with the pack / NCCL all-gather / scatter path."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import itensor_b200 as itb
    from itensor_b200 import synth
    from itensor_b200._lib import check, lib
    from itensor_b200.shard import shard_chain

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = itb.Context(rank)
    structs = synth.heff_chain(synth.gaussian_sectors(400, 7))
    hosts = [synth.random_values(s, 10 + i) for i, s in enumerate(structs)]
    dts = [itb.QTensor.from_host(ctx, s, h) for s, h in zip(structs, hosts)]
    # unsharded reference on this rank
    cur = dts[0]
    for t in dts[1:]:
        cur = itb.contract(cur, t)
    want = cur.data.clone()
    plans, s = [], structs[0]
    for t in structs[1:]:
        p = itb.ContractPlan(s, t)
        plans.append(p)
        s = p.C
    sh = shard_chain(plans, world, rank).prepare(ctx.empty)
    bufs = sh.prepare_p2p(ctx)
    assert bufs is not None, "peer memory unavailable between two GPUs of one box"
    outs = [itb.QTensor(ctx, p.C, ctx.empty(p.C.nreal)) for p in plans]
    errs = []
    for b in (0, 1, 0):
        bufs[b].fill_(float("nan"))
        dist.barrier()
        cur = dts[0]
        for k, p in enumerate(plans):
            dst = bufs[b] if k == 3 else outs[k].data
            import ctypes as C
            src_ptr = dts[0].ptr if k == 0 else C.c_void_p(prev_ptr)
            check(lib().itb_contract_run(ctx.handle, p._h, src_ptr, dts[k + 1].ptr, C.c_void_p(dst.data_ptr())))
            prev_ptr = dst.data_ptr()
        sh.push(ctx.handle, b)
        torch.cuda.synchronize()
        errs.append(float(((bufs[b] - want).abs().max() / want.abs().max()).item()))
    # the exchange fused into the last contraction's epilogue (itb_contract_run_mirrored)
    import ctypes as C
    for b in (1, 0):
        bufs[b].fill_(float("nan"))
        dist.barrier()
        cur_ptr = dts[0].ptr
        for k, p in enumerate(plans[:3]):
            check(lib().itb_contract_run(ctx.handle, p._h, cur_ptr, dts[k + 1].ptr, outs[k].ptr))
            cur_ptr = outs[k].ptr
        sh.run_last_mirrored(ctx.handle, plans[3], cur_ptr, dts[4].ptr, b)
        torch.cuda.synchronize()
        errs.append(float(((bufs[b] - want).abs().max() / want.abs().max()).item()))
    assert sh.barrier_status()[1] == 0
    # the all-gather path on the same sliced plans
    cur_ptr = dts[0].ptr
    for k, p in enumerate(plans):
        check(lib().itb_contract_run(ctx.handle, p._h, cur_ptr, dts[k + 1].ptr, outs[k].ptr))
        cur_ptr = outs[k].ptr
    sh.allgather(ctx.handle, outs[-1].data)
    torch.cuda.synchronize()
    errs.append(float(((outs[-1].data - want).abs().max() / want.abs().max()).item()))
    q.put((rank, errs))
    dist.barrier()
    sh.close_p2p()
    sh.close()
    dist.destroy_process_group()


def test_p2p_row_exchange_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_worker, args=(r, 2, 29577, q), daemon=True) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=900) for _ in range(2)]   # (a cold box pages in cuSOLVER / cuBLAS / NCCL for minutes)
        for p in procs:
            p.join(timeout=60)
    finally:   # a rank that raised leaves its peer inside a collective: never wait for it
        for p in procs:
            if p.is_alive():
                p.kill()
                p.join(timeout=10)
    for rank, errs in res:
        assert all(e == e and e <= 1e-12 for e in errs), (rank, errs)
