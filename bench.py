#!/usr/bin/env python
"""bench.py — block-sparse contraction FP64 throughput on the DMRG effective-Hamiltonian product.

A "step" is one H_eff*phi = LocalOp::product (itensor/mps/localop.h:324-365): the four chained
block-sparse contractions phi*L, *W1, *W2, *R, on a synthetic S=1/2 Heisenberg centre-bond structure
(Sz-conserving QDense blocks) at maxdim m (BASELINE.json configs[1], default m=2000).

  value        whole-job FP64 TFLOP/s with all operands resident in HBM (algorithmic flops =
               sum over block pairs of 2*M*N*K), CUDA-event timed, L2 flushed between steps
  e2e          same metric through the C ABI with HOST buffers: per step H2D(phi,L,W1,W2,R) from
               pinned memory + table upload + 4 contractions + D2H(H phi)
  roofline     dominant kernel (bsc_gemm_kernel 128x128 DMMA tiles) against the FP64 tensor ceiling
               measured in this run (cuBLAS DGEMM 8192^3 via torch.matmul; MEASURED_PEAKS.json has no
               FP64 entry), per-launch times from CUDA events on the launching stream
  cpu_baseline the UNMODIFIED reference (oracle/_ref/libitref.so, OpenBLAS) on this box's host cores
               on a bounded sample (one H_eff*phi at a smaller m)
  --impl reference   times the reference CPU build on the same metric (bounded sample per step)

Multi-GPU (--gpus N under torchrun): the output blocks of the chain are sharded by the sector of the
persistent primed link l' (SURVEY §8(e)); each rank runs its shard with no communication until one
NCCL all-gather of the H phi shards per step. Total work is fixed -> "scaling": "strong".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "block-sparse contraction FP64 TFLOP/s (H_eff*phi, S=1/2 Heisenberg Sz blocks)"
WORKLOAD = "H_eff*phi (LocalOp::product: phi*L*W1*W2*R) S=1/2 Heisenberg N=100 centre bond, Sz QDense blocks, maxdim %d"


def workload(m: int, nsect: int, dtype: int):
    from itensor_b200 import synth

    sizes = synth.gaussian_sectors(m, nsect)
    structs = synth.heff_chain(sizes, dtype=dtype)
    return sizes, structs


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe). The timed region of this bench
    is ~15 ms, shorter than one nvidia-smi loop period, so the sampler polls NVML in-process (nvidia_ml_py) every
    millisecond from a thread; nvidia-smi -lms is the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []
        self.samples = []  # (sm_mhz, reasons bitmask)
        self.nvml = None
        self.stop_flag = False
        self.max_mhz = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.device]) if vis and vis.split(",")[self.device].isdigit() else self.device
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                mhz = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((float(mhz), int(rs)))
            except Exception:
                pass
            time.sleep(0.001)

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            n = self.nvml
            names = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            reasons = sorted(k for k, bit in names.items() if any(r & bit for _, r in self.samples))
            sm = [s for s, _ in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml polled every 1 ms during the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def measure_dgemm_peak(torch, dev, n=8192, reps=6):
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n**3 / (best * 1e-3) / 1e12


def cpu_modes(structs, hosts, flops, modes=("A", "B", "C"), reps=1):
    """the reference CPU build on this box's host cores in the three thread modes of SURVEY §8(d), on the given workload"""
    from oracle import orc

    out = {}
    for md in modes:
        if md == "C" and not orc.have_ref_omp():
            continue
        _, desc, threads = orc.ref_mode(md)
        secs, _ = orc.ref_time_heff(structs, hosts, reps, mode=md)
        out[md] = {"tflops": flops / secs / 1e12, "seconds_per_step": secs, "threads": threads, "mode": desc}
    return out


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the SAME step (same maxdim, same sectors, same values),
    every step one full H_eff*phi, on all host cores. Thread counts are set through the libraries' own setters, so the
    OMP_NUM_THREADS=1 that torchrun exports cannot shrink them."""
    if rank != 0:
        return
    from itensor_b200 import ITB_C64, ITB_F64, synth
    import itensor_b200 as itb
    from oracle import orc

    if not orc.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libitref.so not built (needs /root/reference at build time)"}))
        return
    dtype = ITB_C64 if args.complex else ITB_F64
    sizes, structs = workload(args.m, args.nsect, dtype)
    hosts = [synth.random_values(s, 10 + i) for i, s in enumerate(structs)]
    flops, s = 0.0, structs[0]
    for t in structs[1:]:
        p = itb.ContractPlan(s, t)
        flops += p.flops
        s = p.C
    # pick the faster of the two all-core modes on one untimed step each (B: threaded BLAS; C: the reference's OpenMP loop
    # over C blocks, its recommended setting for QN tensors, options.mk.sample:114-150), then time that mode
    probe = cpu_modes(structs, hosts, flops, modes=("B", "C"))
    best = max(probe, key=lambda k: probe[k]["tflops"])
    for _ in range(max(args.warmup - 1, 0)):
        orc.ref_time_heff(structs, hosts, 1, mode=best)
    steps = max(1, min(args.steps, 40))
    t0 = time.perf_counter()
    secs = [orc.ref_time_heff(structs, hosts, 1, mode=best)[0] for _ in range(steps)]
    wall = time.perf_counter() - t0
    per = float(np.mean(secs))
    val = flops / per / 1e12
    cores = orc.host_cores()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "c128" if args.complex else "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD % args.m, "maxdim": args.m, "sectors": sizes, "d": 2, "mpo_link_sectors": [3, 1, 1],
                   "flops_per_step": flops},
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": cores, "kind": "reference",
                         "sample": f"unmodified ITensor CPU build, full H_eff*phi at maxdim {args.m} per step (the same workload, not a "
                                   f"reduced sample), mode {best}: {probe[best]['mode']}; {steps} steps, wall {wall:.1f}s",
                         "modes_probe": probe},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--m", type=int, default=2000, help="maxdim (MPS bond dimension)")
    ap.add_argument("--nsect", type=int, default=9, help="Sz sectors on the MPS links")
    ap.add_argument("--complex", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-refine", action="store_true", help="keep the modelled tile partition (no measured re-partitioning of the plans)")
    ap.add_argument("--refine-rounds", type=int, default=4)
    ap.add_argument("--fused-exchange", action="store_true", help="N > 1: store the owned rows into the peers' buffers from the last contraction's epilogue (itb_contract_run_mirrored) instead of a separate block-copy launch; measured equal or slower (the remote stores stall the consumer warps: 0.760 vs 0.751 ms per step on 2 GPUs), so off by default")
    ap.add_argument("--no-p2p", action="store_true", help="N > 1: re-replicate H*phi with pack / NCCL all-gather / scatter instead of direct peer-memory stores")
    ap.add_argument("--no-rebalance", action="store_true", help="N > 1: keep equal flops per rank (no measured re-balancing of the row partition)")
    ap.add_argument("--rebalance-rounds", type=int, default=3)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    import itensor_b200 as itb
    from itensor_b200 import ITB_C64, ITB_F64, synth
    from itensor_b200._lib import check, lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line (NCCL prints its version banner)
        dist.init_process_group("nccl", device_id=dev)
    ctx = itb.Context(local)
    dtype = ITB_C64 if args.complex else ITB_F64
    sizes, structs = workload(args.m, args.nsect, dtype)
    hosts = [synth.random_values(s, 10 + i) for i, s in enumerate(structs)]

    # ---- plans (host integer work, done once: structure is fixed across Davidson iterations) -----
    plans, s = [], structs[0]
    for t in structs[1:]:
        p = itb.ContractPlan(s, t)
        plans.append(p)
        s = p.C
    total_flops = sum(p.flops for p in plans)
    # multi-GPU: shard the C blocks of every step by the sector of l' (last index of step 1's C, and it
    # stays an uncontracted index of every later intermediate) -> contiguous C-block ranges per rank
    shard = None
    if world > 1:
        from itensor_b200.shard import shard_chain

        shard = shard_chain(plans, world, rank).prepare(ctx.empty)
    max_share = 1.0
    if shard is not None:
        t = torch.tensor([shard.my_flops / shard.total_flops], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        max_share = float(t.item())

    pinned = [torch.from_numpy(np.ascontiguousarray(h).view(np.float64).reshape(-1)).pin_memory() for h in hosts]
    dts = [itb.QTensor(ctx, st, ctx.empty(st.nreal)) for st in structs]
    for d, p in zip(dts, pinned):
        d.data.copy_(p, non_blocking=True)
    outs = [itb.QTensor(ctx, p.C, ctx.empty(p.C.nreal)) for p in plans]
    h_out = torch.empty(plans[-1].C.nreal, dtype=torch.float64).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    p2p_out = None  # N > 1: H*phi in peer-mapped buffers, rows exchanged by direct NVLink stores (set up below)
    step_no = [0]
    fused = [False]  # the row exchange rides the epilogue of the last contraction (itb_contract_run_mirrored)

    def step(b=None):
        if b is None:  # alternate between the two H*phi buffers (see ChainShard.prepare_p2p)
            b = step_no[0] & 1
            step_no[0] += 1
        cur = dts[0]
        for k, p in enumerate(plans):  # sharded: every plan is sliced to this rank's rows of l'
            last = k == len(plans) - 1
            if last and fused[0]:
                # the *R step stores every tile into the peers' buffers from its epilogue; then the arrival barrier
                shard.run_last_mirrored(ctx.handle, p, cur.ptr, dts[k + 1].ptr, b)
                break
            dst = p2p_out[b] if (p2p_out is not None and last) else outs[k]
            check(lib().itb_contract_run(ctx.handle, p._h, cur.ptr, dts[k + 1].ptr, dst.ptr))
            cur = dst
        if shard is not None and not fused[0]:
            if p2p_out is not None:
                shard.push(ctx.handle, b)   # own rows -> every peer's buffer b over NVLink, then the arrival barrier
            else:
                shard.allgather(ctx.handle, outs[-1].data)  # pack own rows -> one NCCL all-gather -> scatter

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Plans that are executed many times refine their static tile partition from measured per-CTA cycles (plan set-up, like
    # the planning itself: itb_contract_plan_refine runs each contraction rounds+1 times on the real operands, in chain order
    # so that every input is valid). --no-refine keeps the modelled partition.
    refine_gain = None

    def refine_plans():
        gains, cur = [], dts[0]
        for k, p in enumerate(plans):
            gains.append(round(p.refine(ctx, cur.ptr, dts[k + 1].ptr, outs[k].ptr, rounds=args.refine_rounds), 4))
            cur = outs[k]
        return gains

    def chain_ms(reps=5):
        # device time of this rank's four contractions (CUDA events, L2 flushed), mean of reps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc = 0.0
        for _ in range(reps):
            flush.zero_()
            e0.record()
            cur = dts[0]
            for k, p in enumerate(plans):
                check(lib().itb_contract_run(ctx.handle, p._h, cur.ptr, dts[k + 1].ptr, outs[k].ptr))
                cur = outs[k]
            e1.record()
            e1.synchronize()
            acc += e0.elapsed_time(e1)
        return acc / reps

    if not args.no_refine:
        refine_gain = refine_plans()
    # N > 1: equal flops per rank is not equal time per rank. A cut inside a sector leaves both sides a remainder tile row
    # that costs almost a full tile per K-chunk, and the all-gather / barrier waits for the slowest rank. The row partition is
    # therefore optimised against the planner's cycle model (ChainShard optimise=True: cuts snap to multiples of the tile
    # height where that pays) and then corrected with measured per-rank factors, one level above the per-CTA refinement;
    # the partition with the shortest slowest rank (measured) is kept.
    rank_balance = None
    if shard is not None and not args.no_rebalance and shard.mode == "rows":
        try:
            def all_times():
                t = torch.tensor([chain_ms()], dtype=torch.float64, device=dev)
                out_t = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(out_t, t)
                return np.array([float(x.item()) for x in out_t])

            def reshard(**kw):
                nonlocal shard, refine_gain
                shard.close()
                shard = shard_chain(plans, world, rank, **kw).prepare(ctx.empty)
                if not args.no_refine:
                    refine_gain = refine_plans()

            history = [(list(shard.cuts), all_times(), None)]          # equal flops
            factors = None
            for _ in range(args.rebalance_rounds):
                reshard(optimise=True, factors=factors)
                times = all_times()
                history.append((list(shard.cuts), times, list(shard.model_ms)))
                f = times / np.asarray(shard.model_ms)
                factors = f / f.mean()
            best = min(range(len(history)), key=lambda i: history[i][1].max())
            if history[best][0] != list(shard.cuts):
                reshard(cuts=history[best][0])
            rank_balance = {"chain_ms_per_rank_equal_flops": [round(x, 4) for x in history[0][1]], "cuts_equal_flops": history[0][0],
                            "chain_ms_per_rank_optimised": [round(x, 4) for x in history[best][1]], "cuts_optimised": history[best][0],
                            "modelled_ms_per_rank": None if history[best][2] is None else [round(x, 4) for x in history[best][2]],
                            "rounds": args.rebalance_rounds, "kept_round": best}
            t = torch.tensor([shard.my_flops / shard.total_flops], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            max_share = float(t.item())
        except Exception as e:  # noqa: BLE001  (keeps the partition it has)
            rank_balance = {"error": str(e)[:200]}
    exchange = "none"
    p2p_check = None
    if shard is not None:
        exchange = "pack -> one NCCL all-gather -> scatter"
        if not args.no_p2p:
            bufs = shard.prepare_p2p(ctx)
            if bufs is not None:
                p2p_out = [itb.QTensor(ctx, plans[-1].C, t) for t in bufs]
                exchange = "direct NVLink stores of the owned rows into every peer's H*phi buffer (CUDA IPC peer memory, block-copy kernel) + flag barrier over peer memory (itb_p2p_barrier); two buffers used alternately"
                if args.fused_exchange:
                    try:   # (every rank plans the same classes for its slice or none does: the decision is all-reduced)
                        fused[0] = True
                        step(0)
                        ok = 1
                    except Exception:  # noqa: BLE001
                        ok = 0
                    fused[0] = False
                    t = torch.tensor([ok], dtype=torch.int32, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MIN)
                    if int(t.item()) == 1:
                        fused[0] = True
                        exchange = "the *R step stores every C tile into the peers' H*phi buffers from its own epilogue (itb_contract_run_mirrored: CUDA IPC peer memory, NVLink stores) + flag barrier over peer memory (itb_p2p_barrier); two buffers used alternately"
                # every rank's assembled H*phi against its own UNSHARDED recomputation of the chain
                for b in (0, 1):
                    p2p_out[b].data.fill_(float("nan"))
                barrier()
                step(0); step(1)
                barrier()
                full, cur_t = [itb.ContractPlan(structs[0], structs[1])], None
                for t in structs[2:]:
                    full.append(itb.ContractPlan(full[-1].C, t))
                tmp = [itb.QTensor(ctx, q.C, ctx.empty(q.C.nreal)) for q in full]
                cur_t = dts[0]
                for k, q in enumerate(full):
                    check(lib().itb_contract_run(ctx.handle, q._h, cur_t.ptr, dts[k + 1].ptr, tmp[k].ptr))
                    cur_t = tmp[k]
                ref_t = tmp[-1].data
                errs = [float(((p2p_out[b].data - ref_t).abs().max() / ref_t.abs().max()).item()) for b in (0, 1)]
                errs = [e if e == e else float("inf") for e in errs]
                t = torch.tensor([max(errs)], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                p2p_check = {"max_rel_err_vs_unsharded_chain_over_ranks": float(t.item()), "bar": 1e-12, "ok": bool(float(t.item()) <= 1e-12)}
                del tmp, full
    for _ in range(args.warmup):
        step()
    barrier()
    # The step is a fixed sequence of launches on fixed buffers (what one Davidson iteration repeats): capture it once in
    # a CUDA graph and replay it in the timed loop, so that the gaps between the 6-10 dependent launches (tile kernel,
    # split-K reduce, row-group kernels on a forked side stream, pack / NCCL all-gather / scatter at N > 1) do not depend
    # on the host. Falls back to eager launches if the capture fails (--no-graph forces that).
    launch_mode = "eager"
    run_step = step
    if not args.no_graph:
        try:
            nbuf = 2 if p2p_out is not None else 1
            graphs = [torch.cuda.CUDAGraph() for _ in range(nbuf)]
            g = graphs[0]
            cap_stream = torch.cuda.Stream(device=dev)
            cap_stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(cap_stream):
                check(lib().itb_ctx_set_stream(ctx.handle, C.c_void_p(cap_stream.cuda_stream)))
                for b in range(nbuf):
                    step(b)  # (warm the capture stream)
                torch.cuda.synchronize()
                for b in range(nbuf):
                    with torch.cuda.graph(graphs[b], stream=cap_stream):
                        step(b)
            check(lib().itb_ctx_set_stream(ctx.handle, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
            torch.cuda.synchronize()
            for gg in graphs:
                gg.replay()
            torch.cuda.synchronize()
            replay_no = [0]

            def run_step():
                graphs[replay_no[0] % nbuf].replay()
                replay_no[0] += 1

            launch_mode = "cuda-graph replay of the captured step"
        except Exception as e:  # noqa: BLE001
            check(lib().itb_ctx_set_stream(ctx.handle, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
            torch.cuda.synchronize()
            launch_mode = "eager (graph capture failed: %s)" % str(e)[:120]
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for e0, e1 in ev:
        flush.zero_()  # flush L2 between timed iterations
        if world > 1:
            dist.barrier()
        e0.record()
        run_step()
        e1.record()
    barrier()
    launches = ctx.launches() - l0
    if run_step is not step:  # a replay does not pass through the C ABI's launch counter: count one captured step
        l1 = ctx.launches()
        step()
        torch.cuda.synchronize()
        launches = (ctx.launches() - l1) * args.steps
    ms = sum(e0.elapsed_time(e1) for e0, e1 in ev) / args.steps
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = total_flops / (ms * 1e-3) / 1e12

    # ---- N > 1: where one step goes (chain of sliced contractions / pack / all-gather / scatter), rank 0's view -------------
    phases = None
    if world > 1:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        acc = np.zeros(4)
        for _ in range(5):
            flush.zero_()
            dist.barrier()
            evs[4].record()
            cur = dts[0]
            for k, p in enumerate(plans):
                if fused[0] and k == len(plans) - 1:
                    shard.run_last_mirrored(ctx.handle, p, cur.ptr, dts[k + 1].ptr, 0, barrier=False)
                    break
                check(lib().itb_contract_run(ctx.handle, p._h, cur.ptr, dts[k + 1].ptr, outs[k].ptr))
                cur = outs[k]
            if fused[0]:
                evs[0].record()
                evs[1].record()
                shard.barrier(ctx.handle)
                evs[2].record()
                torch.cuda.synchronize()
                acc += np.array([evs[4].elapsed_time(evs[0]), 0.0, evs[1].elapsed_time(evs[2]), 0.0])
            elif p2p_out is not None:
                evs[3].record()  # (unused slot)
                shard.push(ctx.handle, 0, evs[:3])
                torch.cuda.synchronize()
                acc += np.array([evs[4].elapsed_time(evs[0]), evs[0].elapsed_time(evs[1]), evs[1].elapsed_time(evs[2]), 0.0])
            else:
                shard.allgather(ctx.handle, outs[-1].data, evs[:4])
                torch.cuda.synchronize()
                acc += np.array([evs[4].elapsed_time(evs[0]), evs[0].elapsed_time(evs[1]), evs[1].elapsed_time(evs[2]), evs[2].elapsed_time(evs[3])])
        acc /= 5
        if p2p_out is not None:
            phases = {"contractions_ms": float(acc[0]), "push_ms": (None if fused[0] else float(acc[1])), "arrival_barrier_ms": float(acc[2]),
                      "exchange": ("inside the last contraction's epilogue" if fused[0] else "separate block-copy launch"),
                      "pushed_bytes_per_rank": int(shard.seg_elems[rank] * (16 if plans[-1].C.is_complex else 8) * (world - 1)),
                      "note": "rank 0, CUDA events, mean of 5 steps (the contractions here write the ordinary output buffer; the barrier includes waiting for the slowest rank)"}
        else:
            phases = {"contractions_ms": float(acc[0]), "pack_ms": float(acc[1]), "allgather_ms": float(acc[2]), "scatter_ms": float(acc[3]),
                      "allgather_bytes_per_rank": int(shard.seg_reals * 8), "note": "rank 0, CUDA events, mean of 5 steps (the all-gather includes waiting for the slowest rank)"}

    # ---- e2e: host buffers in, host buffer out, every step ---------------------------------------
    copy_stream = torch.cuda.Stream(device=dev)
    ev_in = [torch.cuda.Event() for _ in range(5)]

    def e2e_step():
        # what a caller of the reference-facing API pays per product: upload the operands, plan every
        # contraction afresh (host integer work + table upload, as operator* does on every call), run,
        # read H phi back. The operands go up on ONE copy stream in the order the chain consumes them (phi, L, W1, W2, R:
        # one direction of PCIe is a serial resource, so concurrent uploads would only delay the first operands) with an
        # event after each; contraction k waits for operand k+1 only, so steps 1-3 run under the upload of R.
        main = torch.cuda.current_stream(dev)
        copy_stream.wait_stream(main)
        with torch.cuda.stream(copy_stream):
            for d, p, ev in zip(dts, pinned, ev_in):
                d.data.copy_(p, non_blocking=True)
                ev.record(copy_stream)
        if world == 1:
            cur = dts[0]
            main.wait_event(ev_in[0])
            for k in range(4):
                p = itb.ContractPlan(cur.struct, structs[k + 1])
                main.wait_event(ev_in[k + 1])
                check(lib().itb_contract_run(ctx.handle, p._h, cur.ptr, dts[k + 1].ptr, outs[k].ptr))
                cur = itb.QTensor(ctx, p.C, outs[k].data)
                keep.append(p)
        else:
            main.wait_event(ev_in[4])
            step()
        h_out.copy_(outs[-1].data, non_blocking=True)
        torch.cuda.synchronize()
        keep.clear()

    keep = []
    if world > 1:
        # N > 1: every byte crosses PCIe ONCE per step for the whole job. The five operands live back to back in one arena;
        # rank g uploads the g-th 1/N of the arena from pinned host memory, one NCCL all-gather over NVLink completes the
        # arena on every GPU, the sliced chain runs, and every rank reads back only the packed segment of H*phi it owns.
        offs, tot = [], 0
        for pth in pinned:
            offs.append(tot)
            tot += (pth.numel() + 31) // 32 * 32
        chunk = (tot + world - 1) // world
        chunk = (chunk + 31) // 32 * 32
        h_arena = torch.zeros(chunk * world, dtype=torch.float64).pin_memory()
        for o, pth in zip(offs, pinned):
            h_arena[o:o + pth.numel()].copy_(pth)
        d_arena = ctx.empty(chunk * world)
        dts_e = [itb.QTensor(ctx, st, d_arena[o:o + st.nreal]) for o, st in zip(offs, structs)]
        h_seg = torch.empty(shard.seg_reals, dtype=torch.float64).pin_memory()
        mine = slice(rank * chunk, (rank + 1) * chunk)
        e2e_no = [0]

        def e2e_step():  # noqa: F811
            d_arena[mine].copy_(h_arena[mine], non_blocking=True)
            dist.all_gather_into_tensor(d_arena, d_arena[mine])
            cur = dts_e[0]
            b = e2e_no[0] & 1
            e2e_no[0] += 1
            for k, p in enumerate(plans):
                last = k == len(plans) - 1
                if last and fused[0]:
                    shard.run_last_mirrored(ctx.handle, p, cur.ptr, dts_e[k + 1].ptr, b)
                    break
                dst = p2p_out[b] if (p2p_out is not None and last) else outs[k]
                check(lib().itb_contract_run(ctx.handle, p._h, cur.ptr, dts_e[k + 1].ptr, dst.ptr))
                cur = dst
            if p2p_out is not None:
                if not fused[0]:
                    shard.push(ctx.handle, b)
                shard.pack_own(ctx.handle, p2p_out[b].data)
            else:
                shard.allgather(ctx.handle, outs[-1].data)
            h_seg.copy_(shard.send, non_blocking=True)
            torch.cuda.synchronize()
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    h2d = int(sum(p.numel() * 8 for p in pinned))   # whole job: at N > 1 every rank uploads 1/N of it
    d2h = int(h_out.numel() * 8)                    # whole job: at N > 1 every rank reads back the rows it owns

    # ---- roofline of the dominant kernel (128x128 DMMA tile kernel of step 1), live CUDA events ----
    roof = None
    if rank == 0:
        peak = measure_dgemm_peak(torch, dev)
        lib().itb_ctx_set_profile(ctx.handle, 1)
        cls_ms = np.zeros(5)
        cls_fl = np.zeros(5)
        reps = 5
        for _ in range(reps):
            flush.zero_()
            cur = dts[0]
            for k, p in enumerate(plans):
                check(lib().itb_contract_run(ctx.handle, p._h, cur.ptr, dts[k + 1].ptr, outs[k].ptr))
                msv = (C.c_float * 5)()
                lib().itb_contract_last_ms(ctx.handle, msv)
                cls_ms += np.array(list(msv))
                cls_fl += np.array(list(p.info.class_flops))
                cur = outs[k]
        lib().itb_ctx_set_profile(ctx.handle, 0)
        cls_ms /= reps
        cls_fl /= reps
        # timing slots: [0] persistent DMMA tile kernel (+ split-K reduce), [3] streaming kernel, [4] split-K dots
        tile_fl = float(cls_fl[0] + cls_fl[1] + cls_fl[2])
        n_launch = sum(1 for p in plans if p.info.n_gemm_tiles > 0)
        ach = tile_fl / (cls_ms[0] * 1e-3) / 1e12 if cls_ms[0] > 0 else 0.0
        dmma = C.c_double()
        lib().itb_peak_fp64(ctx.handle, 0, 4096, C.byref(dmma))
        dfma = C.c_double()
        lib().itb_peak_fp64(ctx.handle, 1, 4096, C.byref(dfma))
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "r03_traffic.json")
        if os.path.exists(tp) and args.m == 2000 and args.nsect == 9 and not args.complex and world == 1:
            tj = json.load(open(tp))["bsc_gemm_static_kernel"]
            traffic = float(np.mean(tj["per_launch_bytes"]))  # DRAM bytes per launch (mean of the step-1 and step-4 launches)
            traffic_src = "profiles/r03_traffic.json (ncu --set full of this command; algorithmic bytes per launch %.3g)" % float(np.mean(tj["algorithmic_bytes_per_launch"]))
        roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak if peak else None,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": "bsc_gemm_static_kernel (persistent warp-specialised DMMA tiles 128/64/32, stream-K partition)",
                "peak_source": "measured in this run: torch.matmul fp64 8192^3 best of 6 (MEASURED_PEAKS.json has no FP64 entry)",
                "launches_per_step": n_launch, "ms_per_step": {"tile_kernel": float(cls_ms[0]), "streaming_kernel": float(cls_ms[3]), "dot_kernel": float(cls_ms[4])},
                "flops_per_step_by_class": {"tile128": float(cls_fl[0]), "tile64": float(cls_fl[1]), "tile32": float(cls_fl[2]),
                                            "streaming": float(cls_fl[3]), "dot": float(cls_fl[4])},
                "bare_dmma_tflops": dmma.value, "bare_dfma_tflops": dfma.value,
                "hbm_peak_gbs": json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else None,
                "streaming_class_gbs": None}
        # the streaming (MPO) steps are HBM-bound: report their achieved bytes/s too
        sk_bytes = 0.0
        for k, p in enumerate(plans):
            if p.info.class_flops[3] > 0.5 * p.flops:
                sk_bytes += 8.0 * ((structs[0].nreal if k == 0 else plans[k - 1].C.nreal) + p.C.nreal)
        if cls_ms[3] > 0:
            roof["streaming_class_gbs"] = sk_bytes / (cls_ms[3] * 1e-3) / 1e9

    # ---- permute bandwidth (HBM-bound sibling task): reverse the 5 indices of the largest intermediate ----
    perm_info = None
    if rank == 0:
        from itensor_b200.tensor import PermutePlan, permuted_struct

        src = outs[0]  # T1 = phi*L : (s1,s2,r,k0,l')
        D, perm = permuted_struct(src.struct, list(reversed(src.struct.inds)), flux=(0,))
        pp = PermutePlan(src.struct, D, perm)
        # four permutes back to back into alternating destinations inside one event pair: the 254 MB moved per permute
        # exceed the 126 MB L2 (no flush needed between them) and the ~5 us of launch/event latency that a single
        # 50 us kernel would carry is amortised
        dsts = [ctx.empty(D.nreal), ctx.empty(D.nreal)]
        reps_in = 4
        best = 1e9
        for it_ in range(4):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for q in range(reps_in):
                check(lib().itb_permute_run(ctx.handle, pp._h, src.ptr, C.c_void_p(dsts[q & 1].data_ptr()), 1.0, 0.0, 0))
            e1.record(); e1.synchronize()
            if it_ > 0:
                best = min(best, e0.elapsed_time(e1) / reps_in)
        # the permuting ACCUMULATE (PlusEQ with a permutation: what davidson's q += (-Vq)*V[k] hits): dst += P(src),
        # algorithmic bytes = read src + read dst + write dst
        best_acc = 1e9
        for it_ in range(4):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for q in range(reps_in):
                check(lib().itb_permute_run(ctx.handle, pp._h, src.ptr, C.c_void_p(dsts[q & 1].data_ptr()), 0.5, 0.0, 1))
            e1.record(); e1.synchronize()
            if it_ > 0:
                best_acc = min(best_acc, e0.elapsed_time(e1) / reps_in)
        dst = dsts[0]
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        gbs = pp.bytes / (best * 1e-3) / 1e9
        perm_info = {"what": "QDense permute, reverse 5 indices of T1=phi*L (fills all flux-allowed blocks)", "elements": int(src.struct.nelems),
                     "bytes": int(pp.bytes), "ms": best, "achieved_gbs": gbs, "peak_gbs": hbm, "frac": gbs / hbm,
                     "timing": "mean of 4 back-to-back permutes (alternating destinations, 254 MB each > L2) per CUDA-event pair, best of 3",
                     "accumulate": {"what": "same permutation as PlusEQ: dst += 0.5*P(src)", "bytes": int(pp.bytes * 3 // 2), "ms": best_acc,
                                    "achieved_gbs": pp.bytes * 1.5 / (best_acc * 1e-3) / 1e9, "frac": pp.bytes * 1.5 / (best_acc * 1e-3) / 1e9 / hbm}}
        del dst

    # ---- CPU baseline + parity: the reference itself on this box's host cores, on the SAME tensors ---------------------
    cpu, parity = None, None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import orc

        if orc.have_ref():
            t0 = time.perf_counter()
            modes = cpu_modes(structs, hosts, total_flops)
            best = max(modes, key=lambda k: modes[k]["tflops"])
            cpu = {"value": modes[best]["tflops"], "unit": "TFLOP/s", "cores": orc.host_cores(), "kind": "reference",
                   "modes": modes,
                   "sample": f"unmodified ITensor CPU build, one full H_eff*phi of this workload (maxdim {args.m}, {total_flops:.3g} flop) per "
                             f"thread mode A/B/C of SURVEY 8(d); value = fastest mode ({best}); {time.perf_counter() - t0:.1f}s wall"}
            # element-level parity of THIS workload against the reference (north_star: structure bit-exact, values 1e-12
            # relative): every one of the four contractions, each fed with the reference's own previous intermediate
            if world == 1:
                parity = {"tolerance": 1e-12, "steps": []}
                cur_s, cur_h = structs[0], hosts[0]
                for k, p in enumerate(plans):
                    rr = orc.ref_contract(cur_s, cur_h, structs[k + 1], hosts[k + 1])
                    d_in = itb.QTensor.from_host(ctx, cur_s, cur_h)
                    check(lib().itb_contract_run(ctx.handle, p._h, d_in.ptr, dts[k + 1].ptr, outs[k].ptr))
                    got = outs[k].data.cpu().numpy()
                    got = got.view(np.complex128) if p.C.is_complex else got
                    same = bool(np.array_equal(rr.blocks, p.C.blocks) and np.array_equal(rr.offsets, p.C.offsets)
                                and rr.nelems == p.C.nelems and list(rr.labels) == [int(x) for x in p.C.labels])
                    err = float(np.abs(got - rr.data).max() / np.abs(rr.data).max())
                    parity["steps"].append({"step": k + 1, "structure_bit_exact": same, "max_rel_err": err, "elements": int(rr.nelems)})
                    cur_s, cur_h = p.C, rr.data
                parity["ok"] = all(st["structure_bit_exact"] and st["max_rel_err"] <= 1e-12 for st in parity["steps"])
        else:
            cpu = {"value": None, "unit": "TFLOP/s", "cores": os.cpu_count(), "kind": "reference", "sample": "oracle/_ref not built"}

    # ---- the same step through the C++ plugin: ITensor::operator* on QDenseGPU storage (separate process, GPU idle here) ----
    plugin = None
    hb = os.path.join(ROOT, "build", "plugin", "heff_bench")
    if rank == 0 and world == 1 and os.path.exists(hb) and not args.complex:
        import sysconfig

        env = dict(os.environ, ITB_WARM_LIBS="0", LD_LIBRARY_PATH=os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs")
                   + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
        try:
            torch.cuda.synchronize()
            out = subprocess.run([hb, str(args.m), str(args.nsect), str(max(args.steps, 3)), "3"], env=env, capture_output=True, text=True, timeout=600)
            plugin = json.loads(out.stdout.strip().split("\n")[-1])
            plugin["resident_tflops"] = total_flops / (plugin["resident_ms_per_step"] * 1e-3) / 1e12
            plugin["e2e_tflops"] = total_flops / (plugin["e2e_ms_per_step"] * 1e-3) / 1e12
        except Exception as e:  # the Python-mirror numbers above stand on their own
            plugin = {"error": str(e)[:200]}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c128" if args.complex else "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD % args.m,
                       "maxdim": args.m, "sectors": sizes, "d": 2, "mpo_link_sectors": [3, 1, 1],
                       "pairs_per_step": [int(p.npairs) for p in plans], "flops_per_step": total_flops,
                       "l2": "flushed between timed iterations (256 MiB memset)", "launch": launch_mode,
                       "tile_partition": ("modelled" if refine_gain is None else "refined from measured per-CTA cycles at plan set-up (itb_contract_plan_refine, %d rounds; longest-CTA gain per plan %s)" % (args.refine_rounds, refine_gain)),
                       "sharding": ("rows of l' (%s), contiguous row ranges per rank (equal flops, cuts then moved by the cycle model + measured rank times), max rank share %.3f of flops (ideal %.3f); "
                                    "H*phi re-replicated by %s; e2e: 1/N of the operand arena per "
                                    "rank over PCIe + NCCL all-gather" % (shard.mode, max_share, 1.0 / world, exchange)) if world > 1 else "none"},
            "e2e": {"value": total_flops / (e2e_ms * 1e-3) / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "parity_vs_reference": parity, "e2e_plugin": plugin, "multi_gpu_phases": phases, "multi_gpu_rank_balance": rank_balance, "multi_gpu_exchange_check": p2p_check,
            "permute": perm_info,
        }))
    if world > 1 and p2p_out is not None:
        ep, berr = shard.barrier_status()
        if berr:
            print(json.dumps({"error": "rank %d: peer-memory barrier wait expired at epoch %d" % (rank, berr)}), file=sys.stderr, flush=True)
    if world > 1:
        # Tear-down must never outlive the measurement: a captured graph that holds NCCL kernels keeps the communicator
        # busy inside destroy_process_group (seen on 2 GPUs: the JSON line was out, the processes sat in ncclCommDestroy until
        # the launcher's timeout). Release the graph first, and leave through os._exit if the orderly shutdown stalls.
        import threading

        sys.stdout.flush()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        run_step = None
        for gg in locals().get("graphs", []):
            try:
                gg.reset()
            except Exception:  # noqa: BLE001
                pass
        torch.cuda.synchronize()
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass
        os._exit(0)


if __name__ == "__main__":
    main()
