#!/bin/bash
# round-2 session U (2 GPUs): peer-memory row exchange — test, then bench under torchrun with and without it
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
if [ "$N" = "2" ]; then
timeout 300 python -m pytest tests/test_p2p_gpu.py -m gpu -x -q > $OUT/u_pytest_p2p.log 2>&1; echo "p2p test rc=$?"; tail -15 $OUT/u_pytest_p2p.log
fi
for MODE in ${MODES:-p2p nccl}; do
  FLAG=""; [ "$MODE" = "nccl" ] && FLAG="--no-p2p"
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 $FLAG > $OUT/u_bench_n${N}_$MODE.json 2> $OUT/u_bench_n${N}_$MODE.err; echo "bench N=$N $MODE rc=$?"
  tail -2 $OUT/u_bench_n${N}_$MODE.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/u_bench_n${N}_$MODE.json").read().strip().split("\n")[-1])
    print("N=%d $MODE value %.2f TFLOP/s ms %.3f e2e %.2f (%.3f ms) phases %s"%(d["n_gpus"],d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"],d["multi_gpu_phases"]))
    print("   balance", d.get("multi_gpu_rank_balance"), "check", d.get("multi_gpu_exchange_check"))
except Exception as e: print("parse failed", e)
PY
done
