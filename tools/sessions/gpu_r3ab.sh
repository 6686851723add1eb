#!/bin/bash
# round-2 session AB (N GPUs): 16-byte push items vs 8-byte
N=${1:-4}
OUT=gpurun_out; mkdir -p $OUT
for MODE in wide narrow; do
  FLAG=""; [ "$MODE" = "narrow" ] && FLAG="--no-wide-push"
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 $FLAG > $OUT/ab_bench_n${N}_$MODE.json 2> $OUT/ab_bench_n${N}_$MODE.err; echo "bench N=$N $MODE rc=$?"
  tail -1 $OUT/ab_bench_n${N}_$MODE.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_bench_n${N}_$MODE.json").read().strip().split("\n")[-1])
    print("N=%d $MODE value %.2f TFLOP/s ms %.3f e2e %.2f phases %s check %s"%(d["n_gpus"],d["value"],d["ms_per_step"],d["e2e"]["value"],d["multi_gpu_phases"],d.get("multi_gpu_exchange_check")))
except Exception as e: print("parse failed", e)
PY
done
