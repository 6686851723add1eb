#!/bin/bash
# bench.py under torchrun on N GPUs (the driver's launch line)
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/q_bench_n$N.json 2> $OUT/q_bench_n$N.err; echo "bench N=$N rc=$?"
tail -2 $OUT/q_bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/q_bench_n$N.json").read().strip().split("\n")[-1])
print("N=%d value %.2f TFLOP/s ms %.3f e2e %.2f (%.3f ms) phases %s"%(d["n_gpus"],d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"],d["multi_gpu_phases"]))
print(d["config"]["sharding"])
PY
