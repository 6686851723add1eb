#!/bin/bash
# One process per GPU for the plugin binaries: tools/run_ranks.sh <N> <command...>
# Rank r gets ITB_WORLD=N ITB_RANK=r ITB_DEVICE=r and a common ITB_COMM_FILE (the NCCL id is exchanged through it,
# plugin/gpu_storage.cc comm()). Rank 0's stdout/stderr are passed through, the others go to $RANK_LOG_DIR (default /tmp).
N=$1; shift
ID=$(mktemp -u /tmp/itb_comm_XXXXXX)
LOGS=${RANK_LOG_DIR:-/tmp}
pids=()
for r in $(seq 1 $((N-1))); do
  ITB_WORLD=$N ITB_RANK=$r ITB_DEVICE=$r ITB_COMM_FILE=$ID "$@" > $LOGS/rank$r.out 2> $LOGS/rank$r.err &
  pids+=($!)
done
ITB_WORLD=$N ITB_RANK=0 ITB_DEVICE=0 ITB_COMM_FILE=$ID "$@"
rc=$?
for p in "${pids[@]}"; do wait $p || rc=$?; done
rm -f $ID
exit $rc
