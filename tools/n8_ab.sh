#!/bin/bash
# In-step A/B of the row-partition schedules / transports at N GPUs on the full 2 B-edge graph (config P).
# usage: tools/n8_ab.sh N  ->  gpurun_out/r2_ab_n${N}_<variant>.json (SIRGCN_BENCH_VALUE_ONLY lines: value leg only)
#
# GPU time at N GPUs is charged N-fold: the script FAILS FAST.  A 2 % scale canary (parity gate + 2 steps, 150 s limit)
# must pass before any full-size run starts, every run has its own limit, and the first failure stops the script.
# (Round 2 lost its remaining budget to four hung full-size runs: the parity gate of that build let rank 0 exit alone.)
N=${1:-8}
port=29600
run() {
  limit=$1; name=$2; shift 2
  port=$((port+1))
  SIRGCN_BENCH_VALUE_ONLY=1 timeout $limit python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus $N --warmup 3 "$@" > gpurun_out/r2_ab_n${N}_$name.json 2> gpurun_out/r2_ab_n${N}_$name.err
  rc=$?
  echo "$name rc=$rc $(head -c 300 gpurun_out/r2_ab_n${N}_$name.json)"
  if [ $rc -ne 0 ]; then
    grep -v "^\[W\|^$" gpurun_out/r2_ab_n${N}_$name.err | tail -5
    echo "stopping: $name failed"
    exit $rc
  fi
}
run 150 canary_s002 --scale 0.02 --steps 2 --transport collective --chunks 4
run 240 coll_c4_b1 --steps 5 --transport collective --chunks 4 --bwd-chunks 1
run 240 coll_c4_b4 --steps 5 --transport collective --chunks 4 --bwd-chunks 4
run 240 coll_c8_b4 --steps 5 --transport collective --chunks 8 --bwd-chunks 4
run 240 tma_c4_b4 --steps 5 --transport pushtma --chunks 4 --bwd-chunks 4
run 240 sm_c4_b4 --steps 5 --transport pushsm --chunks 4 --bwd-chunks 4
