#!/bin/bash
# In-step A/B of the row-partition schedules / transports at N GPUs on the full 2 B-edge graph (config P).
# usage: tools/n8_ab.sh N  ->  gpurun_out/r2_ab_n${N}_<variant>.json (SIRGCN_BENCH_VALUE_ONLY lines: value leg only)
N=${1:-8}
port=29600
run() {
  name=$1; shift
  port=$((port+1))
  SIRGCN_BENCH_VALUE_ONLY=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/r2_ab_n${N}_$name.json 2> gpurun_out/r2_ab_n${N}_$name.err
  echo "$name rc=$? $(head -c 300 gpurun_out/r2_ab_n${N}_$name.json)"
}
run coll_c4_b1 --transport collective --chunks 4 --bwd-chunks 1
run coll_c4_b4 --transport collective --chunks 4 --bwd-chunks 4
run coll_c8_b4 --transport collective --chunks 8 --bwd-chunks 4
run tma_c4_b4 --transport pushtma --chunks 4 --bwd-chunks 4
run sm_c4_b4 --transport pushsm --chunks 4 --bwd-chunks 4
