#!/bin/bash
# one 8-GPU box session: transport microbenchmark, then bench.py with the fastest transport (chunks 1 and 4)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 150 $TR tools/peer_bw.py > gpurun_out/peer_bw_n8.json 2> gpurun_out/peer_bw_n8.err; echo "bw rc=$?"
cat gpurun_out/peer_bw_n8.json
read -r T CTAS < <(python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/peer_bw_n8.json").read().strip().splitlines()[-1])
    cand = {k: v["gbs_in_per_rank"] for k, v in j.items() if isinstance(v, dict)}
    best = max(cand, key=cand.get)
    if best.startswith("sm_push"):
        print("pushsm", best.split("_")[2].replace("ctas", ""))
    else:
        print({"nccl_all_gather": "collective", "ce_pull": "peer", "ce_push": "push"}[best], 32)
except Exception:
    print("collective", 32)
PY
)
echo "picked transport=$T ctas=$CTAS"
for c in 1 4; do
  SIRGCN_PUSH_CTAS=$CTAS timeout 200 $TR bench.py --gpus 8 --steps 5 --warmup 3 --chunks $c --transport $T --no-cpu-baseline > gpurun_out/bench_n8_${T}_c$c.json 2> gpurun_out/bench_n8_${T}_c$c.err; echo "bench $T chunks=$c rc=$?"
done
