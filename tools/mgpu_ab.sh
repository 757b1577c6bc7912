#!/bin/bash
# A/B of the pose exchange at N GPUs: copy-engine peer push (default) vs NCCL all-gather (MCL_NO_PEER_PUSH=1)
N=${1:-2}
for mode in push nccl; do
  if [ $mode = nccl ]; then export MCL_NO_PEER_PUSH=1; else unset MCL_NO_PEER_PUSH; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$N bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_ab_${mode}_g$N.json
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab_${mode}_g$N.json")); print("$mode", $N, "value %.3e ms %.3f e2e_ms %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d["stage_ms"])
PY
done
