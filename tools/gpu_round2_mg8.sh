#!/bin/bash
tag=${1:-r02h}
mkdir -p gpurun_out
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_bench_config4_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
  tail -2 gpurun_out/${tag}_bench_${n}gpu.err
done
for f in gpurun_out/${tag}_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], "value %.4e ms %.3f e2e %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d.get("stage_ms"), d.get("estimate"), d["digest"]["weights"], d["digest"]["poses"])
    for c in d.get("configs", []): print("   ", c["config"]["workload"][:40], "ms %.4f e2e %.4f"%(c["ms_per_step"], c["e2e"]["ms_per_step"]), c["stage_ms"], c.get("estimate"))
except Exception as ex: print(sys.argv[1], "ERR", ex)
PY
done
