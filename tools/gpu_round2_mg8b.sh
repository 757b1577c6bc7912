#!/bin/bash
# 8-GPU pass: the driver's command line (config 4 + the config-5 entry), then the same with MCL_PROFILE stage timing
tag=${1:-r02aa}
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_config4_8gpu.json 2> gpurun_out/${tag}_bench_8gpu.err
tail -2 gpurun_out/${tag}_bench_8gpu.err
if [ -n "$MCL_WANT_PROFILE" ]; then
MCL_PROFILE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29619 bench.py --gpus 8 --steps 10 --warmup 3 --no-extra > gpurun_out/${tag}_prof_8gpu.json 2> gpurun_out/${tag}_prof_8gpu.err
grep -A20 "MCL_PROFILE rank 0" gpurun_out/${tag}_prof_8gpu.err | head -22
fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 4 --steps 20 --warmup 5 --no-extra > gpurun_out/${tag}_bench_config4_4gpu.json 2> gpurun_out/${tag}_bench_4gpu.err
for f in gpurun_out/${tag}_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], "value %.4e ms %.3f e2e %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d.get("stage_ms"), d.get("estimate"), d["digest"]["scores"], d["digest"]["poses"])
    for c in d.get("configs", []): print("   ", c["config"]["workload"][:40], "ms %.4f e2e %.4f"%(c["ms_per_step"], c["e2e"]["ms_per_step"]), c["stage_ms"], c.get("estimate"), c["details"]["map_tile_used"], c["digest"]["scores"])
except Exception as ex: print(sys.argv[1], "ERR", ex)
PY
done
