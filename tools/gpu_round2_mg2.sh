#!/bin/bash
tag=${1:-r02x}
N=${2:-2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/${tag}_pytest_mg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_mg.log
tail -5 gpurun_out/${tag}_pytest_mg.log
MCL_PROFILE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --steps 10 --warmup 3 --no-extra > gpurun_out/${tag}_bench_config4_${N}gpu.json 2> gpurun_out/${tag}_bench_config4_${N}gpu.err
grep -A24 "MCL_PROFILE rank 0" gpurun_out/${tag}_bench_config4_${N}gpu.err | head -26
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus $N --config config5 --steps 3 --warmup 3 --no-extra > gpurun_out/${tag}_bench_config5_${N}gpu.json 2> gpurun_out/${tag}_bench_config5_${N}gpu.err
for f in gpurun_out/${tag}_bench_*gpu.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1].split('/')[-1], "value %.4e ms %.3f e2e %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d.get("stage_ms"), d["details"]["sensor_path"], d["details"]["map_tile_used"], d["digest"]["scores"])
PY
done
