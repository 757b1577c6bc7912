#!/bin/bash
# full GPU suite + the three sensor-stage shapes (config 4 seeded / interior pose, config 5 on one GPU)
tag=${1:-r02q}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -12 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --config config5 --steps 3 --warmup 3 --no-cpu --no-extra | tail -1 > gpurun_out/${tag}_bench_config5_64M_1gpu.json
timeout 300 python bench.py --pose interior --steps 10 --warmup 3 --no-cpu --no-extra | tail -1 > gpurun_out/${tag}_bench_config4_interior_1gpu.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra | tail -1 > gpurun_out/${tag}_bench_config4_1gpu.json
for f in gpurun_out/${tag}_bench_*.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], "value %.4e ms %.3f e2e %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d.get("stage_ms"), d["details"]["deferred_fraction"], d["details"]["sensor_path"], d["details"]["map_tile_used"])
PY
done
