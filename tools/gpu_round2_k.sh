#!/bin/bash
# batch-mode table kernel: focused tests, then config-5 shape and config-4 bench lines
tag=${1:-r02k}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batch_windows or config5 or sensor_golden or init_uniform or two_pass or randomized" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --config config5 --steps 5 --warmup 3 --no-cpu --no-extra | tail -1 > gpurun_out/${tag}_bench_config5_64M_1gpu.json
MCL_NO_TABLE_BATCH=1 timeout 300 python bench.py --config config5 --steps 3 --warmup 3 --no-cpu --no-extra | tail -1 > gpurun_out/${tag}_bench_config5_64M_1gpu_nobatch.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra | tail -1 > gpurun_out/${tag}_bench_config4_1gpu.json
for f in gpurun_out/${tag}_bench_*.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], "value %.4e ms %.3f"%(d["value"], d["ms_per_step"]), d.get("stage_ms"), d["details"]["deferred_fraction"], d["details"]["sensor_path"], d["details"]["map_tile_used"])
PY
done
