#!/bin/bash
tag=${1:-r02v}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batch_windows or config5" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
MCL_DEBUG_TABLE=1 timeout 300 python bench.py --config config5 --steps 3 --warmup 3 --no-cpu --no-extra 2> gpurun_out/${tag}_dbg5.err | tail -1 > gpurun_out/${tag}_bench_config5_64M_1gpu.json
grep "plan" gpurun_out/${tag}_dbg5.err | sort | uniq -c | sort -rn | head -5 | cut -c1-200
for f in gpurun_out/${tag}_bench_*.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], "value %.4e ms %.3f e2e %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d.get("stage_ms"), d["details"]["deferred_fraction"], d["details"]["sensor_path"], d["details"]["map_tile_used"])
PY
done
