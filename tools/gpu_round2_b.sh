#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -15 gpurun_out/r02d_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02d_bench_config4.json 2> gpurun_out/r02d_bench_config4.err
timeout 300 python bench.py --config config3 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02d_bench_config3.json 2>&1
timeout 300 python bench.py --config config2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r02d_bench_config2.json 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_table_kernel -s 3 -c 1 -f -o gpurun_out/r02d_score_table_kernel_config4 \
    python bench.py --config config4 --steps 1 --warmup 3 --no-cpu > gpurun_out/r02d_ncu.log 2>&1
for f in gpurun_out/r02d_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], "value %.4e ms %.3f"%(d["value"], d["ms_per_step"]), d.get("stage_ms"), d.get("config",{}).get("deferred_fraction"), d["config"].get("sensor_path"))
except Exception as ex: print(sys.argv[1], "ERR", ex)
PY
done
