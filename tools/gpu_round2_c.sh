#!/bin/bash
tag=${1:-r02f}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -8 gpurun_out/${tag}_pytest.log
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
tail -4 gpurun_out/${tag}_bench_default.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_table_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_score_table_kernel_config4 \
    python bench.py --config config4 --steps 1 --warmup 3 --no-cpu --no-extra > gpurun_out/${tag}_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/${tag}_launches_config4.csv \
    python bench.py --config config4 --steps 2 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
python - gpurun_out/${tag}_bench_default.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4e ms %.3f e2e %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d["stage_ms"], d["details"]["deferred_fraction"], d["details"]["sensor_path"])
print("roofline", {k:v for k,v in d["roofline"].items() if k in ("bound","achieved","peak","frac","kernel_ms")}, d["roofline"].get("smem"))
print("cpu", d.get("cpu_baseline"))
for c in d.get("configs", []): print(c["config"]["workload"][:40], "ms %.4f e2e %.4f"%(c["ms_per_step"], c["e2e"]["ms_per_step"]), c["stage_ms"], c["details"]["sensor_path"], c["gpu_launches"], c.get("cpu_baseline",{}).get("value"))
PY
