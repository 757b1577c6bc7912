#!/bin/bash
# what the driver runs at round end, on one GPU: the -m gpu suite, smoke(), the default bench line, the reference arm
tag=${1:-r02ac}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 900 python bench.py ) > gpurun_out/${tag}_bench_default_1gpu.json 2> gpurun_out/${tag}_bench_default_1gpu.err
tail -4 gpurun_out/${tag}_bench_default_1gpu.err
python - gpurun_out/${tag}_bench_default_1gpu.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4e ms %.3f e2e %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d["stage_ms"], d["details"]["deferred_fraction"], d["details"]["sensor_path"], d["roofline"]["frac"], d["roofline"]["kernel"])
print("cpu", d.get("cpu_baseline"))
for c in d.get("configs", []): print(c["config"]["workload"][:40], "ms %.4f e2e %.4f"%(c["ms_per_step"], c["e2e"]["ms_per_step"]), c["details"]["sensor_path"], c["gpu_launches"], c.get("cpu_baseline",{}).get("value"))
PY
