#!/bin/bash
# compute-sanitizer over the score-table kernel's variants, the class-map / window build, the map-following paths and the
# sharded exact sum; + the lanes-per-particle A/B of the exact kernel (north-star mapping)
tag=${1:-r02ab}
mkdir -p gpurun_out
K='sensor_golden or scoring_follows or likelihood_field_mode or (batch_windows and 600) or hostile or resample_golden or normalize'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > gpurun_out/${tag}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/${tag}_sanitizer_memcheck.log
tail -4 gpurun_out/${tag}_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > gpurun_out/${tag}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/${tag}_sanitizer_racecheck.log
tail -4 gpurun_out/${tag}_sanitizer_racecheck.log
for lanes in 1 32; do
  timeout 300 python bench.py --config config3 --sensor-path 1 --lanes $lanes --steps 5 --warmup 3 --no-cpu --no-extra | tail -1 > gpurun_out/${tag}_bench_config3_exact_lanes${lanes}.json
done
for f in gpurun_out/${tag}_bench_*.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], "value %.4e ms %.3f"%(d["value"], d["ms_per_step"]), d["stage_ms"]["score"], d["details"]["lanes_per_particle"], d["details"]["sensor_path"])
PY
done
