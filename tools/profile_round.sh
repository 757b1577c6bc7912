#!/bin/bash
# One-GPU profiling pass for profiles/: launch list + full captures of the two sensor kernels + bench JSON lines.
tag=${1:-r01_v6}
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 80 --csv --log-file gpurun_out/${tag}_launches_config4.csv \
    python bench.py --config config4 --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
for k in score_fast_kernel score_deferred_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/${tag}_${k}_config4 \
      python bench.py --config config4 --steps 1 --warmup 3 --no-cpu > gpurun_out/${tag}_${k}.log 2>&1
done
python bench.py --steps 10 --warmup 3 | tail -1 > gpurun_out/${tag}_bench_config4_1gpu.json
python bench.py --config config3 --steps 10 --warmup 3 | tail -1 > gpurun_out/${tag}_bench_config3_1gpu.json
python bench.py --config config2 --steps 20 --warmup 3 | tail -1 > gpurun_out/${tag}_bench_config2_1gpu.json
python bench.py --config config5 --particles 8000000 --steps 5 --warmup 3 --no-cpu | tail -1 > gpurun_out/${tag}_bench_config5shape_8M_uniform_1gpu.json
python bench.py --sensor-path 1 --steps 5 --warmup 3 --no-cpu | tail -1 > gpurun_out/${tag}_bench_config4_1gpu_exact_only.json
python bench.py --impl reference --steps 2 --warmup 1 | tail -1 > gpurun_out/${tag}_bench_config4_reference.json
for f in gpurun_out/${tag}_bench_*.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], "value %.4e ms %.3f"%(d["value"], d["ms_per_step"]), d.get("stage_ms"), d.get("config",{}).get("deferred_fraction"))
PY
done
