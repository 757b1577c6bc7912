#!/bin/bash
# multi-GPU pass: parity across GPU counts + bench at N ranks
N=${1:-2}
tag=${2:-r02e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py tests/test_gpu_parity.py::test_normalize_and_estimate_golden tests/test_gpu_parity.py::test_lse_weight_mode -m gpu -x -q > gpurun_out/${tag}_pytest_mg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_mg.log
tail -8 gpurun_out/${tag}_pytest_mg.log
for n in 1 $N; do
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_config4_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_bench_config4_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
  fi
  tail -3 gpurun_out/${tag}_bench_${n}gpu.err
done
for f in gpurun_out/${tag}_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], "value %.4e ms %.3f e2e %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d.get("stage_ms"), d["config"].get("sensor_path"), d.get("estimate"))
except Exception as ex: print(sys.argv[1], "ERR", ex)
PY
done
