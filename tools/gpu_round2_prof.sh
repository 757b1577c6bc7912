#!/bin/bash
tag=${1:-r02i}
N=${2:-2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -6 gpurun_out/${tag}_pytest.log
MCL_PROFILE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/${tag}_prof_1gpu.json 2> gpurun_out/${tag}_prof_1gpu.err
grep -A40 "MCL_PROFILE" gpurun_out/${tag}_prof_1gpu.err | head -30
MCL_PROFILE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --steps 10 --warmup 3 --no-extra > gpurun_out/${tag}_prof_${N}gpu.json 2> gpurun_out/${tag}_prof_${N}gpu.err
grep -A40 "MCL_PROFILE" gpurun_out/${tag}_prof_${N}gpu.err | head -30
