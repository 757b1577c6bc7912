#!/bin/bash
tag=${1:-r02m}
mkdir -p gpurun_out
MCL_DEBUG_TABLE=1 timeout 300 python bench.py --config config5 --steps 2 --warmup 2 --no-cpu --no-extra 2> gpurun_out/${tag}_dbg.err | tail -1 > gpurun_out/${tag}_bench_config5_64M_1gpu.json
head -40 gpurun_out/${tag}_dbg.err
MCL_DEBUG_TABLE=1 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "batch_windows" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
grep -v "^\[mcl table\] hint" gpurun_out/${tag}_pytest.log | tail -40
