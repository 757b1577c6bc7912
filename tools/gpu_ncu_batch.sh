#!/bin/bash
tag=${1:-r02s}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_table_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_score_table_batch8_config5 \
    python bench.py --config config5 --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/${tag}_ncu5.log 2>&1
tail -3 gpurun_out/${tag}_ncu5.log
