#!/bin/bash
# ncu --set full capture of the sensor kernel for one or more prebuilt variants (config3), reports into gpurun_out/.
cfg=${CFG:-config3}
for name in "$@"; do
  so=$PWD/botlab_b200/variants/libmcl_$name.so
  [ "$name" = product ] && so=$PWD/botlab_b200/libmcl_cuda.so
  MCL_LIB=$so ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 3 -c 1 -f \
     -o gpurun_out/score_${name}_$cfg python bench.py --config $cfg --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_${name}_$cfg.log 2>&1
  tail -2 gpurun_out/ncu_${name}_$cfg.log | cut -c1-200
done
