#!/bin/bash
# Benches every prebuilt variant in botlab_b200/variants/ (see tools/build_variants.sh) on the GPU box.
#   tools/sweep_variants.sh "config4 config3" [name-glob]
cfgs=${1:-config4}
glob=${2:-*}
for so in botlab_b200/variants/libmcl_$glob.so; do
  name=$(basename $so .so); name=${name#libmcl_}
  for cfg in $cfgs; do
    MCL_LIB=$PWD/$so python bench.py --config $cfg --steps 4 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.load(sys.stdin); print('$name $cfg score_ms %.3f step_ms %.3f value %.4e tile %s'%(d['stage_ms']['score'], d['ms_per_step'], d['value'], d['config']['map_tile_used']))
except Exception as e: print('$name $cfg FAILED', e)
"
  done
done
