#!/bin/bash
# A/B of prebuilt library variants (botlab_b200/variants/libmcl_<name>.so) on config 4, seeded and interior pose
tag=${1:-r02ai}; shift
mkdir -p gpurun_out
for name in product "$@"; do
  so=$PWD/botlab_b200/variants/libmcl_$name.so
  [ "$name" = product ] && so=$PWD/botlab_b200/libmcl_cuda.so
  for pose in seeded interior; do
    MCL_LIB=$so timeout 300 python bench.py --pose $pose --steps 10 --warmup 3 --no-cpu --no-extra | tail -1 > gpurun_out/${tag}_ab_${name}_${pose}.json
    python - gpurun_out/${tag}_ab_${name}_${pose}.json $name $pose <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[2], sys.argv[3], "ms %.3f score %.3f"%(d["ms_per_step"], d["stage_ms"]["score"]), d["digest"]["scores"])
PY
  done
done
