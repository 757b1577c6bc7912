#!/bin/bash
# Rebuilds the CUDA library with different tuning macros on the GPU box and benches config 3 for each.
for v in "-DMCL_BEAM_UNROLL=1" "-DMCL_BEAM_UNROLL=2" "-DMCL_BEAM_UNROLL=4" "-DMCL_BEAM_UNROLL=2 -DMCL_SCORE_MIN_CTAS=3"; do
  touch botlab_b200/csrc/mcl_engine.cu
  make -s -C botlab_b200/csrc EXTRA="$v" > /dev/null 2>&1
  regs=$(grep -A2 "score_kernelILi1ELb1ELb1ELb0" botlab_b200/csrc/build.log | grep -o "Used [0-9]* registers")
  python bench.py --config config3 --steps 5 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('$v', '$regs', 'score_ms %.3f'%d['stage_ms']['score'], 'value %.3e'%d['value'])"
done
touch botlab_b200/csrc/mcl_engine.cu; make -s -C botlab_b200/csrc > /dev/null 2>&1
for l in 1 2 4; do python bench.py --config config3 --steps 5 --no-cpu --lanes $l 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('lanes $l score_ms %.3f'%d['stage_ms']['score'], 'value %.3e'%d['value'])"; done
