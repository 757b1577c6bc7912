for so in botlab_b200/variants/libmcl_*.so; do
  name=$(basename $so .so); name=${name#libmcl_}
  MCL_LIB=$PWD/$so python bench.py --config config4 --steps 4 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.load(sys.stdin); print('$name score_ms %.3f step_ms %.3f'%(d['stage_ms']['score'], d['ms_per_step']))
except Exception as e: print('$name FAILED', e)
"
done
