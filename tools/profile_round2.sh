#!/bin/bash
# One-GPU profiling pass for profiles/ (round 2): full ncu captures of the score-table kernel's three measured variants,
# the launch list of two config-4 steps, and the bench JSON lines.  Run under gpurun; summarise here with
# tools/ncu_summary.py and tools/kernel_figures.py.
tag=${1:-r02w}
mkdir -p gpurun_out
B="python bench.py --warmup 3 --no-cpu --no-extra"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_table_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_score_table_v0_config4 \
    $B --config config4 --steps 1 > gpurun_out/${tag}_ncu_v0.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_table_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_score_table_v1_config4_interior \
    $B --config config4 --pose interior --steps 1 > gpurun_out/${tag}_ncu_v1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_table_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_score_table_v3_config5 \
    $B --config config5 --steps 2 > gpurun_out/${tag}_ncu_v3.log 2>&1
# summaries on the box (the raw reports are 20 MB each; gpurun brings back at most 64 MB): keep the config-4 report only
E4=$((16000000*357)); E5=$((64000000*357))
python tools/ncu_summary.py gpurun_out/${tag}_score_table_v0_config4.ncu-rep $E4 > gpurun_out/${tag}_score_table_v0_config4.txt 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_score_table_v1_config4_interior.ncu-rep $E4 > gpurun_out/${tag}_score_table_v1_config4_interior.txt 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_score_table_v3_config5.ncu-rep $E5 > gpurun_out/${tag}_score_table_v3_config5.txt 2>&1
rm -f gpurun_out/${tag}_score_table_v1_config4_interior.ncu-rep gpurun_out/${tag}_score_table_v3_config5.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/${tag}_launches_config4.csv \
    $B --config config4 --steps 2 > /dev/null 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 | tail -1 > gpurun_out/${tag}_bench_default_1gpu.json
timeout 300 python bench.py --pose interior --steps 10 --warmup 3 --no-cpu --no-extra | tail -1 > gpurun_out/${tag}_bench_config4_interior_1gpu.json
timeout 300 python bench.py --config config5 --steps 3 --warmup 3 --no-cpu --no-extra | tail -1 > gpurun_out/${tag}_bench_config5_64M_1gpu.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | tail -1 > gpurun_out/${tag}_bench_config4_reference.json
for f in gpurun_out/${tag}_bench_*.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1].split('/')[-1], "value %.4e ms %.3f"%(d["value"], d["ms_per_step"]), d.get("stage_ms"), d.get("roofline",{}).get("frac"))
for c in d.get("configs", []): print("   ", c["config"]["workload"][:40], "ms %.4f e2e %.4f"%(c["ms_per_step"], c["e2e"]["ms_per_step"]), c["gpu_launches"])
PY
done
