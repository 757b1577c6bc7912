#!/usr/bin/env python
"""Per-kernel figures for bench.py's roofline (profiles/r02_kernel_figures.json) from ncu summaries made by
tools/ncu_summary.py:   python tools/kernel_figures.py <key>=<summary.txt>:<evals in the launch> ...
Existing keys are kept unless given again."""
import json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "profiles", "r02_kernel_figures.json")
fig = json.load(open(path)) if os.path.exists(path) else {}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "": 1.0, "inst": 1.0}


def metric(text, name):
    m = re.search(r"^" + re.escape(name) + r"\s+(\S*?)\s+([0-9.eE+-]+)\s*$", text, re.M)
    if not m:
        raise SystemExit(f"{name} not found")
    return float(m.group(2)) * UNIT.get(m.group(1), 1.0)


for arg in sys.argv[1:]:
    key, rest = arg.split("=", 1)
    fn, evals = rest.rsplit(":", 1)
    text = open(os.path.join(ROOT, fn)).read()
    w = float(evals) / 32.0
    fig[key] = {
        "warp_inst_per_32_evals": round(metric(text, "smsp__inst_executed.sum") / w, 2),
        "smem_wavefronts_per_32_evals": round(metric(text, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / w, 2),
        "dram_bytes_per_launch": int(metric(text, "dram__bytes_read.sum") + metric(text, "dram__bytes_write.sum")),
        "issue_active_pct_ncu": metric(text, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "kernel_ms_ncu": metric(text, "gpu__time_duration.sum"),
        "source": f"{fn} ({int(float(evals))} evaluations in the launch, one B200)",
    }
json.dump(fig, open(path, "w"), indent=1)
print(json.dumps({k: v for k, v in fig.items() if not k.startswith("_")}, indent=1))
