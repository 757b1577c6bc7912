"""Long offline run of the CPU certification model (tests/certification_model.py) against the oracle's per-ray scores:
many geometries, many particles, adversarial sine/cosine perturbation.  Prints one line per geometry and a total.
    python tools/certification_stress.py [cases] [repeats]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import port  # noqa: E402
import certification_model as cm  # noqa: E402
import test_certification_model as T  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
tot_e = tot_c = tot_w = 0
for case in range(cases):
    grid, cloud, r, th, t = T.make_case(case)
    ratios = (t - int(cloud["parent_pose"]["utime"][0])).astype(np.float64) / float(
        int(cloud["pose"]["utime"][0]) - int(cloud["parent_pose"]["utime"][0]))
    window = (0, 0, grid.width, grid.height) if case % 2 else cm.cloud_window(grid, cloud, r, 0.15)
    plan = cm.Plan(grid, r, th, ratios, 0.15, *window)
    if not plan.enabled:
        print(case, "float pass not applicable")
        continue
    fc = cm.derive_fast_map(grid.cells)
    pg = port.Grid(grid.cells, grid.origin_x, grid.origin_y, grid.cells_per_meter)
    rng = np.random.default_rng(10_000 + case)
    ev = ce = wr = 0
    want = [np.rint(2 * port.ray_scores(pg, cloud[i], r, th, t)).astype(np.int64) for i in range(len(cloud))]
    for _ in range(reps):
        for i in range(len(cloud)):
            v2, c = cm.fast_pass(grid, plan, cloud[i], r, th, ratios, 0.15, fc, rng)
            wr += int((c & (v2 != want[i])).sum()); ev += len(v2); ce += int(c.sum())
    print(case, f"{grid.width}x{grid.height} eps {plan.eps:.2e} band {2 * plan.kb}/{1 << plan.fb} certain {ce / ev:.3f} wrong {wr} evals {ev}", flush=True)
    tot_e += ev; tot_c += ce; tot_w += wr
print(f"TOTAL evals {tot_e} certain {tot_c / max(tot_e, 1):.3f} wrong {tot_w}")
